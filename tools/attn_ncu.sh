#!/usr/bin/env bash
# ncu --set full of one attention launch (B=512 candidates x S=31 rows, P=616 prefix keys, 32 heads)
# for both head_dim-128 kernels; raw CSV pages land in gpurun_out/.
set -u
mkdir -p gpurun_out
for SPLIT in 1 0; do
  BMA_ATTN_SPLIT=$SPLIT ncu --set full --clock-control none -k regex:attention -c 1 -o gpurun_out/attn_split$SPLIT -f \
    python tools/atc_bisect.py 512 31 616 32 > gpurun_out/attn_split$SPLIT.log 2>&1
  echo "split=$SPLIT rc=$?"; tail -n 1 gpurun_out/attn_split$SPLIT.log
done
