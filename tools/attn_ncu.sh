#!/usr/bin/env bash
# ncu --set full of one attention launch (B=512 candidates x S=31 rows, P=616 prefix keys, 32 heads) for the three
# head_dim-128 kernels (BMA_ATTN = occ (default) | split | tc); raw CSV pages land in gpurun_out/.
set -u
mkdir -p gpurun_out
for K in occ split tc; do
  BMA_ATTN=$K ncu --set full --clock-control none -k regex:attention -s 5 -c 1 -o gpurun_out/attn_$K -f \
    python tools/bench_attention.py 512 31 616 32 > gpurun_out/attn_$K.log 2>&1
  echo "$K rc=$?"; tail -n 1 gpurun_out/attn_$K.log
  ncu -i gpurun_out/attn_$K.ncu-rep --page raw --csv > gpurun_out/attn_$K.raw.csv 2>/dev/null
done
