#!/usr/bin/env bash
# attention A/B: BMA_OCC_MODE variants, timing
mkdir -p gpurun_out
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for m in ${MODES:-1}; do
  BMA_OCC_MODE=$m timeout 120 python tools/bench_attention.py 512 31 616 32
  BMA_OCC_MODE=$m BMA_OCC_TIMING=1 timeout 120 python tools/bench_attention.py 512 31 616 32 2>&1 | tail -n 2
done
} > gpurun_out/attn_lab.log 2>&1
cat gpurun_out/attn_lab.log
