#!/usr/bin/env bash
mkdir -p gpurun_out
{
for m in ${MODES:-1}; do for s in ${SLEEPS:-0,0}; do
  echo "mode $m sleep $s"
  BMA_OCC_MODE=$m BMA_OCC_SLEEP=$s timeout 120 python tools/bench_attention.py 512 31 616 32 2>&1 | tail -n 1
done; done
} > gpurun_out/attn_sleep.log 2>&1
cat gpurun_out/attn_sleep.log
