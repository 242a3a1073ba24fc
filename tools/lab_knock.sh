#!/usr/bin/env bash
# knock-out experiments on the attention kernel (timing only; results are wrong by construction)
mkdir -p gpurun_out
{
for m in ${MODES:-1}; do for k in ${KNOCKS:-0}; do
  BMA_OCC_MODE=$m BMA_OCC_TIMING=1 BMA_OCC_KNOCK=$k timeout 120 python tools/bench_attention.py 512 31 616 32 2>&1 | tail -n 5
done; done
} > gpurun_out/attn_knock.log 2>&1
cat gpurun_out/attn_knock.log
