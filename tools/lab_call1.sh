#!/usr/bin/env bash
# attention A/B (BMA_OCC_MODE 0 = two passes, 1 = scores in registers) + phase timing, tests, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
{
for m in 0 1; do
  BMA_OCC_MODE=$m timeout 120 python tools/bench_attention.py 512 31 616 32
  BMA_OCC_MODE=$m BMA_OCC_TIMING=1 timeout 120 python tools/bench_attention.py 512 31 616 32 2>&1 | tail -n 3
done
} > gpurun_out/attn_lab.log 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "attention" > gpurun_out/t_attn.log 2>&1
echo "attn tests exit $?"
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_mode1.json 2> gpurun_out/bench_mode1.err
echo "bench mode1 exit $?"
BMA_OCC_MODE=0 timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --full-steps 0 --no-hbm-kernels > gpurun_out/bench_mode0.json 2> gpurun_out/bench_mode0.err
echo "bench mode0 exit $?"
timeout 400 python -m pytest tests/test_gpu_attack.py tests/test_gpu_round2.py -m gpu -x -q > gpurun_out/t_attack.log 2>&1
echo "attack tests exit $?"
tail -n 3 gpurun_out/t_attn.log gpurun_out/t_attack.log
cat gpurun_out/attn_lab.log
