#!/usr/bin/env bash
# One call: the whole GPU suite as the driver runs it, then the default bench. Logs in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv,noheader > gpurun_out/smi.txt 2>&1
t0=$(date +%s)
timeout ${SUITE_TIMEOUT:-1500} python -m pytest tests/ -x -q -m gpu > gpurun_out/t_all.log 2>&1
echo "suite exit $? in $(( $(date +%s) - t0 )) s: $(tail -n 1 gpurun_out/t_all.log)"
t0=$(date +%s)
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench exit $? in $(( $(date +%s) - t0 )) s"
python tools/show_bench.py gpurun_out/bench_default.json 2>&1 | head -30
