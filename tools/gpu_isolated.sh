#!/usr/bin/env bash
# Runs each GPU test function in its own process (a device trap poisons the CUDA context) with a
# per-test timeout; logs land in gpurun_out/. Usage: tools/gpu_isolated.sh tests/test_x.py [more ...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/smi.txt 2>&1
for f in "$@"; do
  for t in $(python -m pytest "$f" -m gpu --collect-only -q 2>/dev/null | grep "::" | sed 's/\[.*//' | sort -u); do
    name=$(echo "$t" | sed 's#[/:.]#_#g')
    echo "=== $t"
    timeout 300 python -m pytest "$t" -m gpu -x -q > "gpurun_out/$name.log" 2>&1
    echo "exit $? : $(tail -n 1 gpurun_out/$name.log)"
  done
done
