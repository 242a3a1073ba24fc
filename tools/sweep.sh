#!/usr/bin/env bash
# BASELINE.json config 5: search-width sweep of the headline workload on the GPUs of this box.
# Usage: tools/sweep.sh [n_gpus] ; one JSON line per search width into gpurun_out/sweep_n<N>.jsonl
set -u
N="${1:-1}"
mkdir -p gpurun_out
OUT="gpurun_out/sweep_n${N}.jsonl"
: > "$OUT"
for SW in 128 256 512 1024 2048 4096; do
  if [ "$N" -gt 1 ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus "$N" --steps 3 --warmup 3 --search-width "$SW" --full-steps 0 --no-cpu-baseline 2>> gpurun_out/sweep.err | tail -n 1 >> "$OUT"
  else
    python bench.py --steps 3 --warmup 3 --search-width "$SW" --full-steps 0 --no-cpu-baseline 2>> gpurun_out/sweep.err | tail -n 1 >> "$OUT"
  fi
  echo "sw=$SW rc=$?"
done
