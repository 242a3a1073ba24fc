#!/usr/bin/env bash
# bench_attention under a list of environment settings: ENVS="A=1,B=2 A=3" tools/lab_env.sh
mkdir -p gpurun_out
{
for e in ${ENVS:-X=0}; do
  echo "== $e"
  env $(echo "$e" | tr ',' ' ') timeout 120 python tools/bench_attention.py ${SHAPE:-512 31 616 32} 2>&1 | tail -n ${TAILN:-1}
done
} > gpurun_out/attn_env.log 2>&1
cat gpurun_out/attn_env.log
