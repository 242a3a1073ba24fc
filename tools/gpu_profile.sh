#!/usr/bin/env bash
# One gpurun call: plain bench run (must exit 0), then the ncu launch list of one timed scoring
# step and full-set captures of tcgen05 GEMM launches from the candidate pass (the first ~125 GEMM
# launches of a step belong to the 596-row prefix pass) and of one attention launch.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --full-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err &&
ncu --nvtx --nvtx-include "bench_timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --nvtx --nvtx-include "bench_timed/" --set full --clock-control none --import-source on \
    -k regex:gemm_bf16_tn -s 133 -c 5 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
ncu --nvtx --nvtx-include "bench_timed/" --set full --clock-control none --import-source on \
    -k regex:attention -s 40 -c 1 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn capture rc=$?"
