#!/usr/bin/env bash
# One gpurun call (round 2): plain runs (must exit 0), then
#   1. the ncu launch list of one timed scoring step (NVTX range bench_timed),
#   2. ONE --set full capture of the first 10 launches of that step: row planner, tail-embedding gather,
#      RMSNorm, QKV+RoPE GEMM, prefix-K/V publish, attention, o_proj GEMM, RMSNorm, gate/up+GLU GEMM, down GEMM,
#   3. --set full captures of the small kernels at the headline shapes (tools/ncu_small_kernels.py).
# Raw-page CSVs are written next to the reports (the reports stay under gpurun's 64 MiB return limit).
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --full-steps 0 --no-cpu-baseline --no-hbm-kernels"
$CMD > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err || { echo "plain run failed"; tail -n 5 gpurun_out/r2_prof_plain.err; exit 1; }
ncu --nvtx --nvtx-include "bench_timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --nvtx --nvtx-include "bench_timed/" --set full --clock-control none --import-source on \
    -c 10 -f -o gpurun_out/r2_prof_step $CMD > gpurun_out/r2_ncu_step.log 2>&1
echo "step capture rc=$?"
python tools/ncu_small_kernels.py > gpurun_out/r2_small_plain.log 2>&1 || { echo "small plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"topk_|sample_replace|pgd_step|gemm_bf16_tn_kernel<32" -c 24 -f -o gpurun_out/r2_prof_small \
    python tools/ncu_small_kernels.py > gpurun_out/r2_ncu_small.log 2>&1
echo "small capture rc=$?"
for r in step small; do
  ncu -i gpurun_out/r2_prof_$r.ncu-rep --page raw --csv > gpurun_out/r2_prof_$r.raw.csv 2>/dev/null
done
ls -la gpurun_out/ | head -n 30
du -sm gpurun_out
