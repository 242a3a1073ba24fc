#!/usr/bin/env bash
# One gpurun call (round 2): plain bench run (must exit 0), then
#   1. the ncu launch list of one timed scoring step (NVTX range bench_timed),
#   2. --set full captures of one launch of each scorer GEMM epilogue (QKV+RoPE, o_proj, gate/up+GLU, down)
#      and of one attention launch inside the timed step,
#   3. --set full captures of the HBM/latency-bound kernels of the path: RMSNorm, token-gradient GEMM,
#      one-launch top-k, Philox replacement, PGD step, CE finalise, row planner, tail-embedding gather.
# Reports land in gpurun_out/ (r2_*.ncu-rep, r2_launches.csv); tools/summarize_profile_r2.py turns them
# into the text files under profiles/.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --full-steps 0 --no-cpu-baseline"
$CMD > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err || { echo "plain run failed"; tail -n 5 gpurun_out/r2_prof_plain.err; exit 1; }
ncu --nvtx --nvtx-include "bench_timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --nvtx --nvtx-include "bench_timed/" --set full --clock-control none --import-source on \
    -k regex:gemm_bf16_tn -s 4 -c 4 -f -o gpurun_out/r2_prof_gemm $CMD > gpurun_out/r2_ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
ncu --nvtx --nvtx-include "bench_timed/" --set full --clock-control none --import-source on \
    -k regex:attention -s 2 -c 1 -f -o gpurun_out/r2_prof_attn $CMD > gpurun_out/r2_ncu_attn.log 2>&1
echo "attn capture rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"rmsnorm_kernel|topk_row16|sample_replace|pgd_step|ce_finalize|plan_rows|build_tail_embeds|gemm_bf16_tn_kernel<32" \
    -s 40 -c 60 -f -o gpurun_out/r2_prof_hbm $CMD > gpurun_out/r2_ncu_hbm.log 2>&1
echo "hbm capture rc=$?"
ls -la gpurun_out/r2_prof_*.ncu-rep 2>/dev/null
