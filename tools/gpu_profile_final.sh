#!/usr/bin/env bash
# One gpurun call (final code of round 2): a plain run (must exit 0), then
#   1. the ncu launch list of one timed scoring step (NVTX range bench_timed),
#   2. ONE --set full capture of the first launches of the scorer's hot kernels inside that step (RMSNorm, QKV+RoPE
#      GEMM, attention_occ, o_proj GEMM, gate/up+GLU GEMM, down GEMM),
#   3. --set full of the attention kernel alone at the headline shape (tools/bench_attention.py).
# Raw-page CSVs are written next to the reports; tools/summarize_ncu.py turns them into profiles/*.txt.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --full-steps 0 --no-cpu-baseline --no-hbm-kernels"
$CMD > gpurun_out/r2f_prof_plain.json 2> gpurun_out/r2f_prof_plain.err || { echo "plain run failed"; tail -n 5 gpurun_out/r2f_prof_plain.err; exit 1; }
ncu --nvtx --nvtx-include "bench_timed/" --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --nvtx --nvtx-include "bench_timed/" --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"gemm_bf16_tn_kernel<256|attention_occ_kernel|rmsnorm_kernel" -c 8 -f -o gpurun_out/r2f_prof_scorer \
    $CMD > gpurun_out/r2f_ncu_scorer.log 2>&1
echo "scorer capture rc=$?"
python tools/bench_attention.py 512 31 616 32 > gpurun_out/r2f_attn_plain.log 2>&1 || { echo "attention plain run failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:attention_occ -s 5 -c 1 -f -o gpurun_out/r2f_prof_attn \
    python tools/bench_attention.py 512 31 616 32 > gpurun_out/r2f_ncu_attn.log 2>&1
echo "attention capture rc=$?"
for r in scorer attn; do
  ncu -i gpurun_out/r2f_prof_$r.ncu-rep --page raw --csv > gpurun_out/r2f_prof_$r.raw.csv 2>/dev/null
done
rm -f gpurun_out/r2f_prof_scorer.ncu-rep
ls -la gpurun_out/ | head -n 30
