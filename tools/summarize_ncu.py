#!/usr/bin/env python
"""Text summary of tools/gpu_profile_final.sh's captures (gpurun_out/r2f_*) for profiles/.
Run in the build container (no GPU needed): python tools/summarize_ncu.py > profiles/r2_final_ncu_summary.txt"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from summarize_profile_r2 import launches, raw, ROOT, WANT

WANT += ["smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
         "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
         "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]

if __name__ == "__main__":
    go = os.path.join(ROOT, "gpurun_out")
    print(launches(os.path.join(go, "r2f_launches.csv")))
    print("\n=== --set full, first launches of the scorer's hot kernels inside one timed step (bench.py --steps 1) ===")
    txt, t1 = raw(os.path.join(go, "r2f_prof_scorer.raw.csv"))
    print(txt)
    print("\n=== --set full, attention_occ alone (tools/bench_attention.py 512 31 616 32: 15 872 rows, 616 prefix keys, 32 heads) ===")
    txt, t2 = raw(os.path.join(go, "r2f_prof_attn.raw.csv"))
    print(txt)
    json.dump({"scorer": t1, "attention": t2}, open(os.path.join(go, "r2f_traffic.json"), "w"), indent=1)
