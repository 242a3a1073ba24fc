#!/usr/bin/env python
"""Prints the interesting fields of bench.py JSON lines (files given as arguments)."""
import json
import sys

for f in sys.argv[1:]:
    for line in open(f):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        r = d.get("roofline") or {}
        print(f"== {f}: sw {d['config'].get('search_width')} n_gpus {d['n_gpus']}  value {d['value']:.1f} {d['unit']}  "
              f"{d['ms_per_step']:.2f} ms/step  e2e {d['e2e']['value']:.1f}  launches {d.get('gpu_launches')}  clocks {d.get('clocks')}")
        if r:
            ag = r.get("all_tcgen05_gemms") or {}
            print(f"   top kernel {r['achieved']:.1f} TF/s frac {r['frac']:.3f}  all gemms {ag.get('achieved')} share {r.get('share_of_scoring_kernels')}  step_alg {r.get('step_tflops_algorithmic'):.1f}  traffic {r.get('traffic')}"
                  f"  sharing {r.get('suffix_head_sharing')}")
            for k, v in r["per_class"].items():
                print(f"   {k:18s} {v['ms_per_step']:8.3f} ms  n={v['launches_per_step']:.0f}  {v['tflops']}")
        if d.get("cpu_baseline"):
            print("   cpu", d["cpu_baseline"])
        if d.get("full_step"):
            print("   full", {k: d["full_step"][k] for k in ("steps_per_sec", "ms_per_step", "phase_s_last")})
