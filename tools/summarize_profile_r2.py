#!/usr/bin/env python
"""Text summaries of the round-2 ncu captures (gpurun_out/r2_*.raw.csv, r2_launches.csv) for profiles/.
Run in the build container (no GPU needed): python tools/summarize_profile_r2.py > profiles/r2_ncu_summary.txt"""
import collections, csv, io, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg.per_second"]


def launches(path):
    lines = open(path).read().splitlines()
    hi = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for r in csv.DictReader(io.StringIO("\n".join(lines[hi:]))):
        try:
            val = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        key = r["Kernel Name"].split("(")[0][:80]
        agg[key][0] += 1
        agg[key][1] += val
        n += 1
    tot = sum(v[1] for v in agg.values())
    out = [f"one timed scoring step under ncu --metrics gpu__time_duration.sum --clock-control none (cold caches, serialised; "
           f"compare SHARES): {n} launches, {tot / 1e6:.1f} ms"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{v[1] / tot * 100:6.2f}%  n={v[0]:4d}  {v[1] / 1e6:8.3f} ms  {k}")
    return "\n".join(out)


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def raw(path, pick_last_of_each=False):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    if pick_last_of_each:            # three launches per kernel: keep the warm one
        last = {}
        for r in body:
            last[r[idx["Kernel Name"]]] = r
        body = list(last.values())
    out, traffic = [], {}
    for r in body:
        name = r[idx["Kernel Name"]]
        out.append("--- " + name[:110])
        for w in WANT:
            if w in idx:
                out.append(f"    {w} = {r[idx[w]]} {units[idx[w]]}")
        if "dram__bytes_read.sum" in idx:
            t = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
                to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            out.append(f"    dram traffic (read + write) = {t / 1e6:.3f} MB")
            traffic[name.split("(")[0]] = t
    return "\n".join(out), traffic


if __name__ == "__main__":
    go = os.path.join(ROOT, "gpurun_out")
    print(launches(os.path.join(go, "r2_launches.csv")))
    print("\n=== --set full, first 10 launches of the candidate pass of one timed step (bench.py --steps 1) ===")
    txt, t1 = raw(os.path.join(go, "r2_prof_scorer.raw.csv"))
    print(txt)
    print("\n=== --set full, small kernels at the headline shapes (tools/ncu_small_kernels.py, third = warm launch) ===")
    txt, t2 = raw(os.path.join(go, "r2_prof_small.raw.csv"), pick_last_of_each=True)
    print(txt)
    print("\n=== --set full, first 10 launches of the native CLIP tower inside the same step ===")
    txt, t3 = raw(os.path.join(go, "r2_prof_step.raw.csv"))
    print(txt)
    json.dump({"scorer": t1, "small": t2}, open(os.path.join(go, "r2_traffic.json"), "w"), indent=1)
