"""Summaries of gpurun_out/{launches.csv, prof_*.ncu-rep, bench_7b.json} for profiles/ (run in the
build container; ncu -i works without a GPU)."""
import collections, csv, io, json, subprocess, sys

def launches(path):
    lines = open(path).read().splitlines()
    hi = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in csv.DictReader(io.StringIO("\n".join(lines[hi:]))):
        try:
            val = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        agg[r["Kernel Name"].split("(")[0][:80]][0] += 1
        agg[r["Kernel Name"].split("(")[0][:80]][1] += val
    tot = sum(v[1] for v in agg.values())
    out = [f"one timed scoring step under ncu (cold caches, serialised): {sum(v[0] for v in agg.values())} launches, {tot/1e6:.1f} ms"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        out.append(f"{v[1]/tot*100:6.2f}%  n={v[0]:4d}  {v[1]/1e6:8.3f} ms  {k}")
    return "\n".join(out)

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "sm__cycles_elapsed.avg.per_second", "launch__shared_mem_per_block_dynamic"]

def rep(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        out.append("--- " + r[idx["Kernel Name"]][:90])
        for w in WANT:
            if w in idx:
                out.append(f"    {w} = {r[idx[w]]} {units[idx[w]]}")
    return "\n".join(out)

if __name__ == "__main__":
    print(launches("gpurun_out/launches.csv"))
    for p in sys.argv[1:]:
        print(rep(p))
