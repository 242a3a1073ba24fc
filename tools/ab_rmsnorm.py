#!/usr/bin/env python
"""In-process A/B of the RMSNorm kernel's occupancy switch (BMA_NORM_MINB = 3 | 4) at the headline shape
(16 023 rows x 4096): four input/output buffer pairs are rotated so every launch reads rows that are not in L2."""
import os, sys, json, statistics
os.environ["BMA_NORM_RELOAD"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bimodalattack_b200 import native
lib = native.load()
dev = torch.device("cuda:0")
rows, d = 16023, 4096
xs = [torch.randn(rows, d, device=dev).bfloat16() for _ in range(4)]
ys = [torch.empty_like(x) for x in xs]
w = torch.randn(d, device=dev).bfloat16()
st = torch.cuda.current_stream().cuda_stream
def run(i):
    native.check(lib.bma_rmsnorm(xs[i % 4].data_ptr(), d, w.data_ptr(), ys[i % 4].data_ptr(), d, rows, d, 1e-5, 0, st), "rmsnorm")
for i in range(200):
    run(i)
torch.cuda.synchronize()
res, ref = {"3": [], "4": []}, None
for rnd in range(5):
    for cfg in ("3", "4"):
        os.environ["BMA_NORM_MINB"] = cfg
        run(0); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(40):
            run(i)
        b.record(); b.synchronize()
        res[cfg].append(a.elapsed_time(b) / 40 * 1e3)
        if rnd == 0:
            o = ys[0].float().clone()
            if ref is None: ref = o
            else: print("# max |y4 - y3| =", (o - ref).abs().max().item())
nbytes = 2 * rows * d * 2
for cfg in ("3", "4"):
    med = statistics.median(res[cfg])
    print(json.dumps({"minb": cfg, "us_median": round(med, 2), "us_all": [round(x, 2) for x in res[cfg]], "GBps": round(nbytes / med / 1e3, 1)}))
