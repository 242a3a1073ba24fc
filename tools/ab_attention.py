#!/usr/bin/env python
"""In-process A/B of the attention kernel's tuning switches (same box, same clocks): after a long warm-up
every configuration is timed ROUNDS times in rotation, 40 launches each. Usage:
  ab_attention.py B S P H "DEFER=1,PVFIRST=1" "DEFER=0" "ATTN=split" ...
Keys: DEFER, PVFIRST (BMA_OCC_<KEY>, re-read per launch through BMA_OCC_RELOAD) -- ATTN=split|tc selects the older
kernels and is read once per process, so give it its own invocation."""
import os, sys, json, statistics
os.environ["BMA_OCC_RELOAD"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bimodalattack_b200 import native
lib = native.load()
dev = torch.device("cuda:0")
B, S, P, H = [int(x) for x in sys.argv[1:5]]
configs = sys.argv[5:] or ["DEFER=1"]
D, M = 128, B * S
g = torch.Generator().manual_seed(1)
mk = lambda r, c: torch.randn((r, c), generator=g).to(dev).bfloat16()
q, k, kp = mk(M, H * D), mk(M, H * D), mk(P, H * D)
ldt, ldp = (M + 7) // 8 * 8, (P + 7) // 8 * 8
vT, vpT = mk(H * D, ldt), mk(H * D, ldp)
out = torch.zeros((M, H * D), dtype=torch.bfloat16, device=dev)
st = torch.cuda.current_stream().cuda_stream
KEYS = ["DEFER", "PVFIRST", "PACKED"]
def apply(cfg):
    for key in KEYS:
        os.environ.pop("BMA_OCC_" + key, None)
    for kv in cfg.split(","):
        key, val = kv.split("=")
        os.environ["BMA_OCC_" + key] = val
def run():
    native.check(lib.bma_attention(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), vT.data_ptr(), vT.stride(0),
                                   kp.data_ptr(), kp.stride(0), vpT.data_ptr(), vpT.stride(0), out.data_ptr(), out.stride(0),
                                   M, S, P, H, 1, D, D ** -0.5, st), "attention")
apply(configs[0])
for _ in range(400):
    run()
torch.cuda.synchronize()
res = {c: [] for c in configs}
ref = None
for rnd in range(int(os.environ.get("ROUNDS", "5"))):
    for c in configs:
        apply(c)
        run(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(40):
            run()
        b.record(); b.synchronize()
        res[c].append(a.elapsed_time(b) / 40 * 1e3)
        if rnd == 0:
            o = out.float().clone()
            if ref is None:
                ref = o
            else:
                print(f"# {c}: max |out - out[{configs[0]}]| = {(o - ref).abs().max().item():.4g}")
flops = 4.0 * M * (P + S / 2) * H * D
for c in configs:
    med = statistics.median(res[c])
    print(json.dumps({"config": c, "us_median": round(med, 1), "us_all": [round(x, 1) for x in res[c]], "TFLOPs": round(flops / med / 1e6, 1)}))
