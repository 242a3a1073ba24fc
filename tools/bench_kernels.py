#!/usr/bin/env python
"""Stand-alone timings of the HBM-bound kernels of the path against the measured copy bandwidth
(MEASURED_PEAKS.json): token-gradient GEMM (W streamed once), top-k + replacement sampler, fused PGD
step, for the LLaVA-1.5-7B and Gemma-3-4b shapes of SURVEY.md section 8(d). CUDA events on the
launching stream, 5 warm-up + 50 timed launches, a 512 MB write between launches so every launch
starts with a cold L2. One JSON line per kernel. Run on the GPU box."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bimodalattack_b200 import native

dev = torch.device("cuda:0")
lib = native.load()
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=50):
    for _ in range(5):
        fn()
    evs = []
    for _ in range(n):
        flush.fill_(1)                      # evict L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def report(name, shape, nbytes, fn):
    med, best = timed(fn)
    print(json.dumps({"kernel": name, "shape": shape, "algorithmic_bytes": nbytes, "median_us": 1e3 * med,
                      "best_us": 1e3 * best, "achieved_gbs": nbytes / (med * 1e-3) / 1e9,
                      "peak_gbs": peaks["hbm_gbs"], "frac": nbytes / (med * 1e-3) / 1e9 / peaks["hbm_gbs"],
                      "floor_us_at_peak": nbytes / peaks["hbm_gbs"] / 1e3, "l2": "flushed between launches"}), flush=True)


g = torch.Generator().manual_seed(0)
for model, V, d in (("llava-1.5-7b", 32064, 4096), ("gemma-3-4b", 262208, 2560)):
    L, topk, sw = 20, 256, 512
    W = (torch.randn((V, d), generator=g) * 0.02).to(dev).bfloat16()
    ge = torch.randn((L, d), generator=g).to(dev).bfloat16()
    report("token_grad (tcgen05 GEMM, W streamed once)", f"{model}: [{L},{d}] x [{V},{d}]^T", (V * d + L * d + L * V) * 2,
           lambda: native.token_grad(ge, W))
    grad = torch.randn((L, V), generator=g).to(dev).bfloat16()
    ids = torch.randint(0, V, (L,), generator=g).to(dev)
    report("sample_ids (top-k + Philox replacement)", f"{model}: grad [{L},{V}] bf16, topk {topk}, sw {sw}",
           L * V * 2 + L * topk * 8 + sw * L * 8, lambda: native.sample_ids(ids, grad, sw, topk, 1, None, 1, 0))
    del W
for px in (224, 336, 896):
    x = torch.rand((1, 3, px, px), generator=g).to(dev)
    gr = torch.randn((1, 3, px, px), generator=g).to(dev)
    x0 = x.clone()
    report("pgd_step (fused sign/step/project/clamp)", f"image {px}x{px} fp32", 4 * 3 * px * px * 4,
           lambda: native.pgd_step_(x, gr, x0, 4 / 255 * 64 / 255, 64 / 255))
