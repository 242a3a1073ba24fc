#!/usr/bin/env python
"""Token-gradient GEMM alone, back to back (W = 263 MB > L2): achieved HBM GB/s for the given
BMA_TOKGRAD_STAGGER (0 = all tiles walk K in the same order)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bimodalattack_b200 import native
dev = torch.device("cuda:0")
for V, d in ((32064, 4096), (262208, 2560)):
    W = (torch.randn(V, d, device=dev) * 0.02).bfloat16()
    g = (torch.randn(20, d, device=dev) * 1e-3).bfloat16()
    for _ in range(5):
        native.token_grad(g, W)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        native.token_grad(g, W)
    b.record(); b.synchronize()
    us = a.elapsed_time(b) / 20 * 1e3
    nbytes = 2 * (V * d + 20 * d + 20 * V)
    print(json.dumps({"stagger": os.environ.get("BMA_TOKGRAD_STAGGER", "default(7)"), "V": V, "d": d, "us": us,
                      "GBps": nbytes / us / 1e3}))
    del W
