#!/usr/bin/env python
"""Attention kernel alone at the headline shape (B candidates x S tail rows, P prefix keys, H heads of
128): 20 back-to-back launches between CUDA events. BMA_ATTN = occ (default) | split | tc selects the
head_dim-128 kernel. Usage: bench_attention.py B S P H"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bimodalattack_b200 import native
lib = native.load()
dev = torch.device("cuda:0")
B, S, P, H = [int(x) for x in sys.argv[1:5]]
D, M = 128, B * S
g = torch.Generator().manual_seed(1)
mk = lambda r, c: torch.randn((r, c), generator=g).to(dev).bfloat16()
q, k, kp = mk(M, H * D), mk(M, H * D), mk(P, H * D)
ldt, ldp = (M + 7) // 8 * 8, (P + 7) // 8 * 8
vT, vpT = mk(H * D, ldt), mk(H * D, ldp)
out = torch.zeros((M, H * D), dtype=torch.bfloat16, device=dev)
st = torch.cuda.current_stream().cuda_stream
def run():
    native.check(lib.bma_attention(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), vT.data_ptr(), vT.stride(0),
                                   kp.data_ptr(), kp.stride(0), vpT.data_ptr(), vpT.stride(0), out.data_ptr(), out.stride(0),
                                   M, S, P, H, 1, D, D ** -0.5, st), "attention")
for _ in range(3):
    run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    run()
b.record(); b.synchronize()
us = a.elapsed_time(b) / 20 * 1e3
flops = 4.0 * M * (P + S / 2) * H * D
print(json.dumps({"kernel": os.environ.get("BMA_ATTN", "occ"), "B": B, "S": S, "P": P, "H": H, "us": us, "TFLOPs": flops / us / 1e6}))
