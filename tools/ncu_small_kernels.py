#!/usr/bin/env python
"""The HBM/latency-bound kernels of the path at the headline shapes, three launches each (the last one
is the warm one to read in an ncu capture): token-gradient GEMM (32 064 x 4096 table), one-launch
top-k + Philox replacement (20 x 32 064 bf16 gradient, topk 256, sw 512), two-level top-k (Gemma's
262 208-token rows), fused PGD step (336 px)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bimodalattack_b200 import native
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
V, d, L = 32064, 4096, 20
W = (torch.randn(V, d, generator=g) * 0.02).to(dev).bfloat16()
ge = (torch.randn(L, d, generator=g) * 1e-3).to(dev).bfloat16()
ids = torch.randint(0, V, (L,), generator=g).to(dev)
for _ in range(3):
    grad = native.token_grad(ge, W)
for _ in range(3):
    native.sample_ids(ids, grad, 512, 256, 1, None, 1, 0)
big = (torch.randn(L, 262208, generator=g) * 1e-3).to(dev).bfloat16()
ids2 = torch.randint(0, 262208, (L,), generator=g).to(dev)
for _ in range(3):
    native.sample_ids(ids2, big, 512, 256, 1, None, 1, 0)
x = torch.rand((1, 3, 336, 336), generator=g).to(dev)
gr = torch.randn((1, 3, 336, 336), generator=g).to(dev)
x0 = x.clone()
for _ in range(3):
    native.pgd_step_(x, gr, x0, 4 / 255 * 64 / 255, 64 / 255, [0.48, 0.45, 0.40], [0.26, 0.26, 0.27])
torch.cuda.synchronize()
print("ok")
