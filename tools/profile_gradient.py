#!/usr/bin/env python
"""Time split of the native gradient step at the headline size (LLaVA-1.5-7B arch, L = 638):
CUDA-graph replay of the whole step, the decoder forward+backward alone, the vision-tower part
alone, and the torch.profiler kernel table of one eager call (which kernels, how long). GPU box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bimodalattack_b200 import BimodalAttack, BimodalAttackConfig, synthetic as syn
from bimodalattack_b200.attack import install_patch_embed_gemm

dev = torch.device("cuda:0")
tok, proc, norm, image_cpu, messages, target, init = syn.headline_case(32000, 336)
model = syn.make_llava("7b", dtype=torch.bfloat16, device=dev, seed=0, head_gain=2.0, attn_implementation="sdpa")
cfg = BimodalAttackConfig(num_steps=1, search_width=512, pgd_attack=True, gcg_attack=True, joint_eval=True,
                          allow_non_ascii=True, filter_ids=False, optim_str_init=init, seed=1, verbosity="ERROR",
                          images_folder=None)
atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
atk._prepare(messages, target)
install_patch_embed_gemm(model)
torch.use_deterministic_algorithms(True, warn_only=True)
torch.utils.deterministic.fill_uninitialized_memory = False
image = image_cpu.to(dev)
optim_ids = tok(init, add_special_tokens=False, return_tensors="pt")["input_ids"].to(dev)


def timeit(fn, label, n=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1)/n:.2f} ms GPU-event, {(time.perf_counter()-t0)/n*1e3:.2f} ms wall", flush=True)


full = lambda: atk.compute_gradient(optim_ids, image.clone().requires_grad_())
timeit(full, "compute_gradient (graph replay)", n=8)
ng = atk._native_gradient()
print("native:", ng is not None, getattr(atk, "_native_grad_why", None))
if ng is not None:
    x0 = torch.randn(638, ng.d, device=dev).to(torch.bfloat16) * 0.02
    dec = lambda: ng.forward_backward(x0, atk.target_ids)
    timeit(dec, "decoder forward+backward, eager launches")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        dec()
    timeit(g.replay, "decoder forward+backward, graph replay", n=8)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        dec(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))
from bimodalattack_b200.attack import image_features
def tower():
    img = image.clone().requires_grad_()
    f = image_features(model, atk._normalize_pixels(img), False)
    return torch.autograd.grad(f, img, grad_outputs=torch.ones_like(f))[0]
timeit(tower, "vision tower + projector forward+backward (HF autograd, eager)")
