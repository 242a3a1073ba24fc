#!/usr/bin/env python
"""Where does phase A (HF forward + input-gradient backward at batch 1) spend its time?
CUDA-event time per call eager / deterministic-mode / CUDA-graph, plus the torch.profiler kernel
table of the deterministic eager call. Run on the GPU box."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("BMA_GRAPH_DEBUG", "1")
import torch
from bimodalattack_b200 import BimodalAttack, BimodalAttackConfig, synthetic as syn

dev = torch.device("cuda:0")
tok, proc, norm, image_cpu, messages, target, init = syn.headline_case(32000, 336)
model = syn.make_llava("7b", dtype=torch.bfloat16, device=dev, seed=0, head_gain=2.0, attn_implementation="sdpa")
cfg = BimodalAttackConfig(num_steps=1, search_width=512, pgd_attack=True, gcg_attack=True, joint_eval=True,
                          allow_non_ascii=True, filter_ids=False, optim_str_init=init, seed=1, verbosity="ERROR",
                          images_folder=None)
atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
atk._prepare(messages, target)
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

eager = lambda: atk._gradient_eager(optim_ids, image.clone().requires_grad_())
timeit(eager, "eager")
torch.use_deterministic_algorithms(True, warn_only=True)
timeit(eager, "eager, deterministic mode (fill_uninitialized_memory on)")
torch.utils.deterministic.fill_uninitialized_memory = False
timeit(eager, "eager, deterministic mode, no fill")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    eager(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
graphed = lambda: atk.compute_gradient(optim_ids, image.clone().requires_grad_())
timeit(graphed, "compute_gradient (CUDA graph after warm-up), deterministic mode")
print("graph state:", {k: (v if k in ("calls", "failed") else type(v).__name__) for k, v in (atk._grad_graph or {}).items()})
torch.use_deterministic_algorithms(False)
atk._grad_graph = None
timeit(graphed, "compute_gradient (CUDA graph after warm-up), default mode")
