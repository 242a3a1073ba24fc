#!/usr/bin/env python
"""Synchronised per-phase wall times of full run() steps (headline config). Run on the GPU box."""
import sys, time, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bimodalattack_b200.attack as A
from bimodalattack_b200 import BimodalAttack, BimodalAttackConfig, synthetic as syn

dev = torch.device("cuda:0")
tok, proc, norm, image_cpu, messages, target, init = syn.headline_case(32000, 336)
model = syn.make_llava("7b", dtype=torch.bfloat16, device=dev, seed=0, head_gain=2.0, attn_implementation="sdpa")
cfg = BimodalAttackConfig(num_steps=6, search_width=512, pgd_attack=True, gcg_attack=True, joint_eval=True,
                          allow_non_ascii=True, filter_ids=False, optim_str_init=init, seed=1, verbosity="ERROR",
                          images_folder=None)
atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
acc = collections.defaultdict(list)

def timed(obj, name, label=None):
    fn = getattr(obj, name)
    def wrap(*a, **k):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = fn(*a, **k)
        torch.cuda.synchronize(); acc[label or name].append(time.perf_counter() - t0)
        return r
    setattr(obj, name, wrap)

for n in ("compute_gradient", "perform_pgd_step", "candidate_sampling", "_compute_candidates_loss_original", "_winner"):
    timed(atk, n)
timed(A, "image_features")
timed(tok, "batch_decode", "tok.batch_decode")
marks = []
def hook(i):
    torch.cuda.synchronize(); marks.append(time.perf_counter())
atk.step_hook = hook
atk.run(messages, "goal", target, image_cpu.to(dev))
steps = [b - a for a, b in zip(marks, marks[1:])]
print("step wall ms:", [round(1e3 * s, 1) for s in steps])
for k, v in acc.items():
    print(f"{k:40s} n={len(v):3d}  last3 ms: {[round(1e3*x,2) for x in v[-3:]]}")
