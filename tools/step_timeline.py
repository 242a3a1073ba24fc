#!/usr/bin/env python
"""Host and device timeline of ONE scoring step of the headline config (1 GPU), without extra
synchronisation: perf_counter marks where the host is, CUDA events where the device is, at the
boundaries features | embeds+plan | prefill | score | winner. Answers "where do the milliseconds
outside the scorer's kernels go". Usage: step_timeline.py [search_width]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bimodalattack_b200 import BimodalAttack, BimodalAttackConfig, synthetic as syn

sw = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
tok, proc, norm, image_cpu, messages, target, init = syn.headline_case(32000, 336)
model = syn.make_llava("7b", dtype=torch.bfloat16, device=dev, seed=0, head_gain=2.0, attn_implementation="sdpa")
cfg = BimodalAttackConfig(num_steps=1, search_width=sw, topk=256, n_replace=1, pgd_attack=True, gcg_attack=True,
                          joint_eval=True, allow_non_ascii=True, filter_ids=False, optim_str_init=init, seed=1,
                          alpha=4 / 255, eps=64 / 255, verbosity="ERROR", images_folder=None)
atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
atk._prepare(messages, target)
image = image_cpu.to(dev)
torch.manual_seed(1)
optim_ids = tok(init, add_special_tokens=False, return_tensors="pt")["input_ids"].to(dev)
img = image.clone().requires_grad_()
g_tok, g_img = atk.compute_gradient(optim_ids, img)
image = atk.perform_pgd_step(img, cfg.eps, cfg.alpha, g_img, image.clone()).detach()
sampled, _, _, _ = atk.candidate_sampling(cfg, 0, optim_ids, g_tok, tok, [], 0.0)
atk._base = (sampled, optim_ids.reshape(-1).to(torch.int64).contiguous())

names = ["features", "embeds", "loss_fn", "winner"]


def step(record):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    t = [time.perf_counter()]
    ev[0].record()
    with torch.no_grad():
        atk.engine.prefill_tag = None
        feats = atk.scoring_features(image)
        t.append(time.perf_counter()); ev[1].record()
        batch = atk._build_input_embeds_pgd(sampled, feats, search_width=sw, single=True)
        t.append(time.perf_counter()); ev[2].record()
        atk._compute_candidates_loss_original(sw, batch)
        t.append(time.perf_counter()); ev[3].record()
    w = atk._winner()
    t.append(time.perf_counter()); ev[4].record()
    torch.cuda.synchronize()
    if record:
        host = [1e3 * (b - a) for a, b in zip(t, t[1:])]
        gpu = [ev[i].elapsed_time(ev[i + 1]) for i in range(len(names))]
        print("  ".join(f"{n}: host {h:6.3f} gpu {g:7.3f}" for n, h, g in zip(names, host, gpu)),
              f" | total host {1e3 * (t[-1] - t[0]):.3f} gpu {ev[0].elapsed_time(ev[4]):.3f}")
    return w


for i in range(8):
    step(i >= 4)
