#!/usr/bin/env python
"""BASELINE.json config 3: LLaVA-RC (robust CLIP ViT-L tower swapped into LLaVA-1.5-7B arch),
PGD-only, eps 64/255, alpha 4/255, synthetic image: the vision-tower-gradient path. A step is
gradient (forward + input-gradient backward through LLM and tower) -> fused PGD kernel -> batch-1
scoring of the new image. Prints one JSON line per image size (224 = the reference's RCLIP size,
336 = north_star's synthetic image size) with steps/s; the A100 reference is 1.93 steps/s
(LLaVA-RC PGD-only)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bimodalattack_b200 import BimodalAttack, BimodalAttackConfig, synthetic as syn

dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
model = syn.make_llava("7b", dtype=torch.bfloat16, device=dev, seed=0, head_gain=2.0, attn_implementation="sdpa")
for px in (224, 336):
    tok, proc, norm, image_cpu, messages, target, init = syn.headline_case(32000, px)
    syn.swap_in_robust_clip(model, image_size=px)
    cfg = BimodalAttackConfig(num_steps=steps, search_width=0, pgd_attack=True, gcg_attack=False, joint_eval=False,
                              allow_non_ascii=True, filter_ids=False, optim_str_init=init, seed=1, alpha=4 / 255,
                              eps=64 / 255, verbosity="ERROR", images_folder=None)
    atk = BimodalAttack(model, tok, proc, cfg, normalize=norm)
    marks = []
    def hook(i):
        torch.cuda.synchronize(); marks.append(time.perf_counter())
    atk.step_hook = hook
    res = atk.run(messages, "goal", target, image_cpu.to(dev))
    tail = marks[4:]                       # after the eager warm-up and the CUDA-graph captures
    dt = (tail[-1] - tail[0]) / (len(tail) - 1)
    final = atk.final_image.detach()
    print(json.dumps({"metric": "pgd_only_steps_per_sec", "value": 1.0 / dt, "unit": "steps/s", "ms_per_step": 1e3 * dt,
                      "n_gpus": 1, "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": f"LLaVA-RC (CLIP ViT-L/14 tower at {px}px swapped into LLaVA-1.5-7B arch, "
                                             f"random init) PGD-only, eps 64/255, alpha 4/255, {px // 14 * (px // 14)} image tokens",
                                 "timed_steps": len(tail) - 1},
                      "losses": [round(l, 4) for l in res.losses],
                      "linf": float((final - image_cpu.to(dev)).abs().max()), "eps": 64 / 255,
                      "a100_reference_published": {"steps_per_sec": 1.93, "note": "BASELINE.md: LLaVA-RC PGD-only, A100-80GB fp16"}}),
          flush=True)
    atk.engine.close()
