"""Worker of tests/test_gpu_multi.py: one rank of a 2-GPU run() on the tiny LLaVA model. Writes its
view of the result (every rank must end with the same suffix, losses and image) to argv[1]."""
import json
import os
import sys

import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_dir = sys.argv[1]
    buffer_size = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    import torchvision.transforms as T
    import bimodalattack_b200 as bma
    from bimodalattack_b200 import synthetic as syn
    tok = syn.make_tokenizer(1024)
    proc = syn.make_llava_processor(tok, 28)
    model = syn.make_llava("tiny128", dtype=torch.bfloat16, device=dev, head_gain=20.0, attn_implementation="sdpa")
    norm = T.Normalize(syn.CLIP_MEAN, syn.CLIP_STD)
    image = torch.rand((1, 3, 28, 28), generator=torch.Generator().manual_seed(5)).to(dev)
    cfg = bma.BimodalAttackConfig(num_steps=4, search_width=50, topk=32, seed=1, allow_non_ascii=True, alpha=4 / 255,
                                  eps=64 / 255, pgd_attack=True, gcg_attack=True, joint_eval=True,
                                  optim_str_init="x x x x x x x x x x", verbosity="ERROR", images_folder=None,
                                  buffer_size=buffer_size)
    atk = bma.BimodalAttack(model, tok, proc, cfg, normalize=norm)
    res = atk.run([{"role": "user", "content": "w10 w11 w12 w13 w14 w15"}], "goal", "w20 w21 w22 w23 w24", image.clone())
    with open(os.path.join(out_dir, f"rank{rank}_of{world}_b{buffer_size}.json"), "w") as f:
        json.dump({"losses": res.losses, "suffixes": res.adversarial_suffixes,
                   "image_sum": float(atk.final_image.detach().double().sum()),
                   "final_ids": atk.final_optim_ids.reshape(-1).tolist(), "first_string": res.strings[0]}, f)
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
