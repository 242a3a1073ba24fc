"""ORACLE -- TEST INFRASTRUCTURE ONLY. Run in the BUILD CONTAINER (needs /root/reference):

    python -m oracle.make_golden

Executes the unmodified reference (``bimodalattack.run`` and its module-level helpers) on tiny
random-init models on CPU and writes small fixtures to tests/golden/. tests/test_oracle_golden.py
then checks oracle/hotpath.py against them without the reference present.
"""
from __future__ import annotations

import json
import logging
import os
import tempfile

import numpy as np
import torch
from PIL import Image

from oracle import cases, ref_shims

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def run_case(ref, family: str, mode: str) -> dict:
    model, tok, proc, norm, image = cases.build(family)
    ref_shims.wrap_image_features(model)
    kw = cases.config_kwargs(mode)
    with tempfile.TemporaryDirectory() as tmp:
        cfg = ref.BimodalAttackConfig(**kw, images_folder=tmp)
        res = ref.run(model, tok, proc, [{"role": "user", "content": cases.PROMPT}], "goal", cases.TARGET,
                      image.clone() if kw["pgd_attack"] else None, cfg, normalize=norm)
        last_png = None
        if kw["pgd_attack"]:
            arr = np.asarray(Image.open(os.path.join(tmp, f"{len(res.losses) - 1}.png")))
            last_png = arr.astype(np.int64).sum(axis=(0, 1)).tolist() + [int(arr[3, 5, 1]), int(arr[20, 11, 2])]
    return {"family": family, "mode": mode, "losses": res.losses, "strings": res.strings,
            "best_loss": res.best_loss, "best_string": res.best_string, "image_u8_checks": last_png}


def function_vectors(ref) -> dict:
    """Inputs/outputs of the reference's free functions on the CPU generator."""
    out = {}
    g = torch.Generator().manual_seed(11)
    V, L = 257, 6
    grad = torch.randn((L, V), generator=g)
    ids = torch.randint(0, V, (L,), generator=g)
    banned = torch.tensor([0, 1, 2, 3, 17, 200])
    torch.manual_seed(99)
    new_ids = ref.sample_ids_from_grad(ids.clone(), grad.clone(), 12, topk=16, n_replace=2, not_allowed_ids=banned)
    out["sample"] = {"seed_inputs": 11, "seed_rng": 99, "V": V, "L": L, "sw": 12, "topk": 16, "n_replace": 2,
                     "banned": banned.tolist(), "new_ids": new_ids.tolist()}
    g = torch.Generator().manual_seed(12)
    img = torch.rand((1, 3, 9, 7), generator=g)
    orig = (img + 0.3 * (torch.rand((1, 3, 9, 7), generator=g) - 0.5)).clamp(0, 1)
    gr = torch.randn((1, 3, 9, 7), generator=g)
    gr.view(-1)[::5] = 0.0

    class _Self:      # perform_pgd_step is a method that does not touch self
        pass
    stepped = ref.BimodalAttack.perform_pgd_step(_Self(), img.clone(), 64 / 255, 4 / 255, gr, orig)
    out["pgd"] = {"seed": 12, "shape": [1, 3, 9, 7], "eps": 64 / 255, "alpha": 4 / 255,
                  "out_hex": [float(v).hex() for v in stepped.detach().view(-1).tolist()]}
    ids, tok = cases.filter_case()
    kept = ref.filter_ids(ids.clone(), tok)
    out["filter"] = {"rows": int(ids.shape[0]), "kept": kept.tolist()}
    from bimodalattack_b200 import synthetic as syn
    out["nonascii"] = {"vocab": 1000, "ids": ref.get_nonascii_toks(syn.make_tokenizer(1000)).tolist()}
    return out


def main():
    logging.getLogger("gcg").setLevel(logging.ERROR)
    ref = ref_shims.import_reference()
    os.makedirs(OUT, exist_ok=True)
    runs = []
    for family in ("llava", "gemma"):
        for mode in cases.MODES:
            runs.append(run_case(ref, family, mode))
            print(family, mode, runs[-1]["losses"])
    with open(os.path.join(OUT, "reference_runs.json"), "w") as f:
        json.dump({"generated_by": "oracle/make_golden.py", "torch": torch.__version__, "runs": runs}, f, indent=1)
    with open(os.path.join(OUT, "reference_functions.json"), "w") as f:
        json.dump(function_vectors(ref), f, indent=1)


if __name__ == "__main__":
    main()
