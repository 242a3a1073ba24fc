"""ORACLE -- TEST INFRASTRUCTURE ONLY. Shared definitions of the tiny end-to-end cases used to pin
the oracle against the reference (make_golden.py) and to replay them (tests/test_oracle_golden.py)."""
from __future__ import annotations

import torch

PROMPT = "w10 w11 w12 w13"
TARGET = "w20 w21 w22 w23 w24"
INIT = "x x x x x x x x"

MODES = {
    "gcg_only": dict(pgd_attack=False, gcg_attack=True, joint_eval=False),
    "pgd_only": dict(pgd_attack=True, gcg_attack=False, joint_eval=False),
    "joint": dict(pgd_attack=True, gcg_attack=True, joint_eval=True),
    "pgd_gcg_nojoint": dict(pgd_attack=True, gcg_attack=True, joint_eval=False),
    # pgd_after_gcg=True is not a case: the reference itself raises TypeError at step 0
    # (bimodal_attack.py:661 formats current_loss while it is still None in that branch).
    "joint_masked_dynamic": dict(pgd_attack=True, gcg_attack=True, joint_eval=True, allow_non_ascii=False,
                                 dynamic_search=True, min_search_width=4, n_replace=2),
    # round 2: the option paths the product implements beyond the five mode combinations
    "joint_buffer3": dict(pgd_attack=True, gcg_attack=True, joint_eval=True, buffer_size=3),
    "gcg_buffer2": dict(pgd_attack=False, gcg_attack=True, joint_eval=False, buffer_size=2),
    "joint_early_stop": dict(pgd_attack=True, gcg_attack=True, joint_eval=True, early_stop=True),
    "joint_batch8": dict(pgd_attack=True, gcg_attack=True, joint_eval=True, batch_size=8),
    "gcg_nofilter_space": dict(pgd_attack=False, gcg_attack=True, joint_eval=False, filter_ids=False,
                               add_space_before_target=True),
    "joint_listinit2": dict(pgd_attack=True, gcg_attack=True, joint_eval=True, buffer_size=2, n_replace=3,
                            optim_str_init=["x x x x x x x x", "y y y z z y y y"]),
}
BASE = dict(num_steps=3, search_width=16, topk=32, seed=1, allow_non_ascii=True, alpha=4 / 255,
            eps=64 / 255, optim_str_init=INIT, verbosity="ERROR")


def build(family: str):
    """(model, tokenizer, processor, normalize, image) for 'llava' or 'gemma' tiny fp32 on CPU."""
    import torchvision.transforms as T
    from bimodalattack_b200 import synthetic as syn

    if family == "llava":
        tok = syn.make_tokenizer(1000)
        proc = syn.make_llava_processor(tok, 28)
        model = syn.make_llava("tiny", head_gain=1.0)
        norm = T.Normalize(syn.CLIP_MEAN, syn.CLIP_STD)
        size = 28
    elif family == "gemma":
        tok = syn.make_tokenizer(1024, image_token="<image_soft_token>",
                                 extra_tokens=["<start_of_image>", "<end_of_image>", "<start_of_turn>", "<end_of_turn>"])
        proc = syn.Gemma3Processor(tok)
        model = syn.make_gemma3("tiny")
        norm = T.Normalize(syn.GEMMA_MEAN, syn.GEMMA_STD)
        size = 56
    else:
        raise ValueError(family)
    g = torch.Generator().manual_seed(5)
    image = torch.rand((1, 3, size, size), generator=g)
    return model, tok, proc, norm, image


def config_kwargs(mode: str) -> dict:
    kw = dict(BASE)
    kw.update(MODES[mode])
    return kw


class MergingTokenizer:
    """Tokenizer whose decode -> encode round trip FAILS for ids in ``bad`` (they decode to the text of the next id),
    like BPE merges do for real vocabularies: what the reference's filter_ids (:166-186) exists for."""

    def __init__(self, tok, bad):
        self._tok, self._bad = tok, set(bad)

    def __getattr__(self, name):
        return getattr(self._tok, name)

    def __call__(self, *a, **k):
        return self._tok(*a, **k)

    def _fix(self, ids):
        return [int(i) + 1 if int(i) in self._bad else int(i) for i in ids]

    def decode(self, ids, **k):
        return self._tok.decode(self._fix(ids.tolist() if hasattr(ids, "tolist") else ids), **k)

    def batch_decode(self, seqs, **k):
        seqs = seqs.tolist() if hasattr(seqs, "tolist") else seqs
        return [self._tok.decode(self._fix(s), **k) for s in seqs]


def filter_case():
    """(ids [40, 8], tokenizer) of the filter_ids vector: every 17th id from 100 on breaks the round trip."""
    from bimodalattack_b200 import synthetic as syn
    tok = MergingTokenizer(syn.make_tokenizer(1000), bad=range(100, 1000, 17))
    g = torch.Generator().manual_seed(21)
    ids = torch.randint(16, 990, (40, 8), generator=g)
    return ids, tok
