"""Segment-order tables of the candidate sequences (reference bimodal_attack.py:1150-1215) and the
prefix/tail split the B200 scorer works on.

A candidate sequence is a concatenation of named segments. Only ``optim`` differs between
candidates, so everything before it is a shared *prefix* (its K/V are computed once per step) and
everything from ``optim`` on is the per-candidate *tail*."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

_WITH_IMAGE = {
    True: ["before_img", "optim", "before_suffix", "image", "after"],      # Gemma-3 (:1153-1154)
    False: ["before_img", "image", "before_suffix", "optim", "after"],     # LLaVA   (:1155-1162)
}


def segment_order(mode: str, model_type: str, single: bool = False, no_joint_eval: bool = False,
                  no_target: bool = False) -> List[str]:
    gemma = model_type == "gemma3"
    if mode == "pgd":
        assert single, "PGD mode only supports single=True"
        return _WITH_IMAGE[gemma] + ["target"]
    if mode == "gcg":
        if single:
            return (["before_img", "optim", "before_suffix", "after", "target"] if gemma
                    else ["before_img", "before_suffix", "optim", "after", "target"])
        if no_joint_eval:
            return ["before", "optim", "after", "target"]
        if no_target:
            return ["before", "optim", "after"]
        raise ValueError("Invalid flags for BimodalAttack mode")
    if mode == "gcg_pgd":
        if not single and no_target:
            return list(_WITH_IMAGE[gemma])
        return _WITH_IMAGE[gemma] + ["target"]
    raise ValueError(f"Unknown mode '{mode}'")


@dataclass
class TailPlan:
    """How the scorer assembles one candidate: ``prefix`` [P, d] shared rows, then S tail rows where
    row s is the embedding of ``ids[b, slot[s]]`` if ``slot[s] >= 0`` else ``fixed[s]``."""
    prefix: Optional[torch.Tensor]   # [P, d] or None when the sequence starts with ``optim``
    slot: torch.Tensor               # [S] int32 (device)
    fixed: torch.Tensor              # [S, d] model dtype (device); rows at slot>=0 are unused
    n_prefix: int
    n_tail: int
    n_id_rows: int = 0               # one past the last tail row that depends on the candidate's ids


_SLOT_CACHE: Dict[tuple, torch.Tensor] = {}     # (device, slot tuple) -> int32 device tensor: no H2D copy per step


def plan_tail(order: Sequence[str], parts: Dict[str, torch.Tensor], n_optim: int) -> TailPlan:
    """``parts[name]`` is the [1, L, d] embedding of every fixed segment in ``order``."""
    if "optim" not in order:
        raise ValueError("segment order without an 'optim' segment")
    cut = list(order).index("optim")
    some = next(parts[n] for n in order if n != "optim")
    d, dev, dt = some.shape[-1], some.device, some.dtype
    pre = [parts[n].reshape(-1, d) for n in order[:cut]]
    prefix = torch.cat(pre, 0).contiguous() if pre else None
    slot: List[int] = []
    rows = []
    for name in order[cut:]:
        if name == "optim":
            slot += list(range(n_optim))
            rows.append(torch.zeros((n_optim, d), dtype=dt, device=dev))
        else:
            t = parts[name].reshape(-1, d)
            slot += [-1] * t.shape[0]
            rows.append(t)
    fixed = torch.cat(rows, 0).contiguous()
    n_id_rows = max(i for i, v in enumerate(slot) if v >= 0) + 1
    key = (str(dev), tuple(slot))
    slot_t = _SLOT_CACHE.get(key)
    if slot_t is None:
        if len(_SLOT_CACHE) > 64:
            _SLOT_CACHE.clear()
        slot_t = _SLOT_CACHE[key] = torch.tensor(slot, dtype=torch.int32, device=dev)
    return TailPlan(prefix=prefix, slot=slot_t, fixed=fixed,
                    n_prefix=0 if prefix is None else prefix.shape[0], n_tail=len(slot), n_id_rows=n_id_rows)
