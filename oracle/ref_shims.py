"""ORACLE -- TEST INFRASTRUCTURE ONLY.

Imports the UNMODIFIED reference from /root/reference (build container only; the path does not
exist on the GPU box) and adapts the environment around it -- never the reference itself:
transformers >= 5 returns BaseModelOutputWithPooling from ``get_image_features`` where 4.50.2 (the
reference's pin, requirements.txt:4) returned a tensor, so the model object handed to the
reference gets a wrapped method. Used by ``oracle/make_golden.py`` to pin ``oracle/hotpath.py``.
"""
from __future__ import annotations

import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("BMA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "bimodalattack"))


def import_reference():
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import bimodalattack.bimodal_attack as ref   # noqa: E402
    return ref


def wrap_image_features(model):
    """Give ``model.get_image_features`` its transformers-4.50 return type (a tensor)."""
    inner = model.get_image_features
    if getattr(inner, "_bma_wrapped", False):
        return model

    def tensor_features(*args, **kwargs):
        out = inner(*args, **kwargs)
        if isinstance(out, torch.Tensor):
            return out
        pooled = out.pooler_output
        return torch.stack(list(pooled), 0) if isinstance(pooled, (list, tuple)) else pooled

    tensor_features._bma_wrapped = True
    model.get_image_features = tensor_features
    return model
