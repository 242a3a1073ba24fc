"""Host-side helpers with the reference's semantics (bimodalattack/utils.py)."""
from __future__ import annotations

import functools
import gc
import inspect
from typing import Callable, List, Tuple

import torch

# initial characters for extra buffer entries (reference utils.py:8-12)
INIT_CHARS = [".", ",", "!", "?", ";", ":", "(", ")", "[", "]", "{", "}", "@", "#", "$", "%", "&", "*",
              "w", "x", "y", "z"]

_OOM_MARKERS = ("CUDA out of memory.", "cuDNN error: CUDNN_STATUS_NOT_SUPPORTED.",
                "DefaultCPUAllocator: can't allocate memory")


def get_nonascii_toks(tokenizer, device="cpu") -> torch.Tensor:
    """Token ids that are not printable ASCII, plus bos/eos/pad/unk (reference utils.py:14-33)."""
    texts = tokenizer.batch_decode([[i] for i in range(tokenizer.vocab_size)])
    bad = [i for i, s in enumerate(texts) if not (s.isascii() and s.isprintable())]
    for tid in (tokenizer.bos_token_id, tokenizer.eos_token_id, tokenizer.pad_token_id,
                tokenizer.unk_token_id):
        if tid is not None:
            bad.append(tid)
    return torch.tensor(bad, device=device)


def should_reduce_batch_size(exc: Exception) -> bool:
    return (isinstance(exc, RuntimeError) and len(exc.args) == 1
            and any(m in exc.args[0] for m in _OOM_MARKERS))


def find_executable_batch_size(function: Callable = None, starting_batch_size: int = 128):
    """Retry ``function(batch_size, ...)`` with a halved batch size after an out-of-memory error
    (reference utils.py:57-115). The native scorer has a fixed, pre-allocated workspace, so this
    only ever triggers for the torch-side phases; the interface is kept for drop-in use."""
    if function is None:
        return functools.partial(find_executable_batch_size, starting_batch_size=starting_batch_size)
    batch_size = starting_batch_size

    def wrapper(*args, **kwargs):
        nonlocal batch_size
        params = list(inspect.signature(function).parameters.keys())
        if len(params) < len(args) + 1:
            shown = ", ".join(f"{a}={v}" for a, v in zip(params[1:], args[1:]))
            raise TypeError(f"Batch size was passed into `{function.__name__}` as the first argument when "
                            f"called.Remove this as the decorator already does so: `{function.__name__}({shown})`")
        while True:
            if batch_size == 0:
                raise RuntimeError("No executable batch size found, reached zero.")
            try:
                return function(batch_size, *args, **kwargs)
            except Exception as exc:
                if not should_reduce_batch_size(exc):
                    raise
                gc.collect()
                if torch.cuda.is_available():
                    torch.cuda.empty_cache()
                batch_size //= 2
                print(f"Decreasing batch size to: {batch_size}")

    return wrapper


class AttackBuffer:
    """Best-suffix buffer (reference bimodal_attack.py:91-124); size 0 keeps only the latest entry."""

    def __init__(self, size: int):
        self.buffer: List[Tuple[float, torch.Tensor]] = []
        self.size = size

    def add(self, loss: float, optim_ids: torch.Tensor) -> None:
        if self.size == 0:
            self.buffer = [(loss, optim_ids)]
            return
        if len(self.buffer) < self.size:
            self.buffer.append((loss, optim_ids))
        else:
            self.buffer[-1] = (loss, optim_ids)
        self.buffer.sort(key=lambda e: e[0])

    def get_best_ids(self) -> torch.Tensor:
        return self.buffer[0][1]

    def get_lowest_loss(self) -> float:
        return self.buffer[0][0]

    def get_highest_loss(self) -> float:
        return self.buffer[-1][0]

    def log_buffer(self, tokenizer, logger) -> None:
        if not logger.isEnabledFor(20):      # decoding every entry is only worth it at INFO
            return
        lines = ["buffer:"]
        for loss, ids in self.buffer:
            text = tokenizer.batch_decode(ids)[0].replace("\\", "\\\\").replace("\n", "\\n")
            lines.append(f"loss: {loss} | string: {text}")
        logger.info("\n".join(lines))
