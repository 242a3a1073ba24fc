"""Host-side helpers with the reference's semantics (bimodalattack/utils.py)."""
from __future__ import annotations

from typing import List, Tuple

import torch

# initial characters for extra buffer entries (reference utils.py:8-12)
INIT_CHARS = [".", ",", "!", "?", ";", ":", "(", ")", "[", "]", "{", "}", "@", "#", "$", "%", "&", "*",
              "w", "x", "y", "z"]


def get_nonascii_toks(tokenizer, device="cpu") -> torch.Tensor:
    """Token ids that are not printable ASCII, plus bos/eos/pad/unk (reference utils.py:14-33)."""
    texts = tokenizer.batch_decode([[i] for i in range(tokenizer.vocab_size)])
    bad = [i for i, s in enumerate(texts) if not (s.isascii() and s.isprintable())]
    for tid in (tokenizer.bos_token_id, tokenizer.eos_token_id, tokenizer.pad_token_id,
                tokenizer.unk_token_id):
        if tid is not None:
            bad.append(tid)
    return torch.tensor(bad, device=device)


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
