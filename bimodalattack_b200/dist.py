"""Candidate sharding across the GPUs of one box.

Candidates are independent given (weights, prefix KV, target): rank r scores the contiguous slice
``shard_bounds(n, world)`` against a replicated model and the winner is picked by ONE all-reduce
(MIN) over a packed 64-bit key ``orderable(loss) << 31 | global_index`` -- ties resolve to the
lowest index exactly like ``torch.argmin``. Gradient, PGD and sampling run on rank 0 and their
(small) results are broadcast. Works with the nccl backend on GPUs and with gloo in the CPU tests.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as td


@dataclass(frozen=True)
class World:
    rank: int = 0
    size: int = 1


def world() -> World:
    if td.is_available() and td.is_initialized():
        return World(td.get_rank(), td.get_world_size())
    return World()


def shard_bounds(n: int, w: World) -> Tuple[int, int]:
    """Contiguous, balanced slice of ``range(n)`` owned by ``w.rank`` (first ranks get the extras)."""
    base, extra = divmod(n, w.size)
    lo = w.rank * base + min(w.rank, extra)
    return lo, lo + base + (1 if w.rank < extra else 0)


def balanced_cuts(cost: torch.Tensor, size: int, share: Optional[Sequence[float]] = None) -> list:
    """``size + 1`` cut points of ``range(len(cost))`` into CONTIGUOUS shards of near-equal total cost -- or, with
    ``share`` (one positive weight per rank, e.g. the ranks' measured speeds), of total cost proportional to it
    (``cost[i]`` = decoder rows candidate i will run: with suffix-head sharing a candidate replaced at
    position j owns only the tail rows from j on, so equal COUNTS give the ranks unequal work and the
    slowest rank sets the step time). Deterministic in ``cost``: every rank computes the same cuts from the
    broadcast ids. One small D2H copy when ``cost`` lives on a device."""
    n = int(cost.shape[0])
    if size <= 1 or n == 0:
        return [0] + [n] * max(size, 1)
    cum = torch.cumsum(cost.to(torch.int64), 0)
    k = torch.arange(1, size, device=cost.device, dtype=torch.int64)
    # shard r ends at the first candidate whose running cost reaches r/size (or the ranks' cumulative share) of the total
    if share is None:
        target = (cum[-1] * k + size - 1) // size
    else:
        tot, acc, frac = float(sum(share)), 0.0, []
        for x in list(share)[:-1]:
            acc += float(x)
            frac.append(acc / tot)
        total = int(cum[-1].item())
        target = torch.tensor([min(total, max(0, math.ceil(f * total))) for f in frac], device=cost.device, dtype=torch.int64)
    cuts = (torch.searchsorted(cum, target, right=False) + 1).clamp_(max=n)
    out = [0] + [int(c) for c in cuts.tolist()] + [n]
    for i in range(1, len(out)):           # monotone (zero-cost runs can make searchsorted tie)
        out[i] = max(out[i], out[i - 1])
    return out


def updated_shares(share: Sequence[float], step_ms: Sequence[float], wait_ms: Sequence[float],
                   damping: float = 0.5, spread: float = 1.25) -> list:
    """Straggler-aware shares for the next step. Rank r did ``share[r]`` of the candidate rows in ``step_ms[r] -
    wait_ms[r]`` of its own time (the rest it spent waiting for other ranks in the per-layer prefix exchange), so its
    speed is ``share / busy`` and the shares that would let all ranks finish together are proportional to the speeds.
    Damped (the replicated part of a rank's work -- its prefix slice, the vision tower -- does not move with the
    share) and clamped to ``[1/spread, spread]`` times the equal share. Pure function of its arguments: every rank
    computes the same list from the all-gathered inputs."""
    n = len(share)
    busy = [max(float(t) - float(w), 1e-3) for t, w in zip(step_ms, wait_ms)]
    speed = [float(s) / b for s, b in zip(share, busy)]
    tot = sum(speed)
    new = [(1.0 - damping) * float(s) + damping * v / tot for s, v in zip(share, speed)]
    new = [min(max(x, 1.0 / (spread * n)), spread / n) for x in new]
    tot = sum(new)
    return [x / tot for x in new]


def pack_key(loss: float, index: int) -> int:
    """Host mirror of the device packing (csrc/scoring_ops.cu:pack_loss_key)."""
    import struct
    bits = struct.unpack("<I", struct.pack("<f", loss))[0]
    if loss != loss:
        order = 0xFFFFFFFF
    else:
        order = (~bits) & 0xFFFFFFFF if bits & 0x80000000 else bits | 0x80000000
    return (order << 31) | (index & 0x7FFFFFFF)


_native_comm = None          # (handle, world size) of the library-owned NCCL communicator


def native_comm(device, w: Optional[World] = None):
    """The C library's own NCCL communicator for this process (created once; None on one rank or on
    CPU). Rank 0 makes the 128-byte unique id, torch.distributed carries it to the other ranks."""
    global _native_comm
    w = w or world()
    if w.size == 1 or torch.device(device).type != "cuda":
        return None
    if _native_comm is not None and _native_comm[1] == w.size:
        return _native_comm[0]
    import ctypes as C
    from . import native
    lib = native.load()
    uid = torch.zeros(128, dtype=torch.uint8)
    if w.rank == 0:
        buf = (C.c_ubyte * 128)()
        native.check(lib.bma_comm_unique_id(buf), "comm_unique_id")
        uid = torch.tensor(list(buf), dtype=torch.uint8)
    uid = uid.to(device)
    td.broadcast(uid, src=0)
    raw = bytes(uid.cpu().tolist())
    with torch.cuda.device(device):
        handle = lib.bma_comm_create(raw, w.size, w.rank)
    if not handle:
        raise RuntimeError("bimodal_b200: NCCL communicator creation failed: " +
                           lib.bma_comm_last_error().decode("utf-8", "replace"))
    _native_comm = (handle, w.size)
    return handle


def exchange_peer_handles(lib, scorer, device) -> None:
    """Collective: map every rank's prefix buffers + flag words into every other rank's scorer
    (cudaIpc handles, 192 bytes per rank, carried by torch.distributed) -- the NVLink peer-memory
    exchange of the split prefill (include/bimodal_b200.h, bma_scorer_export_ipc / import_ipc)."""
    import ctypes as C
    from . import native
    w = world()
    if w.size < 2 or w.size > 8:
        return
    buf = (C.c_ubyte * 192)()
    with torch.cuda.device(device):
        native.check(lib.bma_scorer_export_ipc(scorer, buf), "scorer_export_ipc")
    mine = torch.tensor(list(buf), dtype=torch.uint8, device=device)
    gathered = torch.empty(192 * w.size, dtype=torch.uint8, device=device)
    td.all_gather_into_tensor(gathered, mine)        # also orders "flags reset" before anybody's first signal
    raw = bytes(gathered.cpu().tolist())
    with torch.cuda.device(device):
        native.check(lib.bma_scorer_import_ipc(scorer, raw, w.size, w.rank), "scorer_import_ipc")


def reduce_winner(key: torch.Tensor, flag: Optional[torch.Tensor], w: World) -> None:
    """In-place MIN all-reduce of the int64 key (and MAX of the early-stop flag when given): ONE NCCL
    launch on the scorer's stream through the library's communicator; torch.distributed (gloo in
    the CPU tests) otherwise."""
    if w.size == 1:
        return
    if key.is_cuda and _native_comm is not None:
        from . import native
        with torch.cuda.device(key.device):
            native.check(native.load().bma_comm_reduce_winner(
                _native_comm[0], key.data_ptr(), None if flag is None else flag.data_ptr(),
                native.stream_ptr(key.device)), "comm_reduce_winner")
        return
    td.all_reduce(key, op=td.ReduceOp.MIN)
    if flag is not None:
        td.all_reduce(flag, op=td.ReduceOp.MAX)


def broadcast_candidates(sampled: Optional[torch.Tensor], n_tok: int, device, w: World) -> torch.Tensor:
    """Rank 0's candidate ids (their count varies per step after filter_ids) to every rank."""
    if w.size == 1:
        return sampled
    count = torch.tensor([sampled.shape[0] if w.rank == 0 else 0], dtype=torch.int64, device=device)
    td.broadcast(count, src=0)
    if w.rank != 0:
        sampled = torch.empty((int(count.item()), n_tok), dtype=torch.int64, device=device)
    td.broadcast(sampled, src=0)
    return sampled


def broadcast_tensor(t: torch.Tensor, w: World) -> torch.Tensor:
    if w.size > 1:
        td.broadcast(t, src=0)
    return t


def gather_losses(loss: torch.Tensor, lo: int, hi: int, w: World) -> None:
    """Optional: fill the non-local entries of ``loss`` (MIN over ranks; unscored slots are +inf)."""
    if w.size > 1:
        td.all_reduce(loss, op=td.ReduceOp.MIN)
