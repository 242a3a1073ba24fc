"""world_size-2 (gloo, CPU) test of the multi-GPU host logic: shard bounds, the packed-key MIN
all-reduce that picks the winner, candidate broadcast from rank 0."""
import os
import socket

import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, n_cand, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from bimodalattack_b200 import dist as bd
        from bimodalattack_b200.engine import unpack_key
        w = bd.world()
        assert (w.rank, w.size) == (rank, world_size)
        g = torch.Generator().manual_seed(4)
        losses = torch.rand(n_cand, generator=g) * 5          # identical on both ranks
        losses[7] = losses[3] = losses.min() - 0.5            # an exact tie: lowest index must win
        lo, hi = bd.shard_bounds(n_cand, w)
        key = torch.tensor([min([bd.pack_key(float(losses[i]), i) for i in range(lo, hi)], default=2 ** 63 - 1)],
                           dtype=torch.int64)
        flag = torch.tensor([1 if rank == 1 else 0], dtype=torch.int32)
        bd.reduce_winner(key, flag, w)
        loss, idx = unpack_key(int(key.item()))
        assert idx == int(losses.argmin()) == 3 and loss == float(losses[3])
        assert int(flag.item()) == 1
        # candidate broadcast: only rank 0 holds the sampled ids, their count varies per step
        sampled = torch.arange(5 * 6, dtype=torch.int64).view(5, 6) if rank == 0 else None
        got = bd.broadcast_candidates(sampled, 6, torch.device("cpu"), w)
        assert got.shape == (5, 6) and int(got[4, 5]) == 29
        # partial loss vectors merge by MIN
        part = torch.full((n_cand,), float("inf"))
        part[lo:hi] = losses[lo:hi]
        bd.gather_losses(part, lo, hi, w)
        assert torch.equal(part, losses)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        td.destroy_process_group()


def test_shard_bounds_cover_everything():
    from bimodalattack_b200.dist import World, shard_bounds
    for n in (0, 1, 7, 64, 511, 512, 4096):
        for size in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, World(r, size)) for r in range(size)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


@pytest.mark.timeout(180)
def test_two_rank_winner_reduce(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 37, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_balanced_cuts_cover_everything_and_even_out_the_cost():
    """Row-balanced sharding (dist.balanced_cuts): contiguous, complete, monotone, and the most expensive
    shard is close to total / size (within one candidate's cost) -- versus equal counts, which can be far off."""
    import torch
    from bimodalattack_b200.dist import World, balanced_cuts, shard_bounds
    g = torch.Generator().manual_seed(0)
    for n in (0, 1, 5, 64, 512, 513):
        for size in (1, 2, 3, 8):
            cost = torch.randint(1, 42, (n,), generator=g)
            cuts = balanced_cuts(cost, size)
            assert len(cuts) == size + 1 and cuts[0] == 0 and cuts[-1] == n
            assert all(a <= b for a, b in zip(cuts, cuts[1:]))
            if n >= 2 * size:
                worst = max(int(cost[a:b].sum()) for a, b in zip(cuts, cuts[1:]))
                assert worst <= int(cost.sum()) / size + int(cost.max())
    # a skewed batch: the first half of the candidates costs 4x the second half
    cost = torch.cat([torch.full((256,), 40), torch.full((256,), 10)])
    cuts = balanced_cuts(cost, 8)
    bal = max(int(cost[a:b].sum()) for a, b in zip(cuts, cuts[1:]))
    even = max(int(cost[slice(*shard_bounds(512, World(r, 8)))].sum()) for r in range(8))
    assert bal < 0.7 * even
    assert balanced_cuts(torch.ones(7, dtype=torch.int64), 1) == [0, 7]


def test_weighted_cuts_and_share_update():
    """dist.balanced_cuts with per-rank shares, and the straggler-aware share update: a rank that waits (it is
    faster) gains work, the iteration equalises the ranks' busy times, shares stay normalised and clamped."""
    import torch
    from bimodalattack_b200 import dist as d
    cost = torch.randint(1, 40, (512,), generator=torch.Generator().manual_seed(0))
    assert d.balanced_cuts(cost, 4, [0.25] * 4) == d.balanced_cuts(cost, 4)
    cuts = d.balanced_cuts(cost, 4, [0.1, 0.3, 0.3, 0.3])
    tot = int(cost.sum())
    part = [int(cost[a:b].sum()) for a, b in zip(cuts, cuts[1:])]
    assert cuts[0] == 0 and cuts[-1] == 512 and cuts == sorted(cuts)
    assert abs(part[0] / tot - 0.1) < 0.02 and all(abs(x / tot - 0.3) < 0.03 for x in part[1:])
    share, speed = [0.5, 0.5], [0.95, 1.0]
    for _ in range(12):
        busy = [share[i] / speed[i] * 200.0 for i in range(2)]
        t = max(busy)
        share = d.updated_shares(share, [t, t], [t - b for b in busy])
        assert abs(sum(share) - 1.0) < 1e-9
    assert abs(share[0] / share[1] - 0.95) < 2e-3            # busy times equal <=> shares proportional to speeds
    clamped = d.updated_shares([0.5, 0.5], [100.0, 100.0], [0.0, 90.0], damping=1.0)
    assert max(clamped) <= 1.25 / 2 + 1e-9 and min(clamped) >= 1 / (1.25 * 2) - 0.02      # (renormalised after clamping)
