"""Sharding helpers and the single final collective, world_size 2 over gloo on CPU."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tpl_b200 import dist as tdist


def test_shard_ranges_cover_everything():
    for total, world in ((4096, 8), (10, 3), (7, 8), (65536, 8)):
        spans = [tdist.shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    (s_lo, s_hi), (p_lo, p_hi) = tdist.shard_scenes(64, 1024, 3, 8)      # config #3: 8 scenes per GPU
    assert (s_lo, s_hi) == (24, 32) and (p_lo, p_hi) == (24 * 1024, 32 * 1024)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        scenes, per = 6, 5                                    # 3 scenes per rank
        costs = torch.rand(scenes * per, dtype=torch.float64)
        costs[7] = float("nan")
        (s_lo, s_hi), (p_lo, p_hi) = tdist.shard_scenes(scenes, per, rank, world)
        local = costs[p_lo:p_hi]
        # per-scene best on this rank (what tplb_argmin_groups computes on the GPU)
        safe = torch.where(torch.isfinite(local), local, torch.full_like(local, float("inf")))
        mn, am = safe.view(-1, per).min(dim=1)
        am = (am + torch.arange(s_hi - s_lo) * per).to(torch.int32)
        gmin, garg = tdist.gather_best(mn, am, p_lo)
        gathered = tdist.gather_costs(local)
        ok = torch.equal(torch.nan_to_num(gathered, nan=-1.0), torch.nan_to_num(costs, nan=-1.0))
        want = torch.where(torch.isfinite(costs), costs, torch.full_like(costs, float("inf"))).view(scenes, per)
        wmin, warg = want.min(dim=1)
        ok = ok and torch.equal(gmin, wmin) and torch.equal(garg, warg + torch.arange(scenes) * per)
        best, where = tdist.global_best(gathered)
        ok = ok and float(best) == float(wmin.min()) and int(where) == int(garg[wmin.argmin()])
        # uneven shards
        lo, hi = tdist.shard_range(7, rank, world)
        g2 = tdist.gather_costs(costs[lo:hi], counts=[4, 3])
        ok = ok and torch.equal(torch.nan_to_num(g2, nan=-1.0), torch.nan_to_num(costs[:7], nan=-1.0))
        # uneven scene shards: 5 scenes on 2 ranks (3 + 2); the per-scene best still arrives in scene order
        scenes2 = 5
        counts = [hi - lo for lo, hi in (tdist.shard_range(scenes2, r, world) for r in range(world))]
        (s_lo, s_hi), (p_lo, p_hi) = tdist.shard_scenes(scenes2, per, rank, world)
        local = costs[p_lo:p_hi]
        safe = torch.where(torch.isfinite(local), local, torch.full_like(local, float("inf")))
        mn, am = safe.view(-1, per).min(dim=1)
        am = (am + torch.arange(s_hi - s_lo) * per).to(torch.int32)
        gmin, garg = tdist.gather_best(mn, am, p_lo, counts=counts)
        ok = ok and counts == [3, 2] and torch.equal(gmin, wmin[:scenes2])
        ok = ok and torch.equal(garg, (warg + torch.arange(scenes) * per)[:scenes2])
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_final_gather_two_ranks_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_single_process_paths():
    c = torch.tensor([3.0, 1.0, float("inf"), 2.0], dtype=torch.float64)
    assert torch.equal(tdist.gather_costs(c), c)
    mn, arg = tdist.gather_best(torch.tensor([1.0]), torch.tensor([1], dtype=torch.int32), 100)
    assert float(mn) == 1.0 and int(arg) == 101
    assert int(tdist.global_best(c)[1]) == 1
