"""Host-side logic of the multi-GPU path on CPU: SNP shard ranges and the
plan / buffer reductions, exercised with world_size 2 over gloo."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from snprelate_b200 import dist as D
from snprelate_b200._lib import Plan


def test_shard_ranges_tile_exactly():
    for n_snp in (0, 1, 127, 128, 129, 1000, 65536, 1000003):
        for world in (1, 2, 3, 8):
            ranges = [D.shard_range(n_snp, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n_snp
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0 and a0 <= a1
            for lo, hi in ranges[:-1]:
                assert hi % D.SNP_ALIGN == 0 or hi == n_snp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = Plan()
    plan.max_abs = 10.0 + rank
    plan.sum_bound = 100.0 * (rank + 1)
    plan.max_missing = 5 + rank
    plan.total_missing = 50 + rank
    plan.err_weight = 7.0
    plan.err_weight2 = 40.0 + rank
    plan.scale = 2.5
    plan.n_snp = 1000 * (rank + 1)
    plan = D.reduce_plan(plan, device=None)
    # every rank derives the SAME fixed-point format (digits, fractional bits, rounding rule) from the merged
    # statistics, here those of two half shards of config 2 (host-only format choice of the library)
    half = Plan()
    half.frac_bits = half.frac_bits_w = half.frac_bits_d = -1
    for k, v in dict(max_abs=0.80 + 0.007 * rank, max_abs_w=0.085, err_weight=6.3e6 + rank, err_weight2=1.6e8, scale=1.0e6,
                     sum_bound=8.2e6, diag_bound=1.2e6, sum_rest=1.4e4, total_missing=25000000 + rank,
                     max_missing=2700 + rank, n_snp=500000).items():
        setattr(half, k, v)
    half = D.reduce_plan(half, device=None)
    from snprelate_b200._lib import plan_format
    npass = plan_format(0, half, "auto", 10000)
    fmt = (npass, half.digits, half.digits_w, half.frac_bits, half.frac_bits_w, half.rounding, half.n_snp, half.err_weight2)
    # three buffers of the three kinds the library exposes, as raw host pointers
    a = np.arange(6, dtype=np.int64) * (rank + 1) - 3
    b = (np.arange(4, dtype=np.uint32) + 4294967290 + rank).astype(np.uint32)   # wraps mod 2^32
    c = np.array([0.5, 1.25]) * (rank + 1)
    bufs = [(a.ctypes.data, a.size, 0), (b.ctypes.data, b.size, 1), (c.ctypes.data, c.size, 2)]
    D.allreduce_buffers(bufs, device=None)
    # reduce-to-one-rank flavour (row windows finished by a single rank)
    d = np.arange(5, dtype=np.int64) + 10 * rank
    D.allreduce_buffers([(d.ctypes.data, d.size, 0)], device=None, dst=1)
    if rank == 1:
        assert d.tolist() == (2 * np.arange(5) + 10).tolist()
    q.put((rank, plan.max_abs, plan.sum_bound, plan.max_missing, plan.n_snp, plan.total_missing, plan.err_weight, plan.scale, a.tolist(), b.tolist(), c.tolist(),
           plan.err_weight2, fmt))
    dist.destroy_process_group()


def test_plan_and_buffer_reduction_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp_a = ((np.arange(6) * 1 - 3) + (np.arange(6) * 2 - 3)).tolist()
    exp_b = ((np.arange(4, dtype=np.uint64) + 4294967290) + (np.arange(4, dtype=np.uint64) + 4294967291)) % (1 << 32)
    assert res[0][-1] == res[1][-1]                       # one format on all ranks
    assert res[0][-1][:3] == (7, 4, 3) and res[0][-1][5] == 1 and res[0][-1][6] == 1000000 and res[0][-1][7] == 3.2e8
    for rank, mx, sb, mm, ns, tmiss, ew, sc, a, b, c, ew2, fmt in res:
        assert mx == 11.0 and sb == 300.0 and mm == 11 and ns == 3000
        assert tmiss == 101 and ew == 14.0 and sc == 5.0 and ew2 == 81.0
        assert a == exp_a
        assert b == exp_b.astype(np.uint32).tolist()
        assert c == [1.5, 3.75]


def test_window_owner_deals_balanced_shares():
    """Row windows of the upper triangle shrink with their index; the boustrophedon deal keeps
    every rank's pair count within 0.2 % of 1/world at config-5 geometry (round-robin: 2.9 %)."""
    from snprelate_b200._lib import window_owner
    n, rows = 500000, 2048
    starts = list(range(0, n, rows))
    for world in (1, 2, 3, 8):
        share = np.zeros(world)
        seen = []
        for k, r0 in enumerate(starts):
            h = min(rows, n - r0)
            owner = window_owner(k, world)
            assert 0 <= owner < world
            seen.append(owner)
            share[owner] += h * (n - r0) - h * (h - 1) / 2        # pairs (i <= j) in the band
        assert sorted(set(seen)) == list(range(world))
        assert share.sum() == n * (n + 1) / 2
        assert np.max(np.abs(share / share.sum() * world - 1)) < 2e-3
