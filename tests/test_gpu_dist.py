"""Multi-GPU SNP sharding: needs >= 2 visible GPUs (skipped otherwise).  Two ranks
each load half of the SNPs; after the all-reduce every rank must hold exactly the
single-GPU result (integer planes are bit-identical, f64 scalars to 1e-15)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as tdist
    import snprelate_b200 as S
    from snprelate_b200 import dist as D
    from oracle import snprel_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    tdist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n, m = 500, 6000
    g = O.synth_geno(n, m, seed=99, miss_rate=0.01)
    lo, hi = D.shard_range(m, rank, world)
    ctx = S.Context(rank)
    out = {}
    for name, est in (("GCTA", 1), ("EIGMIX", 3), ("Eigenstrat", 0)):
        ctx.geno_begin(n, hi - lo)
        ctx.set_snp_origin(lo)
        ctx.geno_push_u8(g[lo:hi])
        D.accumulate_sharded(ctx, est, device=dev)
        out[name] = ctx.grm(name)[0]
    ctx.geno_begin(n, hi - lo)
    ctx.set_snp_origin(lo)
    ctx.geno_push_u8(g[lo:hi])
    ctx.accumulate(10)
    D.allreduce_buffers(ctx.reduce_buffers(), device=dev)
    torch.cuda.synchronize()
    ctx.mark_reduced()
    out["ibs"] = np.stack(ctx.ibs_num())
    ctx.geno_begin(n, hi - lo)
    ctx.set_snp_origin(lo)
    ctx.geno_push_u8(g[lo:hi])
    out["mom"] = D.ibd_mom_sharded(ctx, kinship_constraint=True, device=dev)[:2]
    # the library's own peer-memory reduction (CUDA IPC) instead of NCCL: all-reduce and reduce-to-root
    ctx.geno_begin(n, hi - lo)
    ctx.set_snp_origin(lo)
    ctx.geno_push_u8(g[lo:hi])
    D.accumulate_sharded(ctx, 1, device=dev, reduce="peer")
    out["gcta_peer"] = ctx.grm("GCTA")[0]
    ctx.accumulate(11)
    D.peer_reduce_buffers(ctx, rank, world, root=1, device=dev)
    ctx.mark_reduced()
    out["king_peer_root1"] = ctx.king_robust_counts() if rank == 1 else None
    # tiled mode: own SNP block from the host, the other block over NCCL straight into the device rows
    recv = D.load_sharded_then_gather(ctx, n, m, lambda a, b: ctx.geno_push_u8(g[a:b]), rank, world, device=dev)
    out["gathered"] = bool(np.array_equal(ctx.geno_copy_u8(), g)) and recv > 0
    out["gcta_after_gather"] = ctx.grm("GCTA")[0]
    q.put((rank, out))
    tdist.destroy_process_group()


def test_two_rank_snp_sharding_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from oracle import snprel_oracle as O
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    g = O.synth_geno(500, 6000, seed=99, miss_rate=0.01)
    ref = {"GCTA": O.grm_gcta(g), "EIGMIX": O.grm_eigmix(g), "Eigenstrat": O.grm_eigenstrat(g)}
    for rank in (0, 1):
        for k, r in ref.items():
            err = np.max(np.abs(res[rank][k] - r) / np.maximum(np.abs(r), 1))
            assert err < 1e-10, (rank, k, err)
        assert np.array_equal(res[rank]["ibs"], O.ibs_counts(g))
        assert res[rank]["gathered"]
        assert np.array_equal(res[rank]["gcta_peer"], res[rank]["GCTA"])          # same exact integers, same epilogue
    assert np.array_equal(res[1]["king_peer_root1"], O.king_robust_counts(g))
    for rank in (0, 1):
        assert np.max(np.abs(res[rank]["gcta_after_gather"] - ref["GCTA"]) / np.maximum(np.abs(ref["GCTA"]), 1)) < 1e-10
        e, _ = O.ibd_mom_tables(g)
        r0, r1 = O.ibd_mom(O.ibs_counts(g), e, True)
        assert np.max(np.abs(res[rank]["mom"][0] - r0)) < 1e-13 and np.max(np.abs(res[rank]["mom"][1] - r1)) < 1e-13
    for k in ref:
        assert np.array_equal(res[0][k], res[1][k])


def _tile_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import snprelate_b200 as S
    from oracle import snprel_oracle as O
    torch.cuda.set_device(rank)
    n, m = 1100, 4000
    g = O.synth_geno(n, m, seed=41, miss_rate=0.01)
    ctx = S.Context(rank)
    ctx.geno_begin(n, m)          # every rank holds ALL genotypes; the N x N output is dealt out
    ctx.geno_push_u8(g)
    parts = ctx.packed_by_windows(lambda: ctx.grm("GCTA", packed=True)[0], 256, rank=rank, world=world)
    ibs = ctx.packed_by_windows(lambda: ctx.ibs_ave(packed=True), 256, rank=rank, world=world)
    q.put((rank, parts, ibs))


def test_output_tiled_across_two_gpus():
    """N x N output tiled across GPUs (second mode): each rank computes
    every second 256-row window of the packed triangle, no collective; the union of the slices
    is the single-GPU result."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from oracle import snprel_oracle as O
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    procs = [ctxm.Process(target=_tile_worker, args=(r, 2, 0, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    n = 1100
    g = O.synth_geno(n, 4000, seed=41, miss_rate=0.01)
    grm = np.full(n * (n + 1) // 2, np.nan)
    ibs = np.full(n * (n + 1) // 2, np.nan)
    for rank, parts, ib in res:
        for off, sl in parts:
            grm[off:off + len(sl)] = sl
        for off, sl in ib:
            ibs[off:off + len(sl)] = sl
    ref = O.to_packed_upper(O.grm_gcta(g))
    assert np.max(np.abs(grm - ref) / np.maximum(np.abs(ref), 1)) < 1e-10
    assert np.array_equal(ibs, O.to_packed_upper(O.ibs_ave(O.ibs_counts(g))))
