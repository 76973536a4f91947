"""Multi-GPU plumbing: SNP-block sharding + one sum-reduction of the partial
accumulators.

One process per GPU (torch.distributed, NCCL over NVLink; gloo on CPU for the
host-logic tests).  Every accumulator of this library is a sum over SNPs of
per-SNP terms that only need that SNP's full sample column, so rank r owns a
contiguous SNP range, accumulates its partial N x N planes with no communication,
and a single all-reduce finishes the job -- the identity the reference exploits in
snpgdsMergeGRM (src/genPCA.cpp:1834-1855).  The partials are exact integers
(uint32 counters, int64 fixed point), so the reduced result is bit-identical for
any GPU count; only the handful of float64 scalars (SumDenominator) are
order dependent at the 1e-16 level.  torch is plumbing here (device tensors over
the library's buffers, the collective); all arithmetic stays in libsnprel_b200.so.
"""
from __future__ import annotations

import numpy as np

SNP_ALIGN = 128     # shard boundaries are multiples of the kernels' SNP stage


def shard_range(n_snp: int, rank: int, world: int):
    """Contiguous [start, stop) SNP range of `rank`; interior boundaries are
    multiples of SNP_ALIGN and the shards tile [0, n_snp) exactly."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    blocks = (n_snp + SNP_ALIGN - 1) // SNP_ALIGN
    lo = (blocks * rank) // world * SNP_ALIGN
    hi = (blocks * (rank + 1)) // world * SNP_ALIGN
    return min(lo, n_snp), min(hi, n_snp)


class _DevArray:
    """Expose a raw device pointer through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


_KIND = {0: ("<i8", np.int64), 1: ("<i4", np.int32), 2: ("<f8", np.float64),   # uint32 sums == int32 sums mod 2^32
         3: ("|u1", np.uint8)}                                                  # raw bytes (packed genotype rows)


def buffer_tensor(ptr, count, kind, device):
    """torch tensor aliasing one of the library's reduce buffers.  `device` None
    means the pointer is host memory (CPU tests)."""
    import torch
    typestr, npdt = _KIND[kind]
    if device is None:
        import ctypes
        buf = (ctypes.c_char * (count * np.dtype(npdt).itemsize)).from_address(ptr)
        return torch.from_numpy(np.frombuffer(buf, dtype=npdt))
    return torch.as_tensor(_DevArray(ptr, count, typestr), device=device)


_PLAN_MAX = ("max_abs", "max_abs_w")
# sums are conservative for the per-sample maxima (max of sums <= sum of maxima)
_PLAN_SUM = ("sum_bound", "err_weight", "scale", "total_missing", "max_missing", "n_snp", "diag_bound", "sum_rest",
             "err_weight2")
_PLAN_INT = ("total_missing", "max_missing", "n_snp")


def _plan_vectors(plan):
    return ([float(getattr(plan, k)) for k in _PLAN_MAX], [float(getattr(plan, k)) for k in _PLAN_SUM])


def _plan_store(plan, mx, sm):
    for k, v in zip(_PLAN_MAX, mx):
        setattr(plan, k, float(v))
    for k, v in zip(_PLAN_SUM, sm):
        setattr(plan, k, int(v) if k in _PLAN_INT else float(v))
    return plan


def reduce_plan(plan, group=None, device=None):
    """All ranks must use one fixed-point format: max-reduce max_abs, sum-reduce
    the other per-rank plan statistics."""
    import torch
    import torch.distributed as dist
    dev = "cpu" if device is None else device
    mxl, sml = _plan_vectors(plan)
    mx = torch.tensor(mxl, dtype=torch.float64, device=dev)
    sm = torch.tensor(sml, dtype=torch.float64, device=dev)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM, group=group)
    return _plan_store(plan, mx.tolist(), sm.tolist())


def merge_plans(plans):
    """The same merge as reduce_plan for plans held by ONE process (several contexts, e.g. one per
    GPU of the box driven from a single host thread): every plan receives the merged statistics."""
    vec = [_plan_vectors(p) for p in plans]
    mx = [max(v[0][k] for v in vec) for k in range(len(_PLAN_MAX))]
    sm = [sum(v[1][k] for v in vec) for k in range(len(_PLAN_SUM))]
    for p in plans:
        _plan_store(p, mx, sm)
    return plans


def accumulate_in_process(contexts, est, bayesian=False):
    """SNP-sharded accumulation over contexts of ONE process (each holding its own SNP range, on
    the same or on different GPUs): plan -> merged format -> local accumulate -> the partial
    buffers are summed into every context (device-to-device adds through torch views of the
    library's buffers) -> mark reduced.  The algebra is exactly accumulate_sharded's, without a
    process group; afterwards any context's finish calls return the global result."""
    import torch
    plans = merge_plans([c.plan_local(est, bayesian) for c in contexts])
    for c, p in zip(contexts, plans):
        c.accumulate(est, p)
    bufs = [c.reduce_buffers() for c in contexts]
    devs = [torch.device("cuda", c.device) for c in contexts]
    for k in range(len(bufs[0])):
        ts = [buffer_tensor(b[k][0], b[k][1], b[k][2], d) for b, d in zip(bufs, devs)]
        if ts[0].numel() == 0:
            continue
        total = ts[0].clone()
        for t in ts[1:]:
            total += t.to(devs[0])
        for t, d in zip(ts, devs):
            t.copy_(total.to(d))
    for d in set(devs):
        torch.cuda.synchronize(d)
    for c in contexts:
        c.mark_reduced()
    return plans[0]


def allreduce_buffers(buffers, group=None, device=None, dst=None):
    """Sum-reduce every (ptr, count, kind) buffer in place across the group.  With `dst` only
    that rank receives the sums (a reduce instead of an all-reduce: half the link traffic when a
    single rank finishes the window, e.g. row windows dealt to ranks for the epilogue)."""
    import torch.distributed as dist
    for ptr, count, kind in buffers:
        if count <= 0:
            continue
        t = buffer_tensor(ptr, count, kind, device)
        if dst is None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)


def peer_reduce_buffers(ctx, rank, world, root=None, group=None, device=None):
    """The library's own reduction over NVLink peer memory for one process per GPU (csrc/multi.cu):
    every rank maps its peers' reduce buffers through CUDA IPC (handles exchanged with one
    all_gather, re-used while the buffers keep their allocations), sums ITS row slice of every
    buffer out of the peers' HBM, and after a barrier pulls the other ranks' reduced slices -- every
    rank (`root` None: an all-reduce) or only the finishing rank (`root`).  Only the live
    upper-triangle columns of the N x N planes cross the links.  NCCL carries the 64-byte handles
    and the barriers only.  Returns the bytes this rank moved over the links."""
    import torch
    import torch.distributed as dist
    dev = "cpu" if device is None else device
    # The handles are exchanged only when some rank's buffers changed (they are persistent allocations,
    # so after the first window this is one tiny all-reduce -- which is also the barrier that tells every
    # rank that its peers have finished accumulating).
    key = tuple((int(p), int(c), int(k)) for p, c, k in ctx.reduce_buffers())
    cache = getattr(ctx, "_peer_cache", None)
    changed = torch.tensor([0 if (cache is not None and cache == (key, rank, world)) else 1], dtype=torch.int32, device=dev)
    dist.all_reduce(changed, op=dist.ReduceOp.MAX, group=group)
    if int(changed.item()):
        hb, off = ctx.reduce_ipc_handles()
        nb = off.size
        th = torch.from_numpy(hb).to(dev)
        to = torch.from_numpy(off).to(dev)
        gh = torch.empty(world * nb * 64, dtype=torch.uint8, device=dev)
        go = torch.empty(world * nb, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gh, th, group=group)
        dist.all_gather_into_tensor(go, to, group=group)
        ctx.peer_reduce_open(rank, world, gh.cpu().numpy(), go.cpu().numpy())
        ctx._peer_cache = (key, rank, world)
    moved = ctx.peer_reduce_phase(1, -1 if root is None else root)
    dist.barrier(group=group)
    moved += ctx.peer_reduce_phase(2, -1 if root is None else root)
    dist.barrier(group=group)                             # nobody re-uses its buffers while a peer still reads them
    return moved


def accumulate_sharded(ctx, est, bayesian=False, group=None, device=None, reduce="nccl", root=None):
    """Plan -> agree on the fixed-point format -> accumulate the local SNP shard ->
    sum the partial accumulators over the ranks -> mark reduced.  Afterwards the usual
    finish calls (ctx.grm / ctx.pca / ctx.ibs_num ...) read the global result.
    reduce: "nccl" (all-reduce / reduce to `root`) or "peer" (the library's peer-memory reduction,
    GPUs of one box only)."""
    import torch
    import torch.distributed as dist
    plan = ctx.plan_local(est, bayesian)
    plan = reduce_plan(plan, group, device)
    ctx.accumulate(est, plan)
    if reduce == "peer":
        peer_reduce_buffers(ctx, dist.get_rank(group), dist.get_world_size(group), root, group, device)
    else:
        allreduce_buffers(ctx.reduce_buffers(), group, device, dst=root)
    if device is not None:
        torch.cuda.synchronize(device)
    ctx.mark_reduced()
    return plan


def ibd_mom_sharded(ctx, allele_freq=None, kinship_constraint=False, packed=False, group=None, device=None):
    """PLINK method of moments over SNP-sharded ranks: the IBS counters and the six
    per-SNP expectation sums (src/genIBD.cpp:253-338) are both plain sums over SNPs.
    `allele_freq` is this rank's slice.  Returns (k0, k1, afreq of the local shard)."""
    import torch
    import torch.distributed as dist
    from ._lib import EST_IBS
    accumulate_sharded(ctx, EST_IBS, group=group, device=device)
    sums, afreq = ctx.ibd_mom_sums(allele_freq)
    t = torch.from_numpy(sums).to("cpu" if device is None else device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    k0, k1 = ctx.ibd_mom_from_sums(t.cpu().numpy(), kinship_constraint, packed)
    return k0, k1, afreq


def load_sharded_then_gather(ctx, n_samp, n_snp, fill, rank, world, group=None, device=None):
    """Tiled N x N output ("when N^2 exceeds one GPU"): every rank finally holds
    ALL SNPs, but only its own SNP block crosses its PCIe link -- `fill(lo, hi)` pushes SNPs [lo, hi)
    into `ctx` (geno_push_2b / geno_push_u8 from the host, or the synthetic generator) -- and the other
    blocks arrive over NVLink: one NCCL broadcast per block straight into the device rows of the
    workspace (an all-gather with unequal block sizes).  Returns the bytes received from peers."""
    import torch.distributed as dist
    ctx.geno_begin(n_samp, n_snp)
    lo, hi = shard_range(n_snp, rank, world)
    ctx.geno_seek(lo)
    if hi > lo:
        fill(lo, hi)
    ptr, row_bytes, _ = ctx.geno_device_rows()
    received = 0
    for s in range(world):
        slo, shi = shard_range(n_snp, s, world)
        if shi <= slo:
            continue
        t = buffer_tensor(ptr + slo * row_bytes, (shi - slo) * row_bytes, 3, device)
        dist.broadcast(t, src=s if group is None else dist.get_global_rank(group, s), group=group)
        if s != rank:
            received += (shi - slo) * row_bytes
    if device is not None:
        import torch
        torch.cuda.synchronize(device)
    ctx.geno_commit(n_snp)
    return received
