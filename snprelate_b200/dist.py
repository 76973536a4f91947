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


_KIND = {0: ("<i8", np.int64), 1: ("<i4", np.int32), 2: ("<f8", np.float64)}   # uint32 sums == int32 sums mod 2^32


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


def reduce_plan(plan, group=None, device=None):
    """All ranks must use one fixed-point format: max-reduce max_abs, sum-reduce
    the other per-rank plan statistics."""
    import torch
    import torch.distributed as dist
    dev = "cpu" if device is None else device
    mx = torch.tensor([plan.max_abs, plan.max_abs_w], dtype=torch.float64, device=dev)
    # sums are conservative for the per-sample maxima (max of sums <= sum of maxima)
    sm = torch.tensor([plan.sum_bound, plan.err_weight, plan.scale, float(plan.total_missing),
                       float(plan.max_missing), float(plan.n_snp)], dtype=torch.float64, device=dev)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM, group=group)
    plan.max_abs = float(mx[0])
    plan.max_abs_w = float(mx[1])
    plan.sum_bound = float(sm[0])
    plan.err_weight = float(sm[1])
    plan.scale = float(sm[2])
    plan.total_missing = int(sm[3])
    plan.max_missing = int(sm[4])
    plan.n_snp = int(sm[5])
    return plan


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


def accumulate_sharded(ctx, est, bayesian=False, group=None, device=None):
    """Plan -> agree on the fixed-point format -> accumulate the local SNP shard ->
    all-reduce the partial accumulators -> mark reduced.  Afterwards the usual
    finish calls (ctx.grm / ctx.pca / ctx.ibs_num ...) read the global result."""
    import torch
    plan = ctx.plan_local(est, bayesian)
    plan = reduce_plan(plan, group, device)
    ctx.accumulate(est, plan)
    allreduce_buffers(ctx.reduce_buffers(), group, device)
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
