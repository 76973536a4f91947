"""BASELINE configs 3 and 4 as library calls: the packed-bit estimators (snpgdsIBS / snpgdsIBDKING
KING-robust) on synthetic data at full size, one GPU or SNP-sharded over the ranks of a torchrun job.
Used by tools/config_run.py (CLI) and by bench.py's optional legs (--gpus 4: config 3, --gpus 8: config 4).

Sharded mode: rank r owns a contiguous SNP block, accumulates its partial uint32 counters with no
communication, and ONE sum-reduction per row window finishes it (exact integers: bit-identical for any
rank count) -- by default the library's own peer-memory reduction to the finishing rank over NVLink
(dist.peer_reduce_buffers: upper-triangle columns only), or an NCCL all-reduce / reduce.  When the
N x N counters do not fit in HBM (config 4: 5 x 100k^2 x 4 B = 200 GB) the matrix is walked in row
windows; window w is finished (counters -> doubles, copy to pinned host memory) by rank w mod world so
the device-to-host traffic is spread over the ranks' links.  Timed region: first window's accumulate ->
last window's result on the host (max over ranks)."""
from __future__ import annotations

import time

import numpy as np

from . import dist as D
from ._lib import Context, EST_IBS, EST_KING_ROBUST

SEED = 20261017


def run_pair_config(local, est_name, n, m, rank=0, world=1, engine="bits", reduce="peer", rows=-1, miss=0.005,
                    check_idx=None, reduce_to_finisher=True):
    """Returns (summary dict, kept rows).  `check_idx`: sorted sample indices whose entries the caller
    wants to verify: kept rows = [(a, [values (check_idx[a], check_idx[a:]) of every result matrix])] for
    the rows that fell into windows THIS rank finished (the checker lives outside the product)."""
    import torch
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as td
    est = EST_IBS if est_name == "ibs" else EST_KING_ROBUST
    bpp = 12 if est_name == "ibs" else 20          # counter bytes per pair (TIBS / TS_KINGRobust)
    nout = 1 if est_name == "ibs" else 2           # float64 result matrices
    lo, hi = D.shard_range(m, rank, world)
    ctx = Context(local)
    try:
        ctx.set_count_engine(engine)
        ctx.set_async_output(True)        # the D2H of a finished window overlaps the next window's accumulate
        ctx.geno_begin(n, hi - lo)
        t0 = time.perf_counter()
        ctx.geno_synth(hi - lo, seed=SEED, miss_rate=miss, snp_start=lo)
        t_synth = time.perf_counter() - t0
        npad = (n + 255) // 256 * 256
        free, _total = ctx.mem_info()
        planes_bytes = (hi - lo + 127) // 128 * 128 // 4 * npad        # bit planes built by the first accumulate
        budget = free - planes_bytes - (4 << 30)
        if rows < 0:
            # counters (bpp) + the finish kernel's float64 output (8 * nout) per pair of a window row
            # (+ the int64 Gram planes of the tensor engine: 4 passes IBS, 6 KING-robust)
            per_row = (bpp + 8 * nout + (0 if engine == "bits" else 8 * (4 if est_name == "ibs" else 6))) * npad
            # (windows are capped at 8192 rows: the pinned host buffer of a window stays around 10 GB)
            rows = 0 if per_row * npad <= 0.8 * budget else max(256, min(8192, int(0.8 * budget / per_row) // 256 * 256))
        wins = list(ctx.windows(rows)) if rows else [(0, 0)]
        max_cnt = 0
        for r0, h in wins:
            ctx.set_row_window(r0, h)
            max_cnt = max(max_cnt, ctx.window_count() if h else n * (n + 1) // 2)
        finisher = any(w % world == rank for w in range(len(wins)))
        host = [torch.empty(max_cnt, dtype=torch.float64, pin_memory=True).numpy() for _ in range(nout)] if finisher else None

        idx = np.zeros(0, dtype=np.int64) if check_idx is None else np.asarray(check_idx, dtype=np.int64)
        got_rows = []
        wanted = []                        # rows of the window whose copy is still in flight: (a, offset in the slice)

        def collect(out):
            ctx.output_wait()
            for a, base in wanted:
                got_rows.append((a, [o[base + idx[a:]].copy() for o in out]))
            wanted.clear()
        last_out = None

        def barrier():
            if world > 1:
                td.barrier()
            torch.cuda.synchronize()

        barrier()
        t_start = time.perf_counter()
        hot_ms, t_acc, t_red, t_fin, link_bytes = 0.0, 0.0, 0.0, 0.0, 0
        for w, (r0, h) in enumerate(wins):
            ctx.set_row_window(r0, h)
            ta = time.perf_counter()
            ctx.accumulate(est)
            hot_ms += ctx.last_hot_kernel()[0]
            tb = time.perf_counter()
            root = w % world
            if world > 1:
                if reduce == "peer":
                    link_bytes += D.peer_reduce_buffers(ctx, rank, world, root=root if reduce_to_finisher else None, device=dev)
                else:
                    D.allreduce_buffers(ctx.reduce_buffers(), device=dev, dst=root if reduce_to_finisher else None)
                    torch.cuda.synchronize()
            ctx.mark_reduced()
            tc = time.perf_counter()
            if root == rank:
                if last_out is not None:
                    collect(last_out)      # the previous window of this rank has long arrived; its host buffer is reused now
                out = (ctx.ibs_ave(packed=True, out=host[0]),) if est_name == "ibs" else ctx.king_robust(None, packed=True, out=host)
                last_out = out
                r1 = min(r0 + h, n) if h else n
                pbase = r0 * (2 * n - r0 - 1) // 2 + r0
                for a, i in enumerate(idx):
                    if r0 <= i < r1:
                        wanted.append((a, int(i) * (2 * n - int(i) - 1) // 2 - pbase))
            td_ = time.perf_counter()
            t_acc += tb - ta
            t_red += tc - tb
            t_fin += td_ - tc
        if last_out is not None:
            collect(last_out)              # (inside the timed region: the last result must be on the host)
        barrier()
        t_job = time.perf_counter() - t_start
        if world > 1:
            tt = torch.tensor([t_job, hot_ms], dtype=torch.float64, device=dev)
            td.all_reduce(tt, op=td.ReduceOp.MAX)
            t_job, hot_ms = float(tt[0]), float(tt[1])
        ctx.set_row_window(0, 0)

        pair_snps = 0.5 * n * n * m
        alg_bytes = n * m / 4 + bpp * n * (n + 1) / 2
        name = "snpgdsIBS (gnrIBSAve)" if est_name == "ibs" else "snpgdsIBDKING KING-robust"
        how = "single GPU"
        if world > 1:
            how = ("SNP-block shards + one " + ("peer-memory reduction (library kernel over NVLink, upper triangle)" if reduce == "peer" else "NCCL reduction")
                   + (" to the finishing rank" if reduce_to_finisher else " to every rank") + " per row window")
        return {
            "workload": f"{name}, synthetic {n} samples x {m} SNPs, missing {miss}, {world} GPU(s), {how}"
                        + (f", {len(wins)} row windows of {rows}" if rows else ", whole matrix"),
            "n_gpus": world, "engine": engine, "job_s": round(t_job, 3), "pair_kernel_s_max_rank": round(hot_ms / 1e3, 3),
            "pair_snps_per_s": pair_snps / t_job, "pair_snps_per_s_kernel_only": pair_snps / (hot_ms / 1e3),
            "phases_s_this_rank": {"accumulate": round(t_acc, 3), "reduce": round(t_red, 3), "finish_d2h": round(t_fin, 3)},
            "link_bytes_this_rank": int(link_bytes),
            "algorithmic_hbm_bytes": alg_bytes, "achieved_hbm_gbs_algorithmic": alg_bytes / t_job / 1e9,
            "synth_s": round(t_synth, 2), "windows": len(wins), "window_rows": rows}, got_rows
    finally:
        ctx.close()
