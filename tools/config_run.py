"""BASELINE configs 3 and 4 at full size: the packed-bit estimators (snpgdsIBS / snpgdsIBDKING
KING-robust) on synthetic data, one GPU or SNP-sharded over the ranks of a torchrun job.

    python tools/config_run.py --est ibs  --samples 50000 --snps 500000            # config 3, 1 GPU
    torchrun --nproc-per-node 4 tools/config_run.py --est ibs --samples 50000 --snps 500000
    torchrun --nproc-per-node 8 tools/config_run.py --est king --samples 100000 --snps 1000000   # config 4

Sharded mode: rank r owns a contiguous SNP block, accumulates its partial uint32 counters with no
communication and ONE NCCL all-reduce per row window sums them (exact integers: bit-identical for
any rank count).  When the N x N counters do not fit in HBM (config 4: 5 x 100k^2 x 4 B = 200 GB)
the matrix is walked in row windows; window w is finished (counters -> doubles, copy to pinned
host memory) by rank w mod world so the device-to-host traffic is spread over the ranks' links.
Timed region: first window's accumulate -> last window's result on the host (max over ranks).
Prints one JSON line on rank 0 with pair-SNPs/s, the algorithmic HBM bytes and an oracle check of
the first rows.  (Option names avoid prefixes of torchrun's own options: argparse abbreviation
matching would otherwise swallow them.)"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import snprelate_b200 as S
from snprelate_b200 import dist as D
from snprelate_b200._lib import EST_IBS, EST_KING_ROBUST

ap = argparse.ArgumentParser()
ap.add_argument("--est", default="ibs", choices=["ibs", "king"])
ap.add_argument("--samples", dest="n", type=int, default=50000)
ap.add_argument("--snps", dest="m", type=int, default=500000)
ap.add_argument("--rows", type=int, default=-1, help="row-window height (multiple of 256), 0 = whole matrix, -1 = auto")
ap.add_argument("--miss", type=float, default=0.005)
ap.add_argument("--check", type=int, default=16)
ap.add_argument("--reduce-to-finisher", action="store_true", help="NCCL reduce to the finishing rank instead of all-reduce")
ap.add_argument("--engine", default="bits", choices=["bits", "tensor"], help="pair-counter engine (snprel_set_count_engine)")
args = ap.parse_args()

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as td
    td.init_process_group("nccl", device_id=dev)

n, m = args.n, args.m
est = EST_IBS if args.est == "ibs" else EST_KING_ROBUST
bpp = 12 if args.est == "ibs" else 20          # counter bytes per pair (TIBS / TS_KINGRobust)
nout = 1 if args.est == "ibs" else 2           # float64 result matrices
lo, hi = D.shard_range(m, rank, world)
ctx = S.Context(local)
ctx.set_count_engine(args.engine)
ctx.geno_begin(n, hi - lo)
t0 = time.perf_counter()
ctx.geno_synth(hi - lo, seed=20261017, miss_rate=args.miss, snp_start=lo)
t_synth = time.perf_counter() - t0
npad = (n + 255) // 256 * 256
free, total = ctx.mem_info()
planes_bytes = (hi - lo + 127) // 128 * 128 // 4 * npad        # bit planes built by the first accumulate
budget = free - planes_bytes - (4 << 30)
rows = args.rows
if rows < 0:
    # counters (bpp) + the finish kernel's float64 output (8 * nout) per pair of a window row
    # (+ the int64 Gram planes of the tensor engine: 4 passes IBS, 6 KING-robust)
    per_row = (bpp + 8 * nout + (0 if args.engine == "bits" else 8 * (4 if args.est == "ibs" else 6))) * npad
    # (windows are capped at 8192 rows: the pinned host buffer of a window stays around 10 GB)
    rows = 0 if per_row * npad <= 0.8 * budget else max(256, min(8192, int(0.8 * budget / per_row) // 256 * 256))
wins = list(ctx.windows(rows)) if rows else [(0, 0)]

max_cnt = 0
for r0, h in wins:
    ctx.set_row_window(r0, h)
    max_cnt = max(max_cnt, ctx.window_count() if h else n * (n + 1) // 2)
finisher = any(w % world == rank for w in range(len(wins)))
host = [torch.empty(max_cnt, dtype=torch.float64, pin_memory=True).numpy() for _ in range(nout)] if finisher else None


def barrier():
    if world > 1:
        td.barrier()
    torch.cuda.synchronize()


barrier()
t_start = time.perf_counter()
hot_ms, first, t_acc, t_red, t_fin = 0.0, None, 0.0, 0.0, 0.0
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for w, (r0, h) in enumerate(wins):
    ctx.set_row_window(r0, h)
    ta = time.perf_counter()
    ctx.accumulate(est)
    hot_ms += ctx.last_hot_kernel()[0]
    tb = time.perf_counter()
    if world > 1:
        # only rank w mod world finishes window w: a reduce to that rank is enough
        D.allreduce_buffers(ctx.reduce_buffers(), device=dev, dst=(w % world) if args.reduce_to_finisher else None)
        torch.cuda.synchronize()
    ctx.mark_reduced()
    tc = time.perf_counter()
    if w % world == rank:
        if args.est == "ibs":
            out = (ctx.ibs_ave(packed=True, out=host[0]),)
        else:
            out = ctx.king_robust(None, packed=True, out=host)
        if r0 == 0 and rank == 0:
            first = [o[: 4 * n].copy() for o in out]
    td_ = time.perf_counter()
    t_acc += tb - ta
    t_red += tc - tb
    t_fin += td_ - tc
barrier()
t_job = time.perf_counter() - t_start
if world > 1:
    tt = torch.tensor([t_job, hot_ms], dtype=torch.float64, device=dev)
    td.all_reduce(tt, op=td.ReduceOp.MAX)
    t_job, hot_ms = float(tt[0]), float(tt[1])
ctx.set_row_window(0, 0)

check = None
if rank == 0 and first is not None and args.check > 0:
    from oracle import snprel_oracle as O        # checker only
    k = min(args.check, n)
    sub = O.synth_geno(k, m, seed=20261017, miss_rate=args.miss)        # samples 0..k-1, ALL SNPs (all shards)
    if args.est == "ibs":
        refs = [O.ibs_ave(O.ibs_counts(sub))]
    else:
        refs = list(O.king_robust(O.king_robust_counts(sub)))
    err = 0.0
    for got, ref in zip(first, refs):
        for i in range(min(k, 4)):
            base = i * n - i * (i - 1) // 2
            d = np.abs(got[base: base + (k - i)] - ref[i, i:k])
            err = max(err, float(np.nanmax(d)))
    check = {"rows": min(k, 4), "cols": k, "max_abs_err": err}

if rank == 0:
    pair_snps = 0.5 * n * n * m
    alg_bytes = n * m / 4 + bpp * n * (n + 1) / 2
    name = "snpgdsIBS (gnrIBSAve)" if args.est == "ibs" else "snpgdsIBDKING KING-robust"
    print(json.dumps({
        "workload": f"{name}, synthetic {n} samples x {m} SNPs, missing {args.miss}, {world} GPU(s), "
                    + (("SNP-block shards + one reduce (to the finishing rank) per row window" if args.reduce_to_finisher
                       else "SNP-block shards + one all-reduce per row window") if world > 1 else "single GPU")
                    + (f", {len(wins)} row windows of {rows}" if rows else ", whole matrix"),
        "n_gpus": world, "engine": args.engine, "job_s": round(t_job, 3), "pair_kernel_s_max_rank": round(hot_ms / 1e3, 3),
        "pair_snps_per_s": pair_snps / t_job, "pair_snps_per_s_kernel_only": pair_snps / (hot_ms / 1e3) ,
        "phases_s_rank0": {"accumulate": round(t_acc, 3), "allreduce": round(t_red, 3), "finish_d2h": round(t_fin, 3)},
        "algorithmic_hbm_bytes": alg_bytes, "achieved_hbm_gbs_algorithmic": alg_bytes / t_job / 1e9,
        "synth_s": round(t_synth, 2), "windows": len(wins), "window_rows": rows,
        "free_gib_after_load": round(free / 2**30, 1), "free_gib_end": round(ctx.mem_info()[0] / 2**30, 1),
        "oracle_check": check}))
if world > 1:
    td.destroy_process_group()
