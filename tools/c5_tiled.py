"""BASELINE config 5 (snpgdsGRM GCTA, 500k samples x 800k SNPs, N x N output tiled across GPUs)
or any other tiled run: every rank holds the whole 2-bit matrix (100 GB at C5), the row windows of
the N x N output are dealt to the ranks in boustrophedon order, no collective.  On one GPU `--world 8 --rank r`
runs exactly the share rank r of an 8-GPU job would run (same windows, same time), so the
8-GPU time can be measured at 1/8 of the GPU-minutes; under torchrun rank / world come from the
environment and the job time is the max over ranks.

    python tools/c5_tiled.py [--samples 500000] [--snps 800000] [--rows 2048] [--world 8] [--rank 0]
                             [--method GCTA] [--max-windows K]
Prints one JSON line (rank 0).  Each window's packed slice is copied to a reused pinned host
buffer (the D2H is inside the timed region); nothing is kept -- the full C5 output is 1 TB.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import snprelate_b200 as S
from snprelate_b200._lib import window_owner

ap = argparse.ArgumentParser()
ap.add_argument("--samples", dest="n", type=int, default=500000)
ap.add_argument("--snps", dest="m", type=int, default=800000)
ap.add_argument("--rows", type=int, default=2048)
ap.add_argument("--world", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
ap.add_argument("--rank", type=int, default=int(os.environ.get("RANK", "0")))
ap.add_argument("--method", default="GCTA")
ap.add_argument("--miss", type=float, default=0.005)
ap.add_argument("--max-windows", type=int, default=0, help="stop after K of this rank's windows (0 = all)")
ap.add_argument("--check", type=int, default=16, help="oracle check on the first K samples of window 0 (rank 0)")
args = ap.parse_args()

dist = "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1
local = int(os.environ.get("LOCAL_RANK", "0")) if dist else 0
torch.cuda.set_device(local)
if dist:
    import torch.distributed as td
    td.init_process_group("nccl", device_id=torch.device("cuda", local))

n, m = args.n, args.m
ctx = S.Context(local)
t0 = time.perf_counter()
ctx.geno_begin(n, m)
ctx.geno_synth(m, seed=20261017, miss_rate=args.miss)
t_synth = time.perf_counter() - t0
free0, total = ctx.mem_info()

wins = list(ctx.windows(args.rows))
mine = [w for k, w in enumerate(wins) if window_owner(k, args.world) == args.rank]
if args.max_windows:
    mine = mine[:args.max_windows]
max_cnt = 0
for r0, h in mine:
    ctx.set_row_window(r0, h)
    max_cnt = max(max_cnt, ctx.window_count())
host = torch.empty(max_cnt, dtype=torch.float64, pin_memory=True)
hb = host.numpy()

if dist:
    td.barrier()
torch.cuda.synchronize()
t_start = time.perf_counter()
per_win, hot_ms, launches0, pairs = [], 0.0, ctx.kernel_launches(), 0
first = None
for k, (r0, h) in enumerate(mine):
    tw = time.perf_counter()
    ctx.set_row_window(r0, h)
    cnt = ctx.window_count()
    out, _ = ctx.grm(args.method, packed=True, out=hb)
    hot_ms += ctx.last_hot_kernel()[0]
    pairs += cnt
    per_win.append(round((time.perf_counter() - tw) * 1e3, 1))
    if first is None and r0 == 0:
        first = out[: min(cnt, 4 * n)].copy()      # rows 0..3 of the packed triangle
torch.cuda.synchronize()
t_job = time.perf_counter() - t_start
if dist:
    tt = torch.tensor([t_job], dtype=torch.float64, device="cuda")
    td.all_reduce(tt, op=td.ReduceOp.MAX)
    t_job_max = float(tt[0])
    pp = torch.tensor([float(pairs)], dtype=torch.float64, device="cuda")
    td.all_reduce(pp)
    pairs_all = float(pp[0])
else:
    t_job_max, pairs_all = t_job, float(pairs)
ctx.set_row_window(0, 0)
pl = ctx.last_plan()

check = None
if args.rank == 0 and first is not None and args.check > 0:
    # oracle (checker only) on the first rows against the device's own all-sample SNP statistics
    from oracle import snprel_oracle as O
    k = min(args.check, n)
    sub = O.synth_geno(k, m, seed=20261017, miss_rate=args.miss)           # samples 0..k-1, all SNPs
    af, _, _ = ctx.snp_ratefreq()
    mu = 2 * af
    poly = (af > 0) & (af < 1)
    w = np.where(poly, 1.0 / np.where(poly, af * (1 - af), 1.0), 0.0)
    z = np.where(sub <= 2, (sub - mu[:, None]) * np.sqrt(w)[:, None], 0.0)
    mm = (sub > 2).astype(np.float64)
    miss = mm * poly[:, None]
    den = miss.sum(0)[:, None] + miss.sum(0)[None, :] - miss.T @ mm
    ref = (z.T @ z) / (2.0 * (poly.sum() - den))
    err = 0.0
    for i in range(min(k, 4)):
        base = i * n - i * (i - 1) // 2                 # packed index of (i, i)
        got = first[base: base + (k - i)]
        err = max(err, float(np.max(np.abs(got - ref[i, i:k]) / np.maximum(np.abs(ref[i, i:k]), 1.0))))
    check = {"rows": min(k, 4), "cols": k, "max_rel_err": err}

if args.rank == 0:
    total_pairs = n * (n + 1) / 2
    print(json.dumps({
        "workload": f"snpgdsGRM {args.method}, synthetic {n} samples x {m} SNPs, missing {args.miss}, "
                    f"N x N output tiled in {args.rows}-row windows dealt boustrophedon to {args.world} rank(s)",
        "rank": args.rank, "world": args.world, "torchrun": dist, "windows_this_rank": len(mine), "windows_total": len(wins),
        "synth_s": round(t_synth, 2), "job_s": round(t_job_max, 3), "hot_kernel_s": round(hot_ms / 1e3, 3),
        "pairs_this_job": pairs_all, "share_of_all_pairs": pairs_all / total_pairs,
        "pair_snps_per_s": pairs_all * m / t_job_max,
        "pair_snps_per_s_per_gpu": pairs_all * m / t_job_max / (args.world if dist else 1),
        "digits": [int(pl.digits), int(pl.digits_w), int(pl.digits_d)], "frac_bits": [int(pl.frac_bits), int(pl.frac_bits_w)],
        "kernel_launches": ctx.kernel_launches() - launches0,
        "free_gib_after_load": round(free0 / 2**30, 1), "free_gib_end": round(ctx.mem_info()[0] / 2**30, 1),
        "ms_per_window": per_win, "oracle_check": check}))
if dist:
    td.destroy_process_group()
