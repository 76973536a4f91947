"""BASELINE config 5 (snpgdsGRM GCTA, 500k samples x 800k SNPs, N x N output tiled across GPUs)
or any other tiled run: every rank ends up holding the whole 2-bit matrix (100 GB at C5) -- under torchrun
each rank loads only its own SNP block and the others arrive over NVLink (snprelate_b200.dist.
load_sharded_then_gather: 100 GB cross PCIe in total instead of 800) -- the row windows of
the N x N output are dealt to the ranks in boustrophedon order, no collective in the compute phase.  On one GPU `--world 8 --rank r`
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
ap.add_argument("--replicated-load", action="store_true", help="every rank loads ALL SNPs itself (round 1 behaviour) instead of "
                "loading its own SNP block and gathering the others over NVLink")
ap.add_argument("--check", type=int, default=48, help="oracle check on K scattered samples (first / middle / last tile rows); "
                "every rank checks the entries of its own windows")
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
gathered_bytes = 0
if dist and not args.replicated_load:
    # SNP-sharded residency: this rank generates (stands in for: reads from the host) only its SNP block;
    # the other blocks come over NVLink
    from snprelate_b200 import dist as D
    gathered_bytes = D.load_sharded_then_gather(
        ctx, n, m, lambda lo, hi: ctx.geno_synth(hi - lo, seed=20261017, miss_rate=args.miss, snp_start=lo),
        args.rank, args.world, device=torch.device("cuda", local))
else:
    ctx.geno_begin(n, m)
    ctx.geno_synth(m, seed=20261017, miss_rate=args.miss)
torch.cuda.synchronize()
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
from oracle import snprel_oracle as O          # checker only (not on the timed path's arithmetic)
idx = O.scattered_samples(n, args.check, seed=7) if args.check > 0 else np.zeros(0, dtype=np.int64)
got_rows = []                                  # (position a in idx, entries (idx[a], idx[a:]) of the packed slice)
for k, (r0, h) in enumerate(mine):
    tw = time.perf_counter()
    ctx.set_row_window(r0, h)
    cnt = ctx.window_count()
    out, _ = ctx.grm(args.method, packed=True, out=hb)
    hot_ms += ctx.last_hot_kernel()[0]
    pairs += cnt
    per_win.append(round((time.perf_counter() - tw) * 1e3, 1))
    pbase = r0 * (2 * n - r0 - 1) // 2 + r0         # packed index of (r0, r0)
    for a, i in enumerate(idx):
        if r0 <= i < min(r0 + h, n):
            base = int(i) * (2 * n - int(i) - 1) // 2 - pbase
            got_rows.append((a, out[base + idx[a:]].copy()))
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
if len(idx):
    # oracle on the scattered samples' columns against the device's own all-sample SNP statistics;
    # every rank checks the rows that fell into its windows, the verdict is the max over ranks
    sub = O.synth_geno(0, m, seed=20261017, miss_rate=args.miss, samples=idx)
    af, _, _ = ctx.snp_ratefreq()
    ref = O.subset_entries(sub, af, "GCTA")
    err, cnt_chk = 0.0, 0
    for a, vals in got_rows:
        r = ref[a, a:]
        err = max(err, float(np.max(np.abs(vals - r) / np.maximum(np.abs(r), 1.0))))
        cnt_chk += len(r)
    if dist:
        tt = torch.tensor([err, -float(cnt_chk)], dtype=torch.float64, device="cuda")
        td.all_reduce(tt, op=td.ReduceOp.MAX)
        err = float(tt[0])
        cc = torch.tensor([float(cnt_chk)], dtype=torch.float64, device="cuda")
        td.all_reduce(cc)
        cnt_chk = int(cc[0])
    check = {"scattered_samples": int(len(idx)), "entries_checked": int(cnt_chk), "max_rel_err": err,
             "first_last_sample": [int(idx[0]), int(idx[-1])], "ok": bool(err < 1e-10)}

if args.rank == 0:
    total_pairs = n * (n + 1) / 2
    print(json.dumps({
        "workload": f"snpgdsGRM {args.method}, synthetic {n} samples x {m} SNPs, missing {args.miss}, "
                    f"N x N output tiled in {args.rows}-row windows dealt boustrophedon to {args.world} rank(s)",
        "rank": args.rank, "world": args.world, "torchrun": dist, "windows_this_rank": len(mine), "windows_total": len(wins),
        "load_s": round(t_synth, 2), "bytes_gathered_from_peers_rank0": int(gathered_bytes),
        "loading": "own SNP block + NCCL broadcast gather" if gathered_bytes else "every rank loads all SNPs", "job_s": round(t_job_max, 3), "hot_kernel_s": round(hot_ms / 1e3, 3),
        "pairs_this_job": pairs_all, "share_of_all_pairs": pairs_all / total_pairs,
        "pair_snps_per_s": pairs_all * m / t_job_max,
        "pair_snps_per_s_per_gpu": pairs_all * m / t_job_max / (args.world if dist else 1),
        "digits": [int(pl.digits), int(pl.digits_w), int(pl.digits_d)], "frac_bits": [int(pl.frac_bits), int(pl.frac_bits_w)],
        "kernel_launches": ctx.kernel_launches() - launches0,
        "free_gib_after_load": round(free0 / 2**30, 1), "free_gib_end": round(ctx.mem_info()[0] / 2**30, 1),
        "ms_per_window": per_win, "oracle_check": check}))
if dist:
    td.destroy_process_group()
