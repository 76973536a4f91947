"""ONE process driving every GPU of the box through snprel_multi_* (what an R session does with
SNPREL_DEVICES=all): config 2 as one fixed problem, and config 4 (KING-robust 100k x 1M) in row windows with
the peer-memory reduction to a rotating root.  Prints one JSON line per leg."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import snprelate_b200 as S
from snprelate_b200._lib import EST_KING_ROBUST

ap = argparse.ArgumentParser()
ap.add_argument("--devices", default="all")
ap.add_argument("--king-samples", type=int, default=100000)
ap.add_argument("--king-snps", type=int, default=1000000)
ap.add_argument("--rows", type=int, default=8192)
ap.add_argument("--engine", default="tensor")
args = ap.parse_args()
nd = torch.cuda.device_count()
devs = list(range(nd)) if args.devices == "all" else [int(x) for x in args.devices.split(",")]

# ---- config 2, fixed problem
m = S.MultiContext(devs)
N, M = 10000, 1000000
m.geno_begin(N, M); m.geno_synth(M)
out = torch.empty((N, N), dtype=torch.float64, pin_memory=True).numpy()
ts = []
for rep in range(4):
    t0 = time.perf_counter(); m.accumulate("Eigenstrat", root=0); t1 = time.perf_counter()
    r = m.ctx(0).pca(genmat_only=True, genmat_out=out); t2 = time.perf_counter()
    ts.append((t1 - t0, t2 - t1, m.last_reduce()))
a, f, red = ts[-1]
print(json.dumps({"leg": "config 2 (10k x 1M PCA covariance) as one fixed problem, one process", "devices": devs,
                  "accumulate_reduce_ms": round(1e3 * min(t[0] for t in ts[1:]), 2), "finish_d2h_ms": round(1e3 * min(t[1] for t in ts[1:]), 2),
                  "reduce_ms": round(red[0], 3), "link_gb": round(red[1] / 1e9, 3),
                  "pair_snps_per_s": 0.5 * N * N * M / min(t[0] + t[1] for t in ts[1:])}), flush=True)
m.close()

# ---- config 4 in row windows
n, msnp = args.king_samples, args.king_snps
m = S.MultiContext(devs)
m.set_count_engine(args.engine)
m.geno_begin(n, msnp); m.geno_synth(msnp)
npad = (n + 255) // 256 * 256
wins = [(r0, min(args.rows, npad - r0)) for r0 in range(0, n, args.rows)]
maxcnt = n * args.rows
host = [torch.empty(maxcnt, dtype=torch.float64, pin_memory=True).numpy() for _ in range(2)]
for d in devs:
    torch.cuda.synchronize(d)
t0 = time.perf_counter(); t_acc = t_fin = 0.0; red_ms = 0.0; link = 0
for w, (r0, h) in enumerate(wins):
    m.set_row_window(r0, h)
    root = w % len(devs)
    ta = time.perf_counter(); m.accumulate(EST_KING_ROBUST, root=root); tb = time.perf_counter()
    m.ctx(root).king_robust(None, packed=True, out=host); tc = time.perf_counter()
    t_acc += tb - ta; t_fin += tc - tb; red_ms += m.last_reduce()[0]; link += m.last_reduce()[1]
job = time.perf_counter() - t0
print(json.dumps({"leg": f"config 4 (KING-robust {n} x {msnp}), one process, {len(wins)} windows of {args.rows} rows, engine {args.engine}",
                  "devices": devs, "job_s": round(job, 3), "accumulate_reduce_s": round(t_acc, 3), "finish_d2h_s": round(t_fin, 3),
                  "peer_reduce_s": round(red_ms / 1e3, 3), "link_gb": round(link / 1e9, 2),
                  "pair_snps_per_s": 0.5 * n * n * msnp / job}), flush=True)
