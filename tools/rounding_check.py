"""Config-2-size check of the rounding modes of the main row table (snprel_set_rounding): fixed-point
format, tensor passes, device time of the accumulate phase and the error of the finished genmat at
scattered samples against the oracle, for round-to-nearest and for randomised rounding over several
draws (the draw is keyed by the SNP origin, so shifting the origin re-draws every table entry).

    python tools/rounding_check.py [n_samp=10000] [n_snp=1000000] [modes=nearest,random,auto] [origins=0,1000003]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import snprelate_b200 as S
from oracle import snprel_oracle as O       # checker only

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["nearest", "random", "auto"]
origins = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 1000003]
SEED, MISS, K = 20261017, 0.005, 64

ctx = S.Context(0)
ctx.geno_begin(n, m)
ctx.geno_synth(m, seed=SEED, miss_rate=MISS)
idx = O.scattered_samples(n, K, seed=5)
t0 = time.time()
sub = O.synth_geno(0, m, seed=SEED, miss_rate=MISS, samples=idx)
af, _, _ = ctx.snp_ratefreq()
cov = O.subset_entries(sub, af, "cov")
print(f"oracle on {K} scattered samples: {time.time() - t0:.1f} s", flush=True)

for mode in modes:
    ctx.set_rounding(mode)
    for org in (origins if mode != "nearest" else origins[:1]):
        if hasattr(ctx, "set_snp_origin"):
            ctx.set_snp_origin(org)
        ctx.invalidate()
        ms = [ctx.time_accumulate(0, 1) for _ in range(3)]
        hot, nl, _ = ctx.last_hot_kernel()
        pl = ctx.last_plan()
        r = ctx.pca(genmat_only=True)
        ref = cov * ((n - 1) / r["TraceXTX"])
        got = r["genmat"][np.ix_(idx, idx)]
        err = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)))
        rnd = getattr(pl, "rounding", "?")
        print(f"{mode:8s} origin {org:8d}: digits T{pl.digits} R{pl.digits_w} D{pl.digits_d} frac_bits {pl.frac_bits}/{pl.frac_bits_w} "
              f"rounding {rnd}  step ms {', '.join(f'{x:.1f}' for x in ms)}  K1 {hot:.1f} ms in {nl} launches  "
              f"max rel err {err:.3e}", flush=True)
