"""Quick device-time check of the accumulate phase at a given size (not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import snprelate_b200 as S
from snprelate_b200._lib import EST_IBS, EST_KING_ROBUST, EST_BETA

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
ests = sys.argv[3].split(",") if len(sys.argv) > 3 else ["pca", "gcta"]
miss = float(sys.argv[4]) if len(sys.argv) > 4 else 0.005
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
ctx = S.Context(0)
ctx.geno_begin(n, m)
t0 = time.time(); ctx.geno_synth(m, miss_rate=miss); print(f"synth {time.time()-t0:.2f}s", flush=True)
ids = {"pca": 0, "gcta": 1, "eigmix": 3, "ibs": EST_IBS, "king": EST_KING_ROBUST, "beta": EST_BETA}
for e in ests:
    t0 = time.time()
    ms = ctx.time_accumulate(ids[e], 1)        # first call of this estimator: includes its buffer allocations
    wall = time.time() - t0
    if reps > 1:
        again = []
        for _ in range(reps - 1):
            t1 = time.time()
            ms = ctx.time_accumulate(ids[e], 1)
            again.append((ms, (time.time() - t1) * 1e3, ctx.last_hot_kernel()[0]))
        print("   repeat calls (device step ms, wall ms, hot ms): " + "; ".join(f"{a:.1f}, {b:.1f}, {c:.1f}" for a, b, c in again))
    hot, nl, units = ctx.last_hot_kernel()
    pl = ctx.last_plan() if ids[e] < 10 else None
    if pl is not None:
        print(f"   plan: digits U{pl.digits}/W{pl.digits_w}/D{pl.digits_d} frac_bits {pl.frac_bits}/{pl.frac_bits_w}/{pl.frac_bits_d} max_abs {pl.max_abs:.3f}/{pl.max_abs_w:.3f} "
              f"err_weight {pl.err_weight:.4g} scale {pl.scale:.4g} sum_bound {pl.sum_bound:.4g} "
              f"missing {pl.total_missing} step {ctx.last_step_ms():.1f} ms")
    print(f"{e}: N={n} M={m} miss={miss} hot {hot:.1f} ms in {nl} launches, wall {wall*1e3:.1f} ms, "
          f"{units/hot*1e3:.3e} pair-SNPs/s (hot)", flush=True)
