"""Larger-than-test-size checks on one GPU: row-window mode vs whole-matrix mode at N = 30k,
and timings of the packed-bit estimators at BASELINE config-3-like sizes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import snprelate_b200 as S

ctx = S.Context(0)
n, m = 30000, 100000
ctx.geno_begin(n, m)
ctx.geno_synth(m, seed=5, miss_rate=0.005)
t0 = time.time(); full = ctx.ibs_ave(packed=True); t_full = time.time() - t0
t0 = time.time(); win = ctx.packed_by_windows(lambda: ctx.ibs_ave(packed=True), 8192); t_win = time.time() - t0
print(f"IBS N={n} M={m}: whole {t_full:.2f}s  windows(8192) {t_win:.2f}s  identical={np.array_equal(full, win, equal_nan=True)}", flush=True)
t0 = time.time(); full = ctx.grm("GCTA", packed=True)[0]; t_full = time.time() - t0
t0 = time.time(); win = ctx.packed_by_windows(lambda: ctx.grm("GCTA", packed=True)[0], 8192); t_win = time.time() - t0
print(f"GCTA N={n} M={m}: whole {t_full:.2f}s  windows(8192) {t_win:.2f}s  identical={np.array_equal(full, win)}", flush=True)
del full, win
# config-3-like: 50k samples; SNP count scaled to keep the run short
n, m = 50000, 100000
ctx.geno_begin(n, m)
ctx.geno_synth(m, seed=6, miss_rate=0.005)
for est, name in ((10, "IBS"), (11, "KING-robust")):
    ms = ctx.time_accumulate(est, 1)
    hot, _, units = ctx.last_hot_kernel()
    print(f"{name} N={n} M={m}: step {ms:.0f} ms, pair kernel {hot:.0f} ms, {units / hot * 1e3:.3e} pair-SNPs/s, free mem {ctx.mem_info()[0] / 2**30:.1f} GiB", flush=True)
