"""Timing of the loadings / projection / correlation calls (csrc/project.cu) through the C ABI."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import snprelate_b200 as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 32
ctx = S.Context(0)
ctx.geno_begin(n, m)
ctx.geno_synth(m, miss_rate=0.005)
rng = np.random.default_rng(1)
vec, _ = np.linalg.qr(rng.standard_normal((n, k)))
val = np.linspace(50.0, 2.0, k)
flop = 2.0 * n * m * k
for rep in range(2):
    t0 = time.perf_counter(); load, avg, scale = ctx.pca_snp_loading(val, vec, 2.0 * m, False); t1 = time.perf_counter()
    k1 = ctx.last_hot_kernel()[0]
    proj = ctx.pca_samp_loading(load, avg, scale); t2 = time.perf_counter()
    k2 = ctx.last_hot_kernel()[0]
    corr = ctx.pca_corr(vec[:, :2]); t3 = time.perf_counter()
    print(f"   kernels: snp_project {k1:.1f} ms ({flop/k1/1e9:.1f} TFLOP/s f64), samp_project {k2:.1f} ms ({flop/k2/1e9:.1f} TFLOP/s)")
    print(f"N={n} M={m} k={k}: snp_loading {1e3*(t1-t0):.1f} ms ({flop/(t1-t0)/1e12:.1f} TFLOP/s f64 incl. copies), "
          f"samp_loading {1e3*(t2-t1):.1f} ms ({flop/(t2-t1)/1e12:.1f}), corr(k=2) {1e3*(t3-t2):.1f} ms", flush=True)
