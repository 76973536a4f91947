"""Eigen step on two 2304-sample data sets (with / without population structure): wall time of
snprel_pca, the solver used, filter rounds, block products and phase times; then the dense solver."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import snprelate_b200 as S
from oracle import snprel_oracle as O      # generator + checker only


def three_populations(n, m, seed):
    rng = np.random.default_rng(seed)
    p0 = rng.uniform(0.1, 0.9, m)
    pops = rng.integers(0, 3, n)
    shift = rng.normal(0, 0.08, (3, m))
    g = np.empty((m, n), dtype=np.uint8)
    for q in range(3):
        idx = np.nonzero(pops == q)[0]
        p = np.clip(p0 + shift[q], 0.02, 0.98)
        g[:, idx] = (rng.random((m, idx.size)) < p[:, None]).astype(np.uint8) + (rng.random((m, idx.size)) < p[:, None])
    g[rng.random((m, n)) < 0.003] = 3
    return g


ctx = S.Context(0)
for structured in (True, False):
    g = three_populations(2304, 6000, 11) if structured else O.synth_geno(2304, 6000, seed=21, miss_rate=0.002)
    ctx.geno_begin(g.shape[1], g.shape[0])
    ctx.geno_push_u8(g)
    t0 = time.time()
    r1 = ctx.pca(eigen_cnt=16, need_genmat=True)
    t1 = time.time()
    print(f"structured={structured}: pca {t1 - t0:.2f} s, eigen info {ctx.last_eigen_info()}, phases {ctx.eigen_phase_ms}", flush=True)
    ctx.debug_flags(4)
    t0 = time.time()
    r0 = ctx.pca(eigen_cnt=16)
    print(f"   dense solver: pca {time.time() - t0:.2f} s, max |eigenvalue difference| "
          f"{np.max(np.abs(r0['eigenval'][:16] - r1['eigenval'][:16])):.2e}", flush=True)
    ctx.debug_flags(0)
