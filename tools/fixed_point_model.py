"""CPU model (numpy, exact integers) of the fixed-point table-Gram arithmetic of csrc/grm.cu, used to
evaluate format changes before touching the kernel.  Two schemes on the bench generator
(MAF ~ U(0.05, 0.5), 0.5 % missing), Eigenstrat weights:

  A (round 1)  C = sum U[g_i] x_j - sum W[g_i] + sum W[g_i] m_j,            U 5 digits, W 4 digits
  B (shipped since round 2) C = sum T[g_i] B_l[g_j] - sum D[g_i] + sum D[g_i] m_j        T 5 digits, D 3 digits
  B' (the default from ~1e5 SNPs on: snprel_set_rounding 'auto') the same with T rounded at random
               (the library's own counter-based draws, keyed by the global SNP index)      T 4 digits, D 3 digits
               with per-SNP integer column tables B_l[g] = s_l g - t_l (|B| <= 127, 0 for missing),
               T = U / s_l, D = (mu_l - t_l / s_l) U: the centring of the column sample moves into
               the main passes and the missing-data pass only carries the rational-approximation
               residual of the SNP mean.

Every digit pass is an exact integer Gram (evaluated here with float64 BLAS on int8-sized digits,
|sum| < 2^53), so the printed errors are exactly what the tensor-core path would produce.

    python tools/fixed_point_model.py [n_samples=48] [n_snps=1000000]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import snprel_oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
g = O.synth_geno(n, m, seed=20261017, miss_rate=0.005)          # [m, n]
valid = g <= 2
x = np.where(valid, g, 0).astype(np.float64)
mis = (~valid).astype(np.float64)
# the real workload takes the SNP statistics over all 10 000 samples; with the few samples modelled
# here the sample frequencies would scatter widely, so mu / w come from the generator's population
# frequencies (same formula as oracle.synth_geno) -- what the 10 000-sample estimates converge to
with np.errstate(over="ignore"):
    hp = O._splitmix64(np.uint64(20261017) ^ (np.arange(m, dtype=np.uint64) * np.uint64(0xD1342543DE82EF95)))
p = 0.05 + 0.45 * ((hp >> np.uint64(11)).astype(np.float64) / 9007199254740992.0)
mu = 2 * p
w = 1.0 / (p * (1 - p))
U = w[:, None] * (np.arange(3)[None, :] - mu[:, None])           # [m, 3]
code = np.where(valid, g, 3).astype(np.int64)                    # 3 = missing


def lookup(tab3):
    """tab3 [m, 3] -> values per (snp, sample), 0 for missing"""
    t4 = np.concatenate([tab3, np.zeros((m, 1))], axis=1)
    return np.take_along_axis(t4, code, axis=1)


# float64 reference: C = sum_l w_l (x_i - mu v_i)(x_j - mu v_j)
Cref = np.zeros((n, n))
for a in range(0, m, 100000):
    zc = (x[a:a + 100000] - mu[a:a + 100000, None] * valid[a:a + 100000]) * np.sqrt(w[a:a + 100000])[:, None]
    Cref += zc.T @ zc
scale = np.trace(Cref) / (n - 1)


def digits_of(q, nd):
    """balanced base-256 digits of an int64 array (as float64 arrays) + overflow flag"""
    q = q.copy()
    out = []
    for _ in range(nd):
        d = ((q + 128) & 255) - 128
        out.append(d.astype(np.float64))
        q = (q - d) >> 8
    return out, bool(np.any(q != 0))


def table_gram(qtab3, nd, bvals):
    """exact sum_l q[l][g_il] * b[l, j] via nd int8 digit passes; qtab3 int64 [m, 3], bvals float64 [m, n]"""
    t4 = np.concatenate([qtab3, np.zeros((m, 1), dtype=np.int64)], axis=1)
    qv = np.take_along_axis(t4, code, axis=1)                    # [m, n] int64
    ds, ovf = digits_of(qv, nd)
    acc = np.zeros((n, n), dtype=object)
    tot = np.zeros((n, n), dtype=np.float64)
    res = [None] * nd
    for k, d in enumerate(ds):
        r = np.zeros((n, n))
        for a in range(0, m, 200000):
            r += d[a:a + 200000].T @ bvals[a:a + 200000]
        res[k] = r
    # combine exactly with Python integers
    big = np.zeros((n, n), dtype=object)
    for k, r in enumerate(res):
        big = big + (np.rint(r).astype(np.int64).astype(object) << (8 * k))
    return big, ovf


def dither_u01(origin, shape):
    """the library's draw for table entry (global SNP index origin + l, genotype g): grm.cu:dither_u01"""
    M64 = np.uint64(0xFFFFFFFFFFFFFFFF)
    l = (np.arange(shape[0], dtype=np.uint64) + np.uint64(origin))[:, None]
    g = np.arange(shape[1], dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        x = np.uint64(0xD17E5 | 1) ^ ((l * np.uint64(4) + g) * np.uint64(0xD1342543DE82EF95))
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return (x >> np.uint64(11)).astype(np.float64) / 9007199254740992.0


def run(name, T3, f, nT, Bv, D3, fw, nD, ew, maxmiss, dither=False, s2=None, origin=0):
    if dither:   # unbiased randomised rounding, one draw per table entry: floor(v 2^f + u), u ~ U[0, 1)
        qT = np.floor(T3 * 2.0 ** f + dither_u01(origin, T3.shape)).astype(np.int64)
    else:
        qT = np.rint(T3 * 2.0 ** f).astype(np.int64)
    qD = np.rint(D3 * 2.0 ** fw).astype(np.int64)
    main, o1 = table_gram(qT, nT, Bv)
    corr, o2 = table_gram(qD, nD, mis)
    vecD = lookup(D3).sum(axis=0)                                # per-sample vector (extended precision on the GPU)
    C = np.array(main, dtype=np.float64) / 2.0 ** f + np.array(corr, dtype=np.float64) / 2.0 ** fw - vecD[:, None]
    err = np.max(np.abs(C - Cref)) / scale
    bound = (2.0 ** -(f + 1) * ew + 2.0 ** -(fw + 1) * maxmiss) / scale
    if dither:   # Hoeffding over the rounding draws, union bound over 1e4^2 pairs, failure probability 1e-12
        bound = (2.0 ** -f * np.sqrt(0.5 * s2 * np.log(2 * 1e8 / 1e-12)) + 2.0 ** -(fw + 1) * maxmiss) / scale
    print(f"{name}: digits {nT}+{nD} (f={f}, fw={fw}) overflow={o1 or o2}  max |err|/scale = {err:.2e}   proven bound {bound:.2e}")


t0 = time.time()
# scheme A: U against x, W = mu U against m
W3 = mu[:, None] * U
ewA = float(np.max(x.sum(axis=0)))
maxmiss = float(np.max(mis.sum(axis=0)))
run("A (round 1) ", U, 33, 5, x, W3, 28, 4, ewA, maxmiss)

# scheme B: per-SNP (s, t): equalise |T| = |U| / s and pick the s in a window that best approximates mu by t / s
umax = np.max(np.abs(U), axis=1)
s_hi = np.minimum(np.floor(127.0 / np.maximum(2.0 - mu, 1e-9)), 127.0)     # 2 s - t <= 127 with t ~ s mu
UREF = 40.0          # shipped rule (grm.cu:coltab_kernel): |U| at MAF 0.05, a constant, so that every rank of a sharded run picks the same table
s_tgt = np.minimum(s_hi, np.maximum(24.0, 127.0 * umax / UREF))
best_s = np.zeros(m)
best_t = np.zeros(m)
best_e = np.full(m, np.inf)
for frac in np.linspace(0.7, 1.0, 31):
    s = np.maximum(1.0, np.floor(s_tgt * frac))
    t = np.rint(s * mu)
    ok = (2 * s - t <= 127) & (t <= 127)
    e = np.where(ok, np.abs(mu - t / s) * umax, np.inf)          # residual table magnitude
    better = e < best_e
    best_s, best_t, best_e = np.where(better, s, best_s), np.where(better, t, best_t), np.where(better, e, best_e)
s, t = best_s, best_t
delta = mu - t / s
T3 = U / s[:, None]
D3 = delta[:, None] * U
Bv = np.where(valid, s[:, None] * x - t[:, None], 0.0)
ewB = float(np.max(np.abs(Bv).sum(axis=0)))
fT = int(np.floor(np.log2((127 * (256.0 ** 5 - 1) / 255 - 1) / np.max(np.abs(T3)))))
fD = int(np.floor(np.log2((127 * (256.0 ** 3 - 1) / 255 - 1) / np.max(np.abs(D3)))))
print(f"   scheme B tables: s in [{s.min():.0f}, {s.max():.0f}], max|T| {np.max(np.abs(T3)):.3f}, max|D| {np.max(np.abs(D3)):.2e}, "
      f"err weight {ewB:.3g} (A: {ewA:.3g}), max missing {maxmiss:.0f}")
run("B (shipped) ", T3, fT, 5, Bv, D3, fD, 3, ewB, maxmiss)
fU4 = int(np.floor(np.log2((127 * (256.0 ** 4 - 1) / 255 - 1) / np.max(np.abs(U)))))
fT4 = int(np.floor(np.log2((127 * (256.0 ** 4 - 1) / 255 - 1) / np.max(np.abs(T3)))))
print("   randomised rounding of the main table (probabilistic bound, failure probability 1e-12):")
run("A', U 4 digits dithered", U, fU4, 4, x, W3, 28, 4, ewA, maxmiss, True, float(np.max((x * x).sum(axis=0))))
s2B = float(np.max((np.ceil(Bv * Bv / 193.0) * 193.0).sum(axis=0)))      # as measured by sample_stats_kernel (unit 193)
for origin in (0, 1000003, 987654321, 2 ** 41 + 5, 31337):         # shipped default at this size ('auto' -> randomised)
    run(f"B', T 4 digits dithered, SNP origin {origin}", T3, fT4, 4, Bv, D3, fD, 3, ewB, maxmiss, True, s2B, origin)
print(f"   ({time.time() - t0:.0f} s)   target: 1e-10")
