"""BASELINE config-2 size (10 000 samples x 1 000 000 SNPs, the bench workload) through the C ABI:
entries at SCATTERED sample indices (first / middle / last 256-sample tile rows, so the tile-index
arithmetic of every tile row and the ragged last tile are covered) against the oracle evaluated on those samples' columns (the per-SNP
statistics over ALL samples come from the device, themselves checked against the generator), plus
size-independent properties of the whole matrix.  The same checks run at configs 3-5 in
tools/config_run.py and tools/c5_tiled.py (results in profiles/).  Sorted last: ~1 minute."""
import numpy as np
import pytest

import snprelate_b200 as S
from oracle import snprel_oracle as O

pytestmark = pytest.mark.gpu

N, M, K, SEED, MISS = 10000, 1000000, 48, 20261017, 0.005
IDX = O.scattered_samples(N, K, seed=11)


@pytest.fixture(scope="module")
def ws():
    ctx = S.Context(0)
    ctx.geno_begin(N, M)
    ctx.geno_synth(M, seed=SEED, miss_rate=MISS)
    sub = O.synth_geno(0, M, seed=SEED, miss_rate=MISS, samples=IDX)   # K scattered samples of the same data set
    yield ctx, sub
    ctx.close()


def test_generator_and_snp_statistics(ws):
    ctx, sub = ws
    back = ctx.geno_copy_2b()
    codes = (back[:, IDX // 4] >> (2 * (IDX % 4)).astype(np.uint8)[None, :]) & 3
    assert np.array_equal(codes, np.minimum(sub, 3))              # device generator == oracle generator
    af, maf, mr = ctx.snp_ratefreq()
    assert np.all((af > 0.02) & (af < 0.55)) and abs(float(mr.mean()) - MISS) < 2e-4
    assert abs(float(af.mean()) - 0.275) < 2e-3                   # MAF ~ U(0.05, 0.5)


def test_gcta_grm_full_size(ws):
    ctx, sub = ws
    grm, _ = ctx.grm("GCTA")
    assert grm.shape == (N, N) and np.array_equal(grm, grm.T)
    d = np.diag(grm)
    assert np.all(d > 0.9) and np.all(d < 1.1) and abs(float(d.mean()) - 1.0) < 1e-3    # E[G_ii] = 1 under HWE
    off_mean = (float(grm.sum()) - float(d.sum())) / (N * (N - 1.0))
    assert abs(off_mean + 1.0 / (N - 1)) < 1e-5                   # centred genotypes: rows sum to ~0
    af, _, _ = ctx.snp_ratefreq()
    ref = O.subset_entries(sub, af, "GCTA")
    got = grm[np.ix_(IDX, IDX)]
    assert IDX[0] == 0 and IDX[-1] == N - 1 and np.sum((IDX > 4000) & (IDX < 6000)) >= 8
    assert float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0))) < 1e-10
    pl = ctx.last_plan()
    assert pl.digits + pl.digits_w + pl.digits_d <= 10            # the fixed-point plan of DESIGN.md section 3


def test_ibs_counts_full_size(ws):
    ctx, sub = ws
    i0, i1, i2 = ctx.ibs_num()
    ref = O.ibs_counts(sub)
    ix = np.ix_(IDX, IDX)
    assert np.array_equal(np.stack([i0[ix], i1[ix], i2[ix]]), ref)                      # bit exact
    tot = (i0.astype(np.int64) + i1 + i2)
    assert np.array_equal(tot, tot.T) and int(tot.max()) <= M
    assert np.array_equal(np.diag(i2), np.diag(tot))              # a sample is IBS2 with itself wherever it is valid


def test_pca_genmat_and_top32_eigenvectors_full_size(ws):
    """Config 2 proper: snpgdsPCA covariance + top-32 eigenvectors.  genmat at the scattered samples vs the
    oracle (1e-10); eigenvectors vs scipy.linalg.eigh of the device genmat (north star: 1e-6, up to sign)."""
    import scipy.linalg as sla
    ctx, sub = ws
    r = ctx.pca(eigen_cnt=32, need_genmat=True)
    gm = r["genmat"]
    af, _, _ = ctx.snp_ratefreq()
    ref = O.subset_entries(sub, af, "Eigenstrat", n_total=N, trace=r["TraceXTX"])
    got = gm[np.ix_(IDX, IDX)]
    assert float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0))) < 1e-10
    assert abs(float(np.trace(gm)) - (N - 1)) < 1e-6 * N and abs(r["TraceVal"] - float(np.trace(gm))) < 1e-6
    w, v = sla.eigh(gm, subset_by_index=[N - 33, N - 1])          # one more: the gap below the 32nd
    w, v = w[::-1], v[:, ::-1]
    assert np.max(np.abs(r["eigenval"][:32] - w[:32]) / w[:32]) < 1e-9
    V = r["eigenvect"]
    for k in range(32):
        sgn = 1.0 if float(V[:, k] @ v[:, k]) >= 0 else -1.0
        # an eigenvector is conditioned by the gap to its neighbours (both solvers stop at residuals of
        # ~1e-11 |C|): 1e-6 wherever the gap allows it, which it does for all 32 here
        gap = min(w[k - 1] - w[k] if k else np.inf, w[k] - w[k + 1])
        tol = max(1e-6, 1e-10 * w[0] / gap)
        assert float(np.max(np.abs(sgn * V[:, k] - v[:, k]))) < tol, (k, gap, tol)
    resid = np.linalg.norm(gm @ V - V * r["eigenval"][None, :32], axis=0)
    assert float(resid.max()) < 1e-9 * w[0]


def test_rounding_modes_and_draws_full_size(ws):
    """Config 2 under the default rounding mode ('auto'): randomised rounding of the main row table saves one
    of eight tensor passes here (Hoeffding bound, failure probability 1e-12).  The
    1e-10 tolerance is checked against the oracle for round-to-nearest and for four independent draws
    (the SNP origin keys the draws), all at the scattered samples."""
    ctx, sub = ws
    af, _, _ = ctx.snp_ratefreq()
    ix = np.ix_(IDX, IDX)

    def run():
        r = ctx.pca(genmat_only=True)
        ref = O.subset_entries(sub, af, "Eigenstrat", n_total=N, trace=r["TraceXTX"])
        pl = ctx.last_plan()
        return float(np.max(np.abs(r["genmat"][ix] - ref) / np.maximum(np.abs(ref), 1.0))), pl, r["genmat"][ix].copy()

    try:
        ctx.set_rounding("nearest")
        e0, p0, g0 = run()
        assert p0.rounding == 0 and e0 < 1e-10
        ctx.set_rounding("auto")
        errs, mats = [], []
        for origin in (0, 1000003, 987654321, 2 ** 41 + 5):
            ctx.set_snp_origin(origin)
            e, pl, gm = run()
            assert pl.rounding == 1, "auto should pick randomised rounding at 1M SNPs"
            assert pl.digits + pl.digits_w + pl.digits_d == p0.digits + p0.digits_w + p0.digits_d - 1 == 7
            assert e < 1e-10, (origin, e)
            errs.append(e)
            mats.append(gm)
        assert not np.array_equal(mats[0], mats[1])
        print(f"config 2: nearest {e0:.2e} (8 passes), randomised draws {', '.join(f'{e:.2e}' for e in errs)} (7 passes)")
    finally:
        ctx.set_rounding("auto")
        ctx.set_snp_origin(0)
