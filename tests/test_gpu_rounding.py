"""Rounding of the main row table of the covariance path (snprel_set_rounding): 'nearest' (worst-case
error bound), 'random' (unbiased randomised rounding, Hoeffding bound with failure probability 1e-12)
and 'auto', the default, which takes randomised rounding only where it needs FEWER tensor passes.
The tolerance of the north star (1e-10 relative on GRM entries) is checked against the oracle for every
mode and over several independent draws; the draws are keyed by the global SNP index, so results are
reproducible and SNP shards reproduce the one-context result bit for bit.  The config-2-size version
of the draws test is in tests/test_gpu_zfull_size.py."""
import numpy as np
import pytest

import snprelate_b200 as S
from snprelate_b200 import dist as D
from oracle import snprel_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-10


def relerr(got, ref):
    return float(np.nanmax(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)))


def passes(pl):
    return pl.digits + pl.digits_w + pl.digits_d


def test_every_mode_meets_the_tolerance_small():
    g = O.synth_geno(300, 20000, seed=5, miss_rate=0.01)
    ref, pref = O.grm_gcta(g), O.pca_genmat(g)[0]
    with S.Context(0) as c:
        c.geno_begin(g.shape[1], g.shape[0])
        c.geno_push_u8(g)
        seen = {}
        for mode in ("nearest", "random", "auto"):
            c.set_rounding(mode)
            got, _ = c.grm("GCTA")
            pl = c.last_plan()
            again, _ = c.grm("GCTA")
            assert relerr(got, ref) < TOL, mode
            assert np.array_equal(got, again)            # the draws are a pure function of (SNP index, genotype)
            assert pl.rounding == {"nearest": 0, "random": 1}.get(mode, pl.rounding)
            seen[mode] = (passes(pl), pl.rounding)
            r = c.pca(eigen_cnt=4, need_genmat=True)
            assert relerr(r["genmat"], pref) < TOL, mode
        assert seen["random"][0] <= seen["nearest"][0]
        # auto on a small problem (n_samp^2 n_snp < 2^36): round to nearest, whatever randomised rounding would save
        assert seen["auto"] == seen["nearest"]
        with pytest.raises(S.SNPRelError):
            c.set_rounding(3)


def test_randomised_rounding_saves_a_pass_at_400k_snps_over_several_draws():
    """1024 samples x 400 000 SNPs: from ~1e5 SNPs on the Hoeffding bound needs one digit of T less than the
    worst-case bound.  Entries at scattered samples vs the oracle for round-to-nearest and for four
    independent draws (shifted SNP origins re-draw every table entry)."""
    n, m, seed, miss = 1024, 400000, 314159, 0.005
    idx = O.scattered_samples(n, 40, seed=2)
    sub = O.synth_geno(0, m, seed=seed, miss_rate=miss, samples=idx)
    with S.Context(0) as c:
        c.geno_begin(n, m)
        c.geno_synth(m, seed=seed, miss_rate=miss)
        af, _, _ = c.snp_ratefreq()
        ix = np.ix_(idx, idx)
        c.set_rounding("nearest")
        r = c.pca(genmat_only=True)
        ref = O.subset_entries(sub, af, "Eigenstrat", n_total=n, trace=r["TraceXTX"])
        p0 = c.last_plan()
        assert p0.rounding == 0 and relerr(r["genmat"][ix], ref) < TOL
        gref = O.subset_entries(sub, af, "GCTA")
        c.set_rounding("random")
        results, prand = [], None
        for origin in (0, 400000, 123456789, 2 ** 40 + 17):
            c.geno_begin(n, m)
            c.set_snp_origin(origin)
            c.geno_synth(m, seed=seed, miss_rate=miss)
            r = c.pca(genmat_only=True)
            prand = c.last_plan()
            assert prand.rounding == 1 and passes(prand) <= passes(p0)
            assert relerr(r["genmat"][ix], ref) < TOL, origin
            assert np.array_equal(r["genmat"], r["genmat"].T)
            assert relerr(c.grm("GCTA")[0][ix], gref) < TOL, origin
            results.append(r["genmat"][ix].copy())
        assert not np.array_equal(results[0], results[1])           # different draws ...
        assert relerr(results[0], results[1]) < TOL                  # ... of the same matrix
        assert passes(prand) == passes(p0) - 1, (passes(prand), passes(p0))   # the digit the sqrt(M) bound saves
        c.set_rounding("auto")
        c.pca(genmat_only=True)
        pa = c.last_plan()
        assert (passes(pa), pa.rounding) == (passes(prand), 1)


def test_shards_reproduce_the_one_context_result_under_random_rounding():
    """Two contexts with half of the SNPs each and the shard offsets as SNP origins: the tables are
    drawn per GLOBAL SNP index, so the reduced planes equal the one-context planes bit for bit."""
    n, m = 600, 12000
    g = O.synth_geno(n, m, seed=8, miss_rate=0.02)
    cut = D.shard_range(m, 0, 2)[1]
    ctxs = [S.Context(0), S.Context(0)]
    try:
        for c, (lo, hi) in zip(ctxs, ((0, cut), (cut, m))):
            c.set_rounding("random")
            c.geno_begin(n, hi - lo)
            c.set_snp_origin(lo)
            c.geno_push_u8(g[lo:hi])
        plan = D.accumulate_in_process(ctxs, 1)
        a = ctxs[0].grm("GCTA")[0]
        assert ctxs[0].last_plan().rounding == 1 and relerr(a, O.grm_gcta(g)) < TOL
        with S.Context(0) as one:
            one.set_rounding("random")
            one.geno_begin(n, m)
            one.geno_push_u8(g)
            p1 = one.plan_local(1)
            # same format as the merged plan -> same integers
            for k in ("max_abs", "max_abs_w", "sum_bound", "err_weight", "err_weight2", "scale", "total_missing",
                      "max_missing", "n_snp", "diag_bound", "sum_rest"):
                setattr(p1, k, getattr(plan, k))
            one.accumulate(1, p1)
            one.mark_reduced()                 # finish from these accumulators (nothing to reduce: one shard)
            whole = one.grm("GCTA")[0]
        assert np.array_equal(a, whole)
    finally:
        for c in ctxs:
            c.close()


def test_plan_statistics_match_a_numpy_restatement():
    """The statistics every error bound rests on -- max_j sum_l |B_l[g_jl]|, max_j sum_l B_l[g_jl]^2 (in units of
    193, rounded up per SNP), the per-sample missing counts -- against a numpy restatement of the per-SNP integer
    column tables (grm.cu:coltab_kernel).  The device may pick a neighbouring s for a handful of SNPs (fused
    multiply-adds in the candidate loop), hence the small relative tolerance on the two sums; they must never be
    BELOW what the data says by more than that."""
    n, m = 333, 6000
    g = O.synth_geno(n, m, seed=12, miss_rate=0.03, maf_lo=0.02)
    valid = g <= 2
    x = np.where(valid, g, 0).astype(np.float64)
    num = valid.sum(axis=1)
    mu = x.sum(axis=1) / np.maximum(num, 1)
    sfrq = mu * 0.5
    ok = (sfrq > 0) & (sfrq < 1)
    r = 1.0 / np.sqrt(np.where(ok, sfrq * (1 - sfrq), 1.0))
    w = np.where(ok, r * r, 0.0)
    umax = w * np.maximum(mu, 2.0 - mu)
    s_hi = np.minimum(np.floor(127.0 / np.maximum(2.0 - mu, 1e-9)), 127.0)
    s_tgt = np.minimum(s_hi, np.maximum(24.0, 127.0 * umax / 40.0))
    bs, bt, be = np.ones(m), np.rint(mu), np.full(m, 1e300)
    for q in range(31):
        sc = np.maximum(1.0, np.floor(s_tgt * (0.7 + 0.01 * q)))
        tc = np.rint(sc * mu)
        good = ~((2 * sc - tc > 127.0) | (tc > 127.0))
        e = np.where(good, np.abs(mu - tc / sc), np.inf)
        better = e < be
        bs, bt, be = np.where(better, sc, bs), np.where(better, tc, bt), np.where(better, e, be)
    live = (w > 0) & (num > 0)
    s, t = np.where(live, bs, 1.0), np.where(live, bt, 0.0)
    B = np.where(valid, s[:, None] * x - t[:, None], 0.0)
    ew = float(np.abs(B).sum(axis=0).max())
    s2 = float((np.ceil(B * B / 193.0) * 193.0).sum(axis=0).max())
    with S.Context(0) as c:
        c.geno_begin(n, m)
        c.geno_push_u8(g)
        p = c.plan_local(0)
    assert p.n_snp == m and p.total_missing == int((~valid).sum()) and p.max_missing == int((~valid).sum(axis=0).max())
    assert abs(p.err_weight - ew) <= 2e-3 * ew, (p.err_weight, ew)
    assert abs(p.err_weight2 - s2) <= 4e-3 * s2, (p.err_weight2, s2)
    assert p.err_weight <= p.err_weight2 <= 320.0 * p.err_weight     # 1 <= |B| <= 127 wherever B != 0, ceil to units of 193
    T = np.where(live[:, None], w[:, None] * (np.arange(3)[None, :] - mu[:, None]) / s[:, None], 0.0)
    assert abs(p.max_abs - float(np.abs(T).max())) <= 0.05 * p.max_abs
