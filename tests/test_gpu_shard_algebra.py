"""The SNP-shard algebra of the multi-GPU path on ONE GPU: two contexts on device 0 each hold half
of the SNPs, their plans are merged exactly as dist.reduce_plan merges them across ranks, each
accumulates its shard, the reduce buffers are summed on the device and marked reduced -- then every
estimator's finish must return the whole-data result (integers bit-identical, float64 to 1e-10).
This is the reference's snpgdsMergeGRM identity (inst/unitTests/test_GRM.R:15-49); it runs where
tests/test_gpu_dist.py has to skip for want of a second GPU."""
import numpy as np
import pytest

import snprelate_b200 as S
from snprelate_b200 import dist as D
from snprelate_b200._lib import EST_IBS, EST_KING_ROBUST, EST_BETA
from oracle import snprel_oracle as O

pytestmark = pytest.mark.gpu


def relerr(got, ref):
    return float(np.nanmax(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)))


@pytest.fixture(scope="module")
def shards():
    n, m = 700, 9000
    g = O.synth_geno(n, m, seed=77, miss_rate=0.02)
    cut = D.shard_range(m, 0, 2)[1]
    assert cut % D.SNP_ALIGN == 0 and 0 < cut < m
    ctxs = [S.Context(0), S.Context(0)]
    yield g, cut, ctxs
    for c in ctxs:
        c.close()


def load(ctxs, g, cut, parts=None):
    parts = parts or [(0, cut), (cut, g.shape[0])]
    for c, (lo, hi) in zip(ctxs, parts):
        c.geno_begin(g.shape[1], hi - lo)
        c.set_snp_origin(lo)                # keys the rounding draws by the global SNP index (snprel_set_rounding)
        c.geno_push_u8(g[lo:hi])


@pytest.mark.parametrize("method,est,ref", [("GCTA", 1, O.grm_gcta), ("EIGMIX", 3, O.grm_eigmix),
                                            ("Eigenstrat", 0, O.grm_eigenstrat)])
def test_covariance_family_over_two_shards(shards, method, est, ref):
    g, cut, ctxs = shards
    load(ctxs, g, cut)
    plan = D.accumulate_in_process(ctxs, est)
    assert plan.n_snp == g.shape[0]
    r = ref(g)
    a, b = ctxs[0].grm(method)[0], ctxs[1].grm(method)[0]
    assert relerr(a, r) < 1e-10
    assert np.array_equal(a, b)                         # both "ranks" hold the same reduced planes
    # and it is the single-context answer, bit for bit (exact integer accumulation)
    with S.Context(0) as one:
        one.geno_begin(g.shape[1], g.shape[0])
        one.geno_push_u8(g)
        plan1 = one.plan_local(est)
        # the merged plan is conservative (sums of per-shard maxima): same digits or more
        assert plan.err_weight >= plan1.err_weight and plan.max_missing >= plan1.max_missing
        whole = one.grm(method)[0]
    assert relerr(a, whole) < 1e-12


def test_pca_over_two_shards(shards):
    g, cut, ctxs = shards
    load(ctxs, g, cut)
    D.accumulate_in_process(ctxs, 0)
    r = ctxs[1].pca(eigen_cnt=4, need_genmat=True)
    gm, tr, _ = O.pca_genmat(g)
    assert relerr(r["genmat"], gm) < 1e-10
    assert abs(r["TraceXTX"] - tr) / tr < 1e-12
    w = np.linalg.eigvalsh(gm)[::-1][:4]
    assert np.max(np.abs(r["eigenval"][:4] - w) / w) < 1e-9


def test_bayesian_flag_is_checked_after_the_reduce(shards):
    g, cut, ctxs = shards
    load(ctxs, g, cut)
    D.accumulate_in_process(ctxs, 0, bayesian=False)
    with pytest.raises(S.SNPRelError, match="bayesian"):
        ctxs[0].pca(eigen_cnt=2, bayesian=True)


def test_counters_over_three_uneven_shards(shards):
    g, cut, ctxs = shards
    m = g.shape[0]
    parts = [(0, 128), (128, 5120), (5120, m)]
    three = ctxs + [S.Context(0)]
    try:
        for est, finish, ref in ((EST_IBS, lambda c: np.stack(c.ibs_num()), O.ibs_counts),
                                 (EST_KING_ROBUST, lambda c: c.king_robust_counts(), O.king_robust_counts),
                                 (EST_BETA, lambda c: c.indiv_beta_counts(), O.beta_counts)):
            load(three, g, cut, parts)
            D.accumulate_in_process(three, est)
            r = ref(g)
            for c in three:
                assert np.array_equal(finish(c), r)
        load(three, g, cut, parts)
        D.accumulate_in_process(three, EST_KING_ROBUST)
        ibs0, kin = three[2].king_robust()
        r0, rk = O.king_robust(O.king_robust_counts(g))
        assert np.array_equal(ibs0, r0) and np.array_equal(kin, rk)
    finally:
        three[2].close()
