"""The numpy oracle against the reference's OWN sources (oracle/_ref, built by
oracle/Makefile from /root/reference/src unmodified).  This pins what the
reference's golden vectors leave open: GCTA / EIGMIX / Corr with missing data,
KING with family ids, 1 vs several threads (test_rel.R:114-117).  Skipped when
oracle/_ref has not been built."""
import numpy as np
import pytest

from oracle import ref_lib as R
from oracle import snprel_oracle as O
from snprel_testutil import hapmap_subset

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")


@pytest.fixture(scope="module")
def data():
    return O.synth_geno(257, 1531, seed=12, miss_rate=0.04, maf_lo=0.005)


def test_grm_methods(data):
    w = R.RefWorkspace(data)
    for method, ref in (("GCTA", O.grm_gcta(data)), ("Eigenstrat", O.grm_eigenstrat(data)),
                        ("EIGMIX", O.grm_eigmix(data)), ("Corr", O.grm_corr(data)),
                        ("IndivBeta", O.grm_indivbeta(O.beta_counts(data))[0])):
        one = w.grm(method, 1)
        assert np.max(np.abs(one - ref)) < 1e-12, method
        assert np.max(np.abs(w.grm(method, 4) - one)) < 1e-12      # thread-count independence


def test_integer_estimators(data):
    w = R.RefWorkspace(data)
    assert all(np.array_equal(a, b) for a, b in zip(w.ibs_num(3), O.ibs_counts(data)))
    fam = np.array([(i // 4) if i % 7 else -2147483648 for i in range(data.shape[1])], dtype=np.int32)
    a, b = w.king_robust(2, fam)
    ra, rb = O.king_robust(O.king_robust_counts(data), np.where(fam < 0, -1, fam))
    assert np.array_equal(a, ra) and np.allclose(b, rb, rtol=0, atol=0, equal_nan=True)
    k0, k1 = w.king_homo(2)
    r0, r1 = O.king_homo(data)
    assert np.nanmax(np.abs(k0 - r0)) < 1e-12 and np.nanmax(np.abs(k1 - r1)) < 1e-12
    for inb in (True, False):
        assert np.max(np.abs(w.indiv_beta(2, inb) - O.indiv_beta(O.beta_counts(data), inb)[0])) < 1e-12


def test_plink_mom(data):
    """gnrIBD_PLINK with counts-derived and caller-supplied frequencies, with and
    without the kinship constraint (src/genIBS.cpp:558-639, src/genIBD.cpp:253-383)."""
    w = R.RefWorkspace(data)
    cnt = O.ibs_counts(data)
    _, af0 = O.ibd_mom_tables(data)
    user = af0.copy()
    user[3] = np.nan
    user[7] = 1.5
    for afin in (None, user):
        for kc in (False, True):
            k0, k1, af = w.ibd_mom(2, afin, kc)
            e, raf = O.ibd_mom_tables(data, afin)
            r0, r1 = O.ibd_mom(cnt, e, kc)
            assert np.array_equal(np.isnan(af), np.isnan(raf)) and np.nanmax(np.abs(af - raf)) == 0
            assert np.nanmax(np.abs(k0 - r0)) < 1e-13 and np.nanmax(np.abs(k1 - r1)) < 1e-13


def test_pca_eigmix_and_selection(data):
    w = R.RefWorkspace(data)
    r = w.pca(2, False, 6)
    ref, tr, _ = O.pca_genmat(data)
    assert np.max(np.abs(r["genmat"] - ref)) < 1e-12 and abs(r["TraceXTX"] - tr) < 1e-8 * tr
    ev, evec = O.pca_eigen(ref, 6)
    assert np.max(np.abs(r["eigenval"][:6] - ev)) < 1e-9
    assert np.max(np.abs(w.pca(2, True, 0)["genmat"] - O.pca_genmat(data, bayesian=True)[0])) < 1e-12
    ibd, af = w.eigmix(2, True)
    ri, raf = O.eigmix_ibd(data, True)
    assert np.max(np.abs(ibd - ri)) < 1e-12 and np.max(np.abs(af - raf)) == 0
    sel, nrm = w.select_snp_base(True, 0.03, 0.05)
    assert np.array_equal(sel, O.select_snp_base(data, True, 0.03, 0.05)) and nrm == int((~sel).sum())


def test_reference_reproduces_its_own_goldens(hapmap, goldens):
    g, _ = hapmap_subset(hapmap, 90)
    w = R.RefWorkspace(g)
    assert np.max(np.abs(w.ibs_ave(2) - goldens["ibs"])) == 0
    assert np.max(np.abs(w.pca(2)["genmat"] - goldens["pca_genmat"])) < 1e-12
    assert np.max(np.abs(w.eigmix(2, True)[0] - goldens["eigmix_ibd"])) < 1e-12
    assert np.max(np.abs(w.indiv_beta(2, True) - goldens["beta"])) < 1e-12
    k0, k1, af = w.ibd_mom(2)
    assert np.max(np.abs(k0 - goldens["mom_k0"])) < 1e-14 and np.max(np.abs(k1 - goldens["mom_k1"])) < 1e-14
    assert np.array_equal(af, goldens["mom_afreq"])


def test_loadings_projection_and_correlation(data):
    """gnrPCASNPLoading / gnrPCASampLoading / gnrPCACorr / gnrEigMixSNPLoading / gnrEigMixSampLoading
     with missing data, Bayesian scaling and 1 vs 3 threads."""
    genmat, tr, _ = O.pca_genmat(data)
    val, vec = O.pca_eigen(genmat, 6)
    new = O.synth_geno(61, data.shape[0], seed=99, miss_rate=0.04, maf_lo=0.005)      # samples to project
    w = R.RefWorkspace(data)
    for bayes in (False, True):
        load, avg, scale = O.pca_snp_loading(data, val, vec, tr, bayes)
        rl, ra, rs = w.pca_snp_loading(val, vec, tr, bayes, nthread=1)
        assert np.max(np.abs(rl - load)) < 1e-12 and np.array_equal(ra, avg) and np.max(np.abs(rs - scale)) < 1e-13
        assert np.max(np.abs(w.pca_snp_loading(val, vec, tr, bayes, nthread=3)[0] - rl)) < 1e-13
    corr, rc = O.pca_corr(data, vec[:, :3]), w.pca_corr(vec[:, :3])
    assert np.array_equal(np.isnan(corr), np.isnan(rc)) and np.nanmax(np.abs(corr - rc)) < 1e-12
    sload = load * np.sqrt(((data.shape[1] - 1) / tr) / val[:6])[:, None]
    ibd, af = O.eigmix_ibd(data, diagadj=False)
    ev, evec = O.pca_eigen(ibd, 5)
    el = O.eigmix_snp_loading(data, ev, evec, af)
    assert np.max(np.abs(w.eigmix_snp_loading(ev, evec, af) - el)) < 1e-13
    w2 = R.RefWorkspace(new)
    assert np.max(np.abs(w2.pca_samp_loading(sload, avg, scale) - O.pca_samp_loading(new, sload, avg, scale))) < 1e-11
    esl = el * np.sqrt(1.0 / ev[:5])[:, None]
    assert np.max(np.abs(w2.eigmix_samp_loading(esl, af) - O.eigmix_samp_loading(new, esl, af))) < 1e-12


def test_randomized_pca(data):
    """gnrPCA algorithm "randomized" (CRandomPCA, src/genPCA.cpp:469-796) with num.thread = 1.
    (With several threads the reference re-adds the per-thread partial matrices AuxMat_mc of every
    earlier block -- they are cleared once per iteration, :722-735 -- so its multi-threaded result
    weights the SNP blocks unevenly; the single-threaded result is the algorithm as written.)
    The dominant subspace is well conditioned, the trailing directions of the Krylov basis are not:
    compare the leading singular values / vectors."""
    rng = np.random.default_rng(3)
    aux_dim, it, k = 16, 10, 8
    n = data.shape[1]
    aux = rng.standard_normal(aux_dim * n)
    sig, vt, tr = O.pca_randomized(data, aux, aux_dim, it)
    w = R.RefWorkspace(data)
    rsig, rvt, rtr = w.pca_randomized(aux, aux_dim, it, nthread=1)
    assert abs(tr - rtr) <= 1e-12 * rtr
    assert rvt.shape == (aux_dim * (it + 1), n)
    assert np.max(np.abs(sig[:k] - rsig[:k]) / rsig[:k]) < 1e-8
    assert np.max(1 - np.abs(np.sum(vt[:k] * rvt[:k], axis=1))) < 1e-10
    # and it approximates the exact decomposition (R/PCA.R:80-89)
    res = O.pca_randomized_result(sig, vt, tr, n, k)
    ev, evec = O.pca_eigen(O.pca_genmat(data)[0], k)
    assert np.max(np.abs(res["eigenval"][:k] - ev[:k]) / ev[:k]) < 1e-3


def test_gds_output_stream_is_the_full_matrix_row_by_row(data):
    """gnrGRM with out.fn (R/IBD.R:570-594): grm_save_to_gds appends n full rows to the "grm" node
    (src/genPCA.cpp:1571-1584).  (method "Corr" never reaches the node in the reference, :1655-1685.)"""
    w = R.RefWorkspace(data)
    for method in ("GCTA", "Eigenstrat", "EIGMIX", "IndivBeta"):
        assert np.array_equal(w.grm_gds(method, 2), w.grm(method, 2)), method
