"""The R binding itself: snprelate_b200/csrc/r_shim.cpp (drop-in bodies of the reference's .Call
entry points gnrGRM, gnrPCA, gnrEigMix, gnrIBSAve, gnrIBSNum, gnrIBD_PLINK, gnrIBD_KING_*, gnrIBD_Beta,
gnrPCACorr, gnrPCASNPLoading, gnrPCASampLoading, gnrEigMix*Loading) compiled against the reference's
OWN workspace layer (dGenGWAS.h / dGenGWAS.cpp, unmodified) and linked to libsnprel_b200.so
(oracle/Makefile -> oracle/_ref/libsnprelate_b200_rshim.so).  The test drives it the way R does --
gnrSetGenoSpace, gnrSelSNP_Base, then the estimator with scalars only -- and compares the R objects
it returns with the oracle.  Skipped where the library has not been built (it needs the reference's
headers, so it is built in the development container and shipped with the snapshot)."""
import numpy as np
import pytest

from oracle import ref_lib as R
from oracle import snprel_oracle as O

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not R.rshim_available(), reason="oracle/_ref/libsnprelate_b200_rshim.so not built")]

TOL = 1e-10


def relerr(got, ref):
    return float(np.nanmax(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)))


@pytest.fixture(scope="module")
def data():
    return O.synth_geno(211, 1777, seed=31, miss_rate=0.03, maf_lo=0.01)


def test_estimators_through_the_r_entry_points(data):
    w = R.RefWorkspace(data, R.RSHIM_PATH)          # gnrSetGenoSpace on the reference's workspace
    assert w.dims() == (data.shape[0], data.shape[1])
    assert relerr(w.grm("GCTA"), O.grm_gcta(data)) < TOL                        # Rf_allocMatrix(REALSXP, n, n)
    assert relerr(w.grm("IndivBeta"), O.grm_indivbeta(O.beta_counts(data))[0]) < 1e-12
    i0, i1, i2 = w.ibs_num()
    assert np.array_equal(np.stack([i0, i1, i2]), O.ibs_counts(data))           # three INTSXP matrices, bit exact
    idx = np.arange(data.shape[1])
    fam = np.where(idx % 7 == 0, -2147483648, idx // 3).astype(np.int32)        # NA_INTEGER for every 7th sample
    ibs0, kin = w.king_robust(family=fam)
    r0, rk = O.king_robust(O.king_robust_counts(data), fam)
    assert np.array_equal(np.isnan(kin), np.isnan(rk)) and relerr(kin, rk) < 1e-14 and relerr(ibs0, r0) < 1e-14
    m0, m1, maf = w.ibd_mom()
    e, raf = O.ibd_mom_tables(data)
    o0, o1 = O.ibd_mom(O.ibs_counts(data), e)
    assert relerr(m0, o0) < 1e-12 and relerr(m1, o1) < 1e-12 and np.allclose(maf, raf, equal_nan=True)


def test_pca_eigmix_and_loadings_through_the_r_entry_points(data):
    w = R.RefWorkspace(data, R.RSHIM_PATH)
    r = w.pca(bayesian=False, eigen_cnt=5)                                      # list(5) of gnrPCA
    genmat, tr, _ = O.pca_genmat(data)
    assert relerr(r["genmat"], genmat) < TOL and abs(r["TraceXTX"] - tr) < 1e-9 * tr
    val, vec = O.pca_eigen(genmat, 5)
    assert np.max(np.abs(r["eigenval"][:5] - val)) < 1e-9 and np.all(np.isnan(r["eigenval"][5:]))
    for k in range(3):
        d = min(np.max(np.abs(r["eigenvect"][:, k] - vec[:, k])), np.max(np.abs(r["eigenvect"][:, k] + vec[:, k])))
        assert d < 1e-6
    ibd, af = w.eigmix(diagadj=True)                                            # list(4) of gnrEigMix
    oibd, oaf = O.eigmix_ibd(data, diagadj=True)
    assert relerr(ibd, oibd) < TOL and np.allclose(af, oaf)
    load, avg, scale = w.pca_snp_loading(val, vec, tr)                          # list(3) of gnrPCASNPLoading
    ol, oavg, oscale = O.pca_snp_loading(data, val, vec, tr)
    assert relerr(load, ol) < TOL and np.array_equal(avg, oavg) and relerr(scale, oscale) < 1e-14
    new = O.synth_geno(57, data.shape[0], seed=77, miss_rate=0.03, maf_lo=0.01)
    sload = ol * np.sqrt(((data.shape[1] - 1) / tr) / val[:5])[:, None]
    w2 = R.RefWorkspace(new, R.RSHIM_PATH)
    assert relerr(w2.pca_samp_loading(sload, oavg, oscale), O.pca_samp_loading(new, sload, oavg, oscale)) < TOL


def test_workspace_selection_is_honoured(data):
    """gnrSelSNP_Base runs in the reference's own workspace code; the shim must read the SELECTED
    genotypes through CdBaseWorkSpace::snpRead (the contract of SURVEY.md section 8b)."""
    w = R.RefWorkspace(data, R.RSHIM_PATH)
    sel, nrm = w.select_snp_base(remove_mono=True, maf=0.05, missrate=0.05)
    keep = O.select_snp_base(data, True, 0.05, 0.05)
    assert np.array_equal(sel, keep) and nrm == int((~keep).sum()) and w.dims()[0] == int(keep.sum())
    sub = np.ascontiguousarray(data[keep])
    assert relerr(w.grm("GCTA"), O.grm_gcta(sub)) < TOL
    assert np.array_equal(np.stack(w.ibs_num()), O.ibs_counts(sub))


def test_randomized_pca_through_the_r_entry_point(data):
    """gnrPCA(algorithm = "randomized") of the binding: list(sigma, V^T [hsize x n], 2 TraceXTX)."""
    n = data.shape[1]
    aux_dim, it = 8, 6
    aux = np.random.default_rng(5).standard_normal(aux_dim * n)
    w = R.RefWorkspace(data, R.RSHIM_PATH)
    sig, vt, tr = w.pca_randomized(aux, aux_dim, it)
    rsig, rvt, rtr = O.pca_randomized(data, aux, aux_dim, it)
    assert abs(tr - rtr) <= 1e-12 * rtr and vt.shape == (aux_dim * (it + 1), n)
    assert np.max(np.abs(sig[:4] - rsig[:4]) / rsig[:4]) < 1e-7
    assert np.max(1 - np.abs(np.sum(vt[:4] * rvt[:4], axis=1))) < 1e-10


def test_gds_output_through_the_r_entry_point(data):
    """gnrGRM(..., GDS = <output file>) of the binding: the same n x n row stream the reference appends
    to the "grm" node of a SNPRELATE_OUTPUT file (grm_save_to_gds, src/genPCA.cpp:1571-1584)."""
    w = R.RefWorkspace(data, R.RSHIM_PATH)
    for method, ref in (("GCTA", O.grm_gcta(data)), ("EIGMIX", O.grm_eigmix(data)),
                        ("IndivBeta", O.grm_indivbeta(O.beta_counts(data))[0])):
        got = w.grm_gds(method)
        assert np.array_equal(got, got.T) and relerr(got, ref) < TOL, method
        assert np.array_equal(got, w.grm(method)), method          # identical to the in-memory result
