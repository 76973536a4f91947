"""Pin the oracle (oracle/snprel_oracle.py) against the reference's own golden
vectors (inst/unitTests/test_rel.R:103-339, goldens decoded by
oracle/make_golden.py).  CPU only."""
import numpy as np

from oracle import snprel_oracle as O
from snprel_testutil import hapmap_subset


def test_selection_counts(hapmap, goldens):
    g90, idx90 = hapmap_subset(hapmap, 90)
    g60, idx60 = hapmap_subset(hapmap, 60)
    assert g90.shape == (8695, 90) and g60.shape == (8658, 60)
    # remove.monosnp=TRUE => returned snp.id == 1-based kept indices
    assert np.array_equal(idx90 + 1, goldens["ibs_snp_id"])
    assert np.array_equal(idx60 + 1, goldens["king_snp_id"])
    g279, _ = hapmap_subset(hapmap, 279, missing_rate=0.01)
    assert g279.shape == (8039, 279)      # BASELINE config 1 working set


def test_ibs_golden(hapmap, goldens):
    g, _ = hapmap_subset(hapmap, 90)
    c = O.ibs_counts(g)
    assert np.array_equal(c, O.ibs_counts_packed(g))     # bit-plane restatement
    assert np.max(np.abs(O.ibs_ave(c) - goldens["ibs"])) == 0.0


def test_plink_mom_golden(hapmap, goldens):          # test_rel.R:197-227
    g, _ = hapmap_subset(hapmap, 90)
    e, af = O.ibd_mom_tables(g)
    k0, k1 = O.ibd_mom(O.ibs_counts(g), e)
    assert np.array_equal(af, goldens["mom_afreq"])
    assert np.max(np.abs(k0 - goldens["mom_k0"])) < 1e-14
    assert np.max(np.abs(k1 - goldens["mom_k1"])) < 1e-14


def test_pca_genmat_golden(hapmap, goldens):
    g, _ = hapmap_subset(hapmap, 90)
    genmat, _, _ = O.pca_genmat(g)
    assert np.max(np.abs(genmat - goldens["pca_genmat"])) < 1e-13


def test_pca_sample_loading_sign_free(hapmap, goldens):
    """Validate.PCA.RData$samploading = eigenvectors of the first 100 of 279
    samples?  No: it is snpgdsPCASampLoading on samp.id[1:100]; only the genmat
    is a hot-path pin.  Here: eigen-decomposition sanity against LAPACK."""
    g, _ = hapmap_subset(hapmap, 90)
    genmat, _, _ = O.pca_genmat(g)
    val, vec = O.pca_eigen(genmat, 8)
    assert np.all(np.diff(val) <= 0)
    assert np.allclose(genmat @ vec, vec * val[None, :], atol=1e-9)


def test_eigmix_golden(hapmap, goldens):
    g, _ = hapmap_subset(hapmap, 90)
    ibd, _ = O.eigmix_ibd(g, diagadj=True)
    assert np.max(np.abs(ibd - goldens["eigmix_ibd"])) < 1e-13


def test_king_robust_golden(hapmap, goldens):
    g, _ = hapmap_subset(hapmap, 60)
    ibs0, kin = O.king_robust(O.king_robust_counts(g))
    assert np.max(np.abs(ibs0 - goldens["king_robust_ibs0"])) == 0.0
    assert np.max(np.abs(kin - goldens["king_robust_kinship"])) == 0.0


def test_king_homo_golden(hapmap, goldens):
    g, _ = hapmap_subset(hapmap, 60)
    k0, k1 = O.king_homo(g)
    assert np.nanmax(np.abs(k0 - goldens["king_homo_k0"])) < 1e-12
    assert np.nanmax(np.abs(k1 - goldens["king_homo_k1"])) < 1e-12


def test_indiv_beta_golden(hapmap, goldens):
    g, _ = hapmap_subset(hapmap, 90)
    beta, _ = O.indiv_beta(O.beta_counts(g), inbreeding=True)
    assert np.max(np.abs(beta - goldens["beta"])) < 1e-13


def test_gcta_merge_identity(hapmap):
    """inst/unitTests/test_GRM.R:15-49: GCTA GRM on SNPs without missing data
    split into 3 subsets and merged (weights = #SNPs) equals the GRM on all."""
    g, _ = hapmap_subset(hapmap, 279, missing_rate=0.0)
    g = g[:3000]
    parts = [g[0::3], g[1::3], g[2::3]]
    merged = O.merge_grm([O.grm_gcta(p) for p in parts], [p.shape[0] for p in parts])
    assert np.max(np.abs(merged - O.grm_gcta(g))) < 1e-12


def test_synth_reproducible():
    a = O.synth_geno(64, 200, seed=7)
    b = O.synth_geno(64, 120, seed=7, snp_start=80)
    assert np.array_equal(a[80:], b)
    assert set(np.unique(a)) <= {0, 1, 2, 3}
    frac_missing = (O.synth_geno(256, 2000, seed=3) == 3).mean()
    assert 0.003 < frac_missing < 0.007


def _sign_free_rounded(got, gold, digits):
    """the goldens are round(x, digits) of vectors whose sign LAPACK leaves free"""
    tol = 0.5 * 10.0 ** (-digits) + 1e-9
    a = np.nanmax(np.abs(got - gold))
    b = np.nanmax(np.abs(-got - gold))
    return min(a, b) <= tol and np.array_equal(np.isnan(got), np.isnan(gold))


def test_pca_loadings_and_corr_golden(hapmap, goldens):          # test_rel.R:20-41,144-160
    g, idx = hapmap_subset(hapmap, 90)
    genmat, tr, _ = O.pca_genmat(g)
    val, vec = O.pca_eigen(genmat, 8)
    load, avg, scale = O.pca_snp_loading(g, val, vec, tr)
    for k in range(8):
        assert _sign_free_rounded(load[k], goldens["pca_snploading"][k], 3), k
    g100 = np.ascontiguousarray(hapmap["geno"][:, :100][idx])     # snpgdsPCASampLoading(sample.id = first 100)
    sload = load * np.sqrt(((90 - 1) / tr) / val[:8])[:, None]    # R/PCA.R:274-277
    sl = O.pca_samp_loading(g100, sload, avg, scale)
    for k in range(8):
        assert _sign_free_rounded(sl[:, k], goldens["pca_samploading"][:, k], 4), k
    corr = O.pca_corr(np.ascontiguousarray(hapmap["geno"][:, :90]), vec[:, :2])    # snp.id = NULL: all 9088 SNPs
    for k in range(2):
        assert _sign_free_rounded(corr[k], goldens["pca_corr"][k], 3), k
