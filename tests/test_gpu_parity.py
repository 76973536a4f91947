"""Parity of the CUDA path (through the C ABI) against the oracle and the
reference's golden vectors.  Mirrors inst/unitTests/test_rel.R and test_GRM.R.

Tolerances: integer counters bit-exact; GRM-type entries
|diff| <= 1e-10 * max(|ref|, 1); eigenvectors 1e-6 up to sign."""
import numpy as np
import pytest

import snprelate_b200 as S
from oracle import snprel_oracle as O
from snprel_testutil import hapmap_subset

pytestmark = pytest.mark.gpu

TOL = 1e-10


def relerr(got, ref):
    return float(np.nanmax(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)))


def load(ctx, g):
    ctx.geno_begin(g.shape[1], g.shape[0])
    ctx.geno_push_u8(g)


@pytest.fixture(scope="module")
def ctx():
    c = S.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def gds(hapmap):
    return S.GenotypeData(hapmap["geno2b"], sample_id=hapmap["sample_id"], snp_id=hapmap["snp_id"],
                          chromosome=hapmap["chromosome"], packed_2bit=True, n_samp=hapmap["nsamp"])


# ---------------------------------------------------------------- goldens (test_rel.R)

def test_ibs_golden(gds, hapmap, goldens):          # test_rel.R:103-129
    samp = hapmap["sample_id"][:90]
    r = S.snpgdsIBS(gds, sample_id=samp, missing_rate=float("nan"))
    assert np.array_equal(r["snp.id"], goldens["ibs_snp_id"])
    assert np.max(np.abs(r["ibs"] - goldens["ibs"])) == 0.0
    rp = S.snpgdsIBS(gds, sample_id=samp, missing_rate=float("nan"), useMatrix=True)
    assert np.array_equal(rp["ibs"]["x"], O.to_packed_upper(r["ibs"]))


def test_plink_mom_golden(gds, hapmap, goldens):    # test_rel.R:197-227
    samp = hapmap["sample_id"][:90]
    r = S.snpgdsIBDMoM(gds, sample_id=samp, missing_rate=float("nan"), kinship=True)
    assert np.array_equal(r["afreq"], goldens["mom_afreq"])
    assert np.max(np.abs(r["k0"] - goldens["mom_k0"])) < 1e-14
    assert np.max(np.abs(r["k1"] - goldens["mom_k1"])) < 1e-14
    assert np.array_equal(r["kinship"], 0.5 * (1 - r["k0"] - r["k1"]) + 0.25 * r["k1"])
    rp = S.snpgdsIBDMoM(gds, sample_id=samp, missing_rate=float("nan"), useMatrix=True)
    assert np.array_equal(rp["k0"]["x"], O.to_packed_upper(r["k0"]))
    assert np.array_equal(rp["k1"]["x"], O.to_packed_upper(r["k1"]))


def test_plink_mom_vs_oracle(ctx, gds, hapmap):
    """Counts-derived and caller-supplied allele frequencies, kinship constraint, row windows."""
    g = O.synth_geno(300, 2500, seed=23, miss_rate=0.04, maf_lo=0.01)
    load(ctx, g)
    cnt = O.ibs_counts(g)
    _, af0 = O.ibd_mom_tables(g)
    user = af0.copy()
    user[5] = np.nan
    user[11] = -0.25
    for afin in (None, user):
        for kc in (False, True):
            k0, k1, af = ctx.ibd_mom(afin, kc)
            e, raf = O.ibd_mom_tables(g, afin)
            r0, r1 = O.ibd_mom(cnt, e, kc)
            assert np.array_equal(af, raf, equal_nan=True)
            assert np.nanmax(np.abs(k0 - r0)) < 1e-13 and np.nanmax(np.abs(k1 - r1)) < 1e-13
            assert np.array_equal(np.isnan(k0), np.isnan(r0))
    sums, _ = ctx.ibd_mom_sums()
    full = ctx.ibd_mom_from_sums(sums, True, packed=True)
    win = ctx.packed_by_windows(lambda: ctx.ibd_mom_from_sums(sums, True, packed=True), 256)
    assert np.array_equal(win[0], full[0]) and np.array_equal(win[1], full[1])
    # allele.freq through the R-level wrapper: selection uses the supplied frequencies
    samp = hapmap["sample_id"][:60]
    base = S.snpgdsIBDMoM(gds, sample_id=samp, missing_rate=float("nan"))
    af_all = np.full(gds.n_snp, np.nan)
    af_all[np.isin(gds.snp_id, base["snp.id"])] = base["afreq"]
    r = S.snpgdsIBDMoM(gds, sample_id=samp, missing_rate=float("nan"), allele_freq=af_all)
    assert np.array_equal(r["snp.id"], base["snp.id"]) and np.array_equal(r["afreq"], base["afreq"])
    gsel, _ = hapmap_subset(hapmap, 60)
    e, _ = O.ibd_mom_tables(gsel, base["afreq"])
    r0, r1 = O.ibd_mom(O.ibs_counts(gsel), e)
    assert np.max(np.abs(r["k0"] - r0)) < 1e-13 and np.max(np.abs(r["k1"] - r1)) < 1e-13
    with pytest.raises(S.SNPRelError, match="allele.freq"):
        S.snpgdsIBDMoM(gds, sample_id=samp, allele_freq=np.zeros(5))


def test_pca_golden(gds, hapmap, goldens):          # test_rel.R:133-195
    samp = hapmap["sample_id"][:90]
    r = S.snpgdsPCA(gds, sample_id=samp, missing_rate=float("nan"), need_genmat=True, eigen_cnt=8)
    assert relerr(r["genmat"], goldens["pca_genmat"]) < TOL
    val, vec = O.pca_eigen(goldens["pca_genmat"], 8)
    assert np.max(np.abs(r["eigenval"][:8] - val)) < 1e-8
    assert np.all(np.isnan(r["eigenval"][8:]))
    for k in range(4):          # well separated leading PCs, up to sign
        d = min(np.max(np.abs(r["eigenvect"][:, k] - vec[:, k])), np.max(np.abs(r["eigenvect"][:, k] + vec[:, k])))
        assert d < 1e-6
    assert abs(r["varprop"][0] - r["eigenval"][0] / np.trace(r["genmat"])) < 1e-12


def _sign_free(got, ref, tol):
    return min(np.nanmax(np.abs(got - ref)), np.nanmax(np.abs(-got - ref))) <= tol and \
        np.array_equal(np.isnan(got), np.isnan(ref))


def test_pca_loadings_golden(gds, hapmap, goldens):          # test_rel.R:144-160
    samp = hapmap["sample_id"]
    pca = S.snpgdsPCA(gds, sample_id=samp[:90], missing_rate=float("nan"), need_genmat=True, eigen_cnt=8)
    corr = S.snpgdsPCACorr(pca, gds, eig_which=[1, 2])["snpcorr"]
    assert corr.shape == (2, 9088)
    for k in range(2):              # goldens are rounded (3 / 3 / 4 decimals), eigenvector signs are free
        assert _sign_free(corr[k], goldens["pca_corr"][k], 0.5e-3 + 1e-9), k
    sl = S.snpgdsPCASNPLoading(pca, gds)
    assert sl["snploading"].shape == (8, 8695)
    for k in range(8):
        assert _sign_free(sl["snploading"][k], goldens["pca_snploading"][k], 0.5e-3 + 1e-9), k
    pr = S.snpgdsPCASampLoading(sl, gds, sample_id=samp[:100])
    assert pr["eigenvect"].shape == (100, 8) and np.all(np.isnan(pr["eigenval"]))
    for k in range(8):
        assert _sign_free(pr["eigenvect"][:, k], goldens["pca_samploading"][:, k], 0.5e-4 + 1e-9), k
    # and to float64 accuracy against the oracle fed with the device's own eigenpairs
    g, idx = hapmap_subset(hapmap, 90)
    load, avg, scale = O.pca_snp_loading(g, pca["eigenval"], pca["eigenvect"], pca["TraceXTX"])
    assert relerr(sl["snploading"], load) < TOL and np.array_equal(sl["avgfreq"], avg) and relerr(sl["scale"], scale) < 1e-14
    # projecting the training samples reproduces their eigenvectors (the identity the method rests on)
    back = S.snpgdsPCASampLoading(sl, gds, sample_id=samp[:90])["eigenvect"]
    assert np.max(np.abs(back - pca["eigenvect"])) < 1e-8


@pytest.mark.parametrize("n,m,k,miss", [(300, 3001, 5, 0.03), (257, 1200, 40, 0.0), (130, 77, 1, 0.2)])
def test_loadings_vs_oracle(ctx, n, m, k, miss):
    """SNP loadings, projection of new samples, SNP-PC correlation and the EIGMIX flavours against
    the oracle on ragged sizes (k > 32 takes two column panels), with missing data."""
    g = O.synth_geno(n, m, seed=n + m, miss_rate=miss, maf_lo=0.01)
    new = O.synth_geno(91, m, seed=5, miss_rate=miss, maf_lo=0.01)
    genmat, tr, _ = O.pca_genmat(g)
    val, vec = O.pca_eigen(genmat, k)
    load(ctx, g)
    for bayes in (False, True):
        got, avg, scale = ctx.pca_snp_loading(val, vec, tr, bayes)
        ref, ravg, rscale = O.pca_snp_loading(g, val, vec, tr, bayes)
        assert relerr(got, ref) < TOL and np.array_equal(avg, ravg) and relerr(scale, rscale) < 1e-14
    corr, rcorr = ctx.pca_corr(vec), O.pca_corr(g, vec)
    assert np.array_equal(np.isnan(corr), np.isnan(rcorr)) and np.nanmax(np.abs(corr - rcorr)) < 1e-9
    ibd, af = O.eigmix_ibd(g, diagadj=False)
    ev, evec = O.pca_eigen(ibd, k)
    el = ctx.eigmix_snp_loading(ev, evec, af)
    assert relerr(el, O.eigmix_snp_loading(g, ev, evec, af)) < TOL
    sload = ref * np.sqrt(((n - 1) / tr) / val[:k])[:, None]
    esl = el * np.sqrt(1.0 / np.abs(ev[:k]))[:, None]
    load(ctx, new)
    assert relerr(ctx.pca_samp_loading(sload, ravg, rscale), O.pca_samp_loading(new, sload, ravg, rscale)) < TOL
    assert relerr(ctx.eigmix_samp_loading(esl, af), O.eigmix_samp_loading(new, esl, af)) < TOL
    with pytest.raises(S.SNPRelError, match="number of samples"):
        ctx.pca_corr(vec)           # 91 samples in the workspace, n rows in vec


def _two_populations(n, m, seed):
    """genotypes of two diverged populations (a few dominant eigenvalues on top of the noise bulk)"""
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


@pytest.mark.parametrize("structured", [True, False])
def test_filtered_subspace_eigen_solver(ctx, structured):
    """The Chebyshev-filtered subspace iteration of the eigen step (csrc/eigen.cu, n >= 2048) against
    the dense cuSOLVER decomposition and the oracle: eigenvalues to 1e-9, eigenvectors to 1e-6 up to
    sign (BASELINE tolerance) -- with population structure and on a structure-free matrix whose
    wanted eigenvalues sit at the edge of the noise bulk."""
    n, m, k = 2304, 6000, 16
    g = _two_populations(n, m, 11) if structured else O.synth_geno(n, m, seed=21, miss_rate=0.002)
    load(ctx, g)
    r1 = ctx.pca(eigen_cnt=k, need_genmat=True)
    solver, rounds, gemms = ctx.last_eigen_info()
    assert solver == 1 and rounds >= 1 and gemms > 0, (solver, rounds, gemms)
    ctx.debug_flags(4)                       # dense solver
    try:
        r0 = ctx.pca(eigen_cnt=k)
        assert ctx.last_eigen_info()[0] == 0
    finally:
        ctx.debug_flags(0)
    val, vec = O.pca_eigen(r1["genmat"], k)
    assert np.max(np.abs(r1["eigenval"][:k] - val)) < 1e-9 * val[0]
    assert np.max(np.abs(r0["eigenval"][:k] - val)) < 1e-9 * val[0]
    assert np.all(np.isnan(r1["eigenval"][k:]))
    for v1 in (r1["eigenvect"], r0["eigenvect"]):
        # the defining property, independent of how close neighbouring eigenvalues are
        assert np.max(np.abs(r1["genmat"] @ v1 - v1 * val[None, :])) < 1e-9 * val[0]
        assert np.max(np.abs(v1.T @ v1 - np.eye(k))) < 1e-10
    gaps = np.minimum(np.abs(np.diff(val, prepend=np.inf)), np.abs(np.diff(val, append=val[-1] - 1)))[: k - 1]
    for i in np.nonzero(gaps > 1e-3 * val[0])[0]:      # isolated eigenvalues: the vectors themselves agree
        d = min(np.max(np.abs(r1["eigenvect"][:, i] - vec[:, i])), np.max(np.abs(r1["eigenvect"][:, i] + vec[:, i])))
        assert d < 1e-6, (i, d)


def test_king_golden(gds, hapmap, goldens):         # test_rel.R:237-283
    samp = hapmap["sample_id"][:60]
    r = S.snpgdsIBDKING(gds, sample_id=samp, missing_rate=float("nan"), type="KING-robust")
    assert np.array_equal(r["snp.id"], goldens["king_snp_id"])
    assert np.max(np.abs(r["IBS0"] - goldens["king_robust_ibs0"])) == 0.0
    assert np.max(np.abs(r["kinship"] - goldens["king_robust_kinship"])) == 0.0
    h = S.snpgdsIBDKING(gds, sample_id=samp, missing_rate=float("nan"), type="KING-homo")
    assert relerr(h["k0"], goldens["king_homo_k0"]) < TOL
    assert relerr(h["k1"], goldens["king_homo_k1"]) < TOL


def test_indiv_beta_golden(gds, hapmap, goldens):   # test_rel.R:287-313
    samp = hapmap["sample_id"][:90]
    r = S.snpgdsIndivBeta(gds, sample_id=samp, missing_rate=float("nan"))
    assert relerr(r["beta"], goldens["beta"]) < TOL


def test_eigmix_golden(gds, hapmap, goldens):       # test_rel.R:317-339
    samp = hapmap["sample_id"][:90]
    r = S.snpgdsEIGMIX(gds, sample_id=samp, missing_rate=float("nan"), ibdmat=True)
    assert relerr(r["ibd"], goldens["eigmix_ibd"]) < TOL


def test_hapmap_config1_all_methods(gds, hapmap):
    """BASELINE config 1: 279 samples, default filters -> 279 x 8039."""
    g, _ = hapmap_subset(hapmap, 279, missing_rate=0.01)
    for method, ref in (("GCTA", O.grm_gcta(g)), ("Eigenstrat", O.grm_eigenstrat(g)),
                        ("EIGMIX", O.grm_eigmix(g)), ("Corr", O.grm_corr(g)),
                        ("IndivBeta", O.grm_indivbeta(O.beta_counts(g))[0])):
        r = S.snpgdsGRM(gds, method=method)
        assert r["grm"].shape == (279, 279) and len(r["snp.id"]) == 8039
        assert relerr(r["grm"], ref) < TOL, method
    w = S.snpgdsGRM(gds, method="Weighted")
    assert w["method"] == "EIGMIX"        # the mapped name, R/IBD.R:552-555,603-604
    assert relerr(w["grm"], O.grm_eigmix(g)) < TOL


def test_grm_merge_identity(gds, hapmap):           # test_GRM.R:15-49
    g, idx = hapmap_subset(hapmap, 279, missing_rate=0.0)
    ids = hapmap["snp_id"][idx]
    parts = [ids[0::3], ids[1::3], ids[2::3]]
    grms = [S.snpgdsGRM(gds, snp_id=p, missing_rate=0.0, method="GCTA")["grm"] for p in parts]
    merged = O.merge_grm(grms, [len(p) for p in parts])
    full = S.snpgdsGRM(gds, snp_id=ids, missing_rate=0.0, method="GCTA")["grm"]
    assert relerr(merged, full) < TOL


def test_grm_files_merge(gds, hapmap, tmp_path):     # test_GRM.R:15-90 (GCTA and IndivBeta)
    """snpgdsGRM(out.fn=) streams the matrix to a SNPRELATE_OUTPUT container (in row bands for
    the windowed estimators) and snpgdsMergeGRM of three SNP subsets reproduces the GRM of
    the union."""
    from snprelate_b200.grmfile import GrmFile
    _, idx = hapmap_subset(hapmap, 279, missing_rate=0.0)
    ids = hapmap["snp_id"][idx]
    parts = [ids[:1000], ids[1000:3000], ids[3000:]]
    for method in ("GCTA", "IndivBeta"):
        names = [str(tmp_path / f"{method}{k}.grm") for k in range(3)]
        for nm, p in zip(names, parts):
            assert S.snpgdsGRM(gds, snp_id=p, method=method, out_fn=nm, window_rows=256) is None
        out = str(tmp_path / f"{method}.grm")
        S.snpgdsMergeGRM(names, out)
        full = S.snpgdsGRM(gds, snp_id=ids, method=method)
        f = GrmFile(out)
        assert relerr(np.array(f.grm), full["grm"]) < TOL, method
        assert np.array_equal(f.snp_id, ids) and np.array_equal(f.sample_id, hapmap["sample_id"])
        one = GrmFile(names[0])
        direct = S.snpgdsGRM(gds, snp_id=parts[0], method=method)
        assert np.array_equal(np.array(one.grm), direct["grm"])      # band writes == direct matrix
        if method == "IndivBeta":
            assert one.avg_val == direct["avg_val"]
            assert abs(f.avg_val - full["avg_val"]) < 1e-12
    S.snpgdsGRM(gds, snp_id=parts[0], method="GCTA", out_fn=str(tmp_path / "s.grm"), out_prec="single")
    assert GrmFile(str(tmp_path / "s.grm")).grm.dtype == np.float32


# ---------------------------------------------------------------- synthetic vs oracle

@pytest.mark.parametrize("n,m,miss", [(300, 1000, 0.02), (513, 4099, 0.0), (70, 130, 0.3)])
def test_tensor_count_engine_bit_identical(n, m, miss):
    """The opt-in tensor-pipe engine (snprel_set_count_engine) reproduces the packed-bit kernels'
    uint32 counters exactly, hence every estimator built on them, also inside row windows."""
    g = O.synth_geno(n, m, seed=n, miss_rate=miss, maf_lo=0.01)
    with S.Context(0) as c:
        load(c, g)
        c.set_count_engine("tensor")
        assert np.array_equal(np.stack(c.ibs_num()), O.ibs_counts(g))
        assert np.array_equal(c.king_robust_counts(), O.king_robust_counts(g))
        tb = c.indiv_beta_counts()
        ob = O.beta_counts(g)
        assert np.array_equal(tb, ob)
        t_ibs, t_king, t_beta = c.ibs_ave(packed=True), c.king_robust(packed=True), c.indiv_beta()[0]
        # KING-homo keeps two float-sum Gram planes alive while the counters run (the
        # tensor engine used to overwrite them)
        k0, k1 = c.king_homo()
        rk0, rk1 = O.king_homo(g)
        assert relerr(k0, rk0) < TOL and relerr(k1, rk1) < TOL
        if n > 256:
            win = c.packed_by_windows(lambda: c.king_robust(packed=True), 256)
            assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(win, t_king))
        c.set_count_engine("bits")
        assert np.array_equal(c.ibs_ave(packed=True), t_ibs, equal_nan=True)
        assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(c.king_robust(packed=True), t_king))
        assert np.array_equal(c.indiv_beta()[0], t_beta, equal_nan=True)


@pytest.mark.parametrize("n,m,miss", [(4, 1, 0.0), (17, 33, 0.1), (130, 129, 0.0), (257, 1000, 0.02),
                                      (513, 777, 0.3), (1000, 4099, 0.005)])
def test_counts_bit_exact(ctx, n, m, miss):
    g = O.synth_geno(n, m, seed=n * 7 + m, miss_rate=miss)
    load(ctx, g)
    assert np.array_equal(np.stack(ctx.ibs_num()), O.ibs_counts(g))
    assert np.array_equal(ctx.king_robust_counts(), O.king_robust_counts(g))
    assert np.array_equal(ctx.indiv_beta_counts(), O.beta_counts(g))


@pytest.mark.parametrize("n,m,miss", [(5, 3, 0.0), (130, 129, 0.0), (300, 2000, 0.02), (513, 777, 0.3),
                                      (1000, 4099, 0.005)])
def test_grm_family_vs_oracle(ctx, n, m, miss):
    g = O.synth_geno(n, m, seed=n + m, miss_rate=miss, maf_lo=0.01)
    load(ctx, g)
    assert relerr(ctx.grm("GCTA")[0], O.grm_gcta(g)) < TOL
    assert relerr(ctx.grm("Eigenstrat")[0], O.grm_eigenstrat(g)) < TOL
    assert relerr(ctx.grm("EIGMIX")[0], O.grm_eigmix(g)) < TOL
    assert relerr(ctx.grm("Corr")[0], O.grm_corr(g)) < TOL
    ibd = ctx.eigmix(eigen_cnt=0, diagadj=True, ibdmat=True)
    ref_ibd, ref_af = O.eigmix_ibd(g, diagadj=True)
    assert relerr(ibd["ibd"], ref_ibd) < TOL
    assert np.max(np.abs(ibd["afreq"] - ref_af)) < 1e-15
    r = ctx.pca(eigen_cnt=2, bayesian=True, need_genmat=True)
    assert relerr(r["genmat"], O.pca_genmat(g, bayesian=True)[0]) < TOL
    full, _ = ctx.grm("GCTA")
    packed, _ = ctx.grm("GCTA", packed=True)
    assert np.array_equal(packed, O.to_packed_upper(full))


def test_estimator_epilogues_vs_oracle(ctx):
    g = O.synth_geno(200, 1500, seed=9, miss_rate=0.05)
    load(ctx, g)
    fam = np.array([(i // 3) if i % 5 else S._lib.NA_INT for i in range(200)], dtype=np.int32)
    ibs0, kin = ctx.king_robust(fam)
    r0, rk = O.king_robust(O.king_robust_counts(g), np.where(fam == S._lib.NA_INT, -1, fam))
    assert np.array_equal(ibs0, r0) and np.allclose(kin, rk, rtol=0, atol=0, equal_nan=True)
    assert np.array_equal(ctx.ibs_ave(), O.ibs_ave(O.ibs_counts(g)))
    for inb in (True, False):
        beta, avg = ctx.indiv_beta(inb)
        rb, ravg = O.indiv_beta(O.beta_counts(g), inb)
        assert relerr(beta, rb) < TOL and abs(avg - ravg) < 1e-13
    k0, k1 = ctx.king_homo()
    rk0, rk1 = O.king_homo(g)
    assert relerr(k0, rk0) < TOL and relerr(k1, rk1) < TOL
    grm, avg = ctx.grm("IndivBeta")
    rg, ravg = O.grm_indivbeta(O.beta_counts(g))
    assert relerr(grm, rg) < TOL and abs(avg - ravg) < 1e-13


def test_table_gram_exact(ctx):
    """The tcgen05 kernel alone: exact int64 Gram with random int8 tables."""
    rng = np.random.default_rng(3)
    n, m = 700, 3001
    g = O.synth_geno(n, m, seed=21, miss_rate=0.05)
    load(ctx, g)
    tA = rng.integers(-128, 128, size=(m, 4)).astype(np.int8)
    for tB in (np.array([0, 1, 2, 0], np.int8), np.array([0, 0, 0, 1], np.int8), np.array([-3, 7, 2, -1], np.int8)):
        gi = g.astype(np.int64)
        a = np.take_along_axis(tA.astype(np.int64), gi, axis=1)
        ref = a.T @ tB.astype(np.int64)[gi]
        assert np.array_equal(ctx.table_gram(tA, tB), ref)


def test_rare_alleles_need_more_digits(ctx):
    """Singletons / doubletons give weights 1/(p(1-p)) in the thousands: the format chooser
    must add digits (tensor passes) instead of losing precision."""
    n, m = 1500, 3000
    g = O.synth_geno(n, m, seed=23, miss_rate=0.01, maf_lo=0.0004, maf_hi=0.02)
    keep = O.select_snp_base(g, True, float("nan"), float("nan"))
    g = np.ascontiguousarray(g[keep])
    load(ctx, g)
    got = ctx.grm("GCTA")[0]
    pl = ctx.last_plan()
    # max_abs is the row table T = U / s of the main passes (s <= 127): weights in the thousands / s
    assert pl.max_abs > 100 and pl.digits >= 6, (pl.max_abs, pl.digits)
    assert relerr(got, O.grm_gcta(g)) < TOL
    assert relerr(ctx.grm("Eigenstrat")[0], O.grm_eigenstrat(g)) < TOL
    assert relerr(ctx.grm("EIGMIX")[0], O.grm_eigmix(g)) < TOL


def test_long_k_loop_single_cta():
    """Hundreds of pipeline stages in ONE CTA (SNP splitting disabled): exercises every
    ring-buffer phase flip of the TMA / producer / MMA pipeline, exact integers."""
    rng = np.random.default_rng(5)
    n, m = 300, 50000
    g = O.synth_geno(n, m, seed=31, miss_rate=0.02)
    c = S.Context(0)
    c.debug_flags(2)
    load(c, g)
    tA = rng.integers(-128, 128, size=(m, 4)).astype(np.int8)
    tB = np.array([0, 1, 2, 0], np.int8)
    gi = g.astype(np.int64)
    ref = np.take_along_axis(tA.astype(np.int64), gi, axis=1).T @ tB.astype(np.int64)[gi]
    for _ in range(3):
        assert np.array_equal(c.table_gram(tA, tB), ref)
    assert relerr(c.grm("GCTA")[0], O.grm_gcta(g)) < TOL
    c.close()


def test_row_windows_reproduce_the_packed_result(ctx):
    """N x N outputs tiled into 256-row windows (the mode for N^2 > HBM and for tiling the
    output across GPUs): the concatenated window slices must be the packed full result --
    bit-identical for the integer estimators and GCTA / EIGMIX, 1e-13 for Eigenstrat (its
    trace comes from the per-SNP counts inside a window)."""
    n, m = 700, 3000
    g = O.synth_geno(n, m, seed=17, miss_rate=0.02)
    load(ctx, g)
    full = {k: ctx.grm(k, packed=True)[0] for k in ("GCTA", "EIGMIX", "Eigenstrat")}
    full_ibs = ctx.ibs_ave(packed=True)
    full_num = [O.to_packed_upper(x) for x in ctx.ibs_num()]
    full_king = ctx.king_robust(packed=True)
    full_homo = ctx.king_homo(packed=True)
    got = {k: [] for k in list(full) + ["ibs", "n0", "n1", "n2", "kin0", "kin1", "h0", "h1"]}
    for r0, rows in ctx.windows(256):
        ctx.set_row_window(r0, rows)
        for k in full:
            got[k].append(ctx.grm(k, packed=True)[0])
        got["ibs"].append(ctx.ibs_ave(packed=True))
        a, b, c3 = ctx.ibs_num()
        got["n0"].append(a); got["n1"].append(b); got["n2"].append(c3)
        k0, k1 = ctx.king_robust(packed=True)
        got["kin0"].append(k0); got["kin1"].append(k1)
        h0, h1 = ctx.king_homo(packed=True)
        got["h0"].append(h0); got["h1"].append(h1)
        with pytest.raises(S.SNPRelError, match="packed"):
            ctx.grm("GCTA", packed=False)
    ctx.set_row_window(0, 0)
    cat = {k: np.concatenate(v) for k, v in got.items()}
    assert np.array_equal(cat["GCTA"], full["GCTA"]) and np.array_equal(cat["EIGMIX"], full["EIGMIX"])
    assert relerr(cat["Eigenstrat"], full["Eigenstrat"]) < 1e-12
    assert np.array_equal(cat["ibs"], full_ibs, equal_nan=True)
    assert all(np.array_equal(cat[k], f) for k, f in zip(("n0", "n1", "n2"), full_num))
    assert np.array_equal(cat["kin0"], full_king[0], equal_nan=True) and np.array_equal(cat["kin1"], full_king[1], equal_nan=True)
    assert np.array_equal(cat["h0"], full_homo[0], equal_nan=True) and np.array_equal(cat["h1"], full_homo[1], equal_nan=True)
    assert relerr(ctx.grm("GCTA")[0], O.grm_gcta(g)) < TOL        # back to the whole matrix


def test_degenerate_inputs(ctx):
    g = O.synth_geno(64, 300, seed=4, miss_rate=0.0)
    g[7, :] = 0            # monomorphic SNP
    g[8, :] = 3            # all-missing SNP
    g[9, :] = 2
    g[:, 5] = 3            # a sample with every genotype missing
    load(ctx, g)
    assert np.array_equal(np.stack(ctx.ibs_num()), O.ibs_counts(g))
    ibs = ctx.ibs_ave()
    assert np.isnan(ibs[5, 5]) and np.isnan(ibs[5, 0])          # 0/0, src/genIBS.cpp:472-473
    assert relerr(ctx.grm("GCTA")[0], O.grm_gcta(g)) < TOL
    assert relerr(ctx.grm("EIGMIX")[0], O.grm_eigmix(g)) < TOL
    ibs0, kin = ctx.king_robust()
    r0, rk = O.king_robust(O.king_robust_counts(g))
    assert np.array_equal(np.isnan(ibs0), np.isnan(r0)) and np.array_equal(np.isnan(kin), np.isnan(rk))


def test_selection_and_ratefreq(ctx, hapmap):
    g = hapmap["geno"][:, :90]
    load(ctx, g)
    af, maf, mr = ctx.snp_ratefreq()
    s, num = O.snp_stats(g)
    with np.errstate(invalid="ignore", divide="ignore"):
        raf = np.where(num > 0, s / (2.0 * num), np.nan)
    assert np.allclose(af, raf, rtol=0, atol=0, equal_nan=True)
    assert np.allclose(maf, np.minimum(raf, 1 - raf), rtol=0, atol=0, equal_nan=True)
    assert np.array_equal(mr, 1 - num / 90.0)
    sel, nrm = ctx.select_snp_base(True, 0.05, 0.02)
    ref = O.select_snp_base(g, True, 0.05, 0.02)
    assert np.array_equal(sel, ref) and nrm == int((~ref).sum())
    assert np.array_equal(ctx.geno_copy_u8(), g[ref])


def test_error_paths(ctx):
    with pytest.raises(S.SNPRelError, match="Invalid 'method'"):
        ctx.grm("nope")
    c2 = S.Context(0)
    with pytest.raises(S.SNPRelError, match="no genotype workspace"):
        c2.ibs_ave()
    c2.geno_begin(8, 4)
    with pytest.raises(S.SNPRelError, match="capacity"):
        c2.geno_push_u8(np.zeros((200, 8), np.uint8))
    c2.close()


# ---------------------------------------------------------------- size-independent properties at scale

def test_large_properties(ctx):
    """N=4096, M=65536: estimator cross-consistency that
    needs no CPU reference, plus a spot check of entries against the oracle."""
    n, m = 4096, 65536
    ctx.geno_begin(n, m)
    ctx.geno_synth(m, seed=77, miss_rate=0.005)
    i0, i1, i2 = [x.astype(np.int64) for x in ctx.ibs_num()]
    kc = ctx.king_robust_counts().astype(np.int64)
    bc = ctx.indiv_beta_counts().astype(np.int64)
    nl = i0 + i1 + i2
    assert np.array_equal(nl, kc[1]) and np.array_equal(nl, bc[1])      # three kernels agree on #jointly valid
    assert np.array_equal(i0, kc[0])
    assert np.array_equal(kc[2], i1 + 4 * i0)                           # sum (gi-gj)^2 = #|d|=1 + 4 #|d|=2
    assert np.array_equal(kc[3], kc[4].T)                               # N1_Aa(i,j) == N2_Aa(j,i)
    assert np.array_equal(np.diag(i2), np.diag(nl))                     # a sample is IBS2 with itself
    idx = O.scattered_samples(n, 40, seed=3)                 # first / middle / last tile rows of the same data set
    sub = O.synth_geno(0, m, seed=77, miss_rate=0.005, samples=idx)
    grm, _ = ctx.grm("GCTA")
    assert np.array_equal(grm, grm.T)
    # oracle on the scattered sub-matrix; the all-sample SNP statistics come from the device
    af, _, mr = ctx.snp_ratefreq()
    ref = O.subset_entries(sub, af, "GCTA")
    ix = np.ix_(idx, idx)
    assert relerr(grm[ix], ref) < TOL
    assert np.array_equal(np.stack([i0[ix], i1[ix], i2[ix]]), O.ibs_counts(sub))


def test_gds_bitstream_ingest(hapmap):
    """The fixture's genotype node is a continuous dBit2 stream of 279-sample rows (279 % 4 = 3)."""
    g = hapmap["geno"]                                  # [9088, 279] codes 0..3
    flat = g.reshape(-1).astype(np.uint8)
    pad = (-flat.size) % 4
    q = np.concatenate([flat, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    stream = (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8)
    with S.Context(0) as c:
        c.geno_begin(g.shape[1], g.shape[0])
        c.geno_push_bitstream(stream, 0, 5000)
        c.geno_push_bitstream(stream, 5000, g.shape[0] - 5000)      # second chunk starts mid-byte
        assert np.array_equal(c.geno_copy_u8(), g)


def _leading_subspace_ok(sig, vt, rsig, rvt, k):
    assert np.max(np.abs(sig[:k] - rsig[:k]) / rsig[:k]) < 1e-7
    assert np.max(1 - np.abs(np.sum(vt[:k] * rvt[:k], axis=1))) < 1e-10       # |cos| of each leading vector pair
    # subspace angle of the whole leading block (north star: 1e-6 on the top-k eigenvectors)
    s = np.linalg.svd(vt[:k] @ rvt[:k].T, compute_uv=False)
    assert np.max(np.sqrt(np.maximum(0, 1 - s ** 2))) < 1e-6


@pytest.mark.parametrize("n,m,aux_dim,it", [(300, 4000, 16, 10), (700, 9000, 8, 4), (120, 2500, 16, 10)])
def test_randomized_pca_vs_oracle(n, m, aux_dim, it):
    """snprel_pca_randomized against the oracle's restatement of CRandomPCA (pinned to the reference
    itself, num.thread = 1, in tests/test_oracle_vs_reference.py).  The third case has
    hsize = aux_dim (it + 1) > n_samp: the wide branch of the final SVD."""
    g = O.synth_geno(n, m, seed=41, miss_rate=0.02)
    aux = np.random.default_rng(n).standard_normal(aux_dim * n)
    rsig, rvt, rtr = O.pca_randomized(g, aux, aux_dim, it)
    with S.Context(0) as c:
        c.geno_begin(n, m)
        c.geno_push_u8(g)
        sig, vt, tr = c.pca_randomized(aux, aux_dim, it)
    assert abs(tr - rtr) <= 1e-12 * rtr
    assert vt.shape == (aux_dim * (it + 1), n) and sig.shape == (n,)
    r = min(n, aux_dim * (it + 1))
    assert np.all(sig[r:] == 0) and np.all(vt[r:] == 0)
    assert np.max(np.abs(vt[:r] @ vt[:r].T - np.eye(r))) < 1e-10               # orthonormal rows
    _leading_subspace_ok(sig, vt, rsig, rvt, 6)


def test_randomized_pca_through_the_api(gds, hapmap):
    """snpgdsPCA(algorithm="randomized") (R/PCA.R:54-62,80-89): same result object as the exact
    algorithm, leading components close to it."""
    n = 279
    aux = np.random.default_rng(1).standard_normal(32 * n)
    r = S.snpgdsPCA(gds, algorithm="randomized", aux_mat=aux)
    e = S.snpgdsPCA(gds, algorithm="exact", eigen_cnt=16)
    assert r["eigenvect"].shape == (n, 16) and r["eigenval"].shape == (n,) and r["Bayesian"] is False
    assert np.max(np.abs(r["eigenval"][:4] - e["eigenval"][:4]) / e["eigenval"][:4]) < 1e-6
    assert np.max(1 - np.abs(np.sum(r["eigenvect"][:, :4] * e["eigenvect"][:, :4], axis=0))) < 1e-6
    assert abs(r["TraceXTX"] - 2.0 * 0.5 * e["TraceXTX"]) / e["TraceXTX"] < 1e-12   # 2 x sum y^2 with y = z / sqrt(2)
    with pytest.raises(S.SNPRelError):
        S.snpgdsPCA(gds, algorithm="fast")


def _pack2b(g, pitch):
    """uint8 [m, n] codes -> 2-bit rows (4 per byte, LSB first) with `pitch` bytes per row, padding = missing"""
    m, n = g.shape
    q = np.full((m, pitch * 4), 3, dtype=np.uint8)
    q[:, :n] = g
    q = q.reshape(m, pitch, 4)
    return np.ascontiguousarray(q[:, :, 0] | (q[:, :, 1] << 2) | (q[:, :, 2] << 4) | (q[:, :, 3] << 6))


@pytest.mark.parametrize("pitch_pad", [0, 7])
def test_streamed_ingest_matches_the_blocking_path(pitch_pad):
    """snprel_geno_push_2b_async + snprel_pca / snprel_grm: the copy chunks are consumed while later chunks
    are in flight, with a speculative fixed-point format verified at the end.  Exact integer planes and the
    same epilogue: the result must equal the blocking path's bit for bit whenever the format agrees, and the
    oracle to 1e-10 in any case."""
    import torch
    n, m = 520, 131072 * 3 + 4321                  # four copy chunks, the last one ragged
    g = O.synth_geno(n, m, seed=77, miss_rate=0.01)
    pitch = (n + 3) // 4 + pitch_pad
    host = torch.from_numpy(_pack2b(g, pitch)).pin_memory()
    idx = O.scattered_samples(n, 40, seed=3)
    with S.Context(0) as c:
        c.geno_begin(n, m)
        c.geno_push_2b(host.numpy())
        ref = c.pca(genmat_only=True)
        plan_ref = c.last_plan()
        gcta_ref = c.grm("GCTA")[0]
        c.geno_begin(n, m)
        c.geno_push_2b_async(host.numpy())
        got = c.pca(genmat_only=True)
        plan = c.last_plan()
        assert c.stream_stats() == (1, 0)
        fmt = lambda p: (p.digits, p.digits_w, p.frac_bits, p.frac_bits_w, p.rounding)
        if fmt(plan) == fmt(plan_ref):
            assert np.array_equal(got["genmat"], ref["genmat"])
        # (different formats: two quantisations of the same matrix, each within the tolerance of the exact one)
        assert relerr(got["genmat"], ref["genmat"]) < 5e-11 and abs(got["TraceXTX"] - ref["TraceXTX"]) <= 1e-12 * ref["TraceXTX"]
        af, _, _ = c.snp_ratefreq()
        sub = g[:, idx]
        oref = O.subset_entries(sub, af, "Eigenstrat", n_total=n, trace=got["TraceXTX"])
        assert relerr(got["genmat"][np.ix_(idx, idx)], oref) < TOL
        # GCTA (one more plane for the missing-pair denominators) and a second use of the same context
        c.geno_begin(n, m)
        c.geno_push_2b_async(host.numpy())
        assert relerr(c.grm("GCTA")[0], gcta_ref) < 5e-11
        assert c.stream_stats() == (2, 0)
        # any other entry point simply waits for the copies
        c.geno_begin(n, m)
        c.geno_push_2b_async(host.numpy())
        assert np.array_equal(np.stack([a[np.ix_(idx, idx)] for a in c.ibs_num()]), O.ibs_counts(sub))


def test_streamed_ingest_falls_back_when_the_guess_fails():
    """Later chunks with much rarer alleles than the first one: the speculative format overflows its digits,
    the verification notices and the ordinary path recomputes the result."""
    import torch
    n, m1 = 300, 131072
    g1 = O.synth_geno(n, m1, seed=5, miss_rate=0.0, maf_lo=0.3, maf_hi=0.5)
    g2 = O.synth_geno(n, 2 * m1, seed=6, miss_rate=0.02, maf_lo=0.002, maf_hi=0.01)
    g = np.concatenate([g1, g2])
    keep = O.select_snp_base(g, True, float("nan"), float("nan"))
    keep[:m1] = True
    g = np.ascontiguousarray(g[keep])
    host = torch.from_numpy(_pack2b(g, (n + 3) // 4)).pin_memory()
    with S.Context(0) as c:
        c.geno_begin(n, g.shape[0])
        c.geno_push_2b_async(host.numpy())
        got = c.grm("GCTA")[0]
        streamed, fallbacks = c.stream_stats()
        assert streamed + fallbacks == 1
        assert relerr(got, O.grm_gcta(g)) < TOL


def test_async_output_of_row_windows():
    """snprel_set_async_output: the device-to-host copy of window w is only queued when the finishing call
    returns (it overlaps the next window's accumulate); after output_wait the slices equal the blocking ones."""
    import torch
    n, m = 1100, 6000
    g = O.synth_geno(n, m, seed=61, miss_rate=0.02)
    with S.Context(0) as c:
        c.geno_begin(n, m)
        c.geno_push_u8(g)
        ref_ibs = c.ibs_ave(packed=True)
        ref_king = c.king_robust(None, packed=True)
        ref_grm = c.grm("GCTA", packed=True)[0]
        c.set_async_output(True)
        bufs = [torch.empty(n * 256, dtype=torch.float64, pin_memory=True).numpy() for _ in range(8)]
        for name, ref in (("ibs", [ref_ibs]), ("king", list(ref_king)), ("grm", [ref_grm])):
            got = [np.full(n * (n + 1) // 2, np.nan) for _ in ref]
            pos, pending = 0, None
            for k, (r0, h) in enumerate(c.windows(256)):
                c.set_row_window(r0, h)
                cnt = c.window_count()
                mine = bufs[2 * (k % 4): 2 * (k % 4) + 2]           # a buffer pair is reused four windows later
                if name == "ibs":
                    out = [c.ibs_ave(packed=True, out=mine[0])]
                elif name == "king":
                    out = list(c.king_robust(None, packed=True, out=mine))
                else:
                    out = [c.grm("GCTA", packed=True, out=mine[0])[0]]
                if pending is not None:                              # the previous window: wait, then read
                    c.output_wait()
                # (the call above already waited for the previous copy before rewriting the device scratch)
                if pending is not None:
                    p0, pc, pout = pending
                    for gk, o in zip(got, pout):
                        gk[p0:p0 + pc] = o[:pc]
                pending = (pos, cnt, out)
                pos += cnt
            c.output_wait()
            p0, pc, pout = pending
            for gk, o in zip(got, pout):
                gk[p0:p0 + pc] = o[:pc]
            c.set_row_window(0, 0)
            for gk, r in zip(got, ref):
                assert np.array_equal(gk, r, equal_nan=True), name
        c.set_async_output(False)
        assert np.array_equal(c.ibs_ave(packed=True), ref_ibs)
