"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

A plain numpy (float64 / int64) restatement of the reference's N x N
relatedness-matrix path.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` leg of ``bench.py`` may import this module; the product
(``snprelate_b200``) must never route through it.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function
below against the reference's own golden vectors
(``inst/unitTests/valid/Validate.{IBS,PCA,KING,Beta,EIGMIX}.RData`` decoded
into ``tests/golden/``) on the bundled HapMap fixture.  GCTA-with-missing has no
golden in the reference; it follows
``src/genPCA.cpp:1148-1237`` literally and is additionally checked against the
compiled reference sources in ``oracle/_ref`` when those are built.

Genotype convention everywhere: ``geno`` is uint8 ``[nsnp, nsamp]`` (SNP-major,
sample fastest -- the layout ``CdBaseWorkSpace::snpRead(..., RDim_Sample_X_SNP)``
produces, ``src/dGenGWAS.cpp:677-733``); values 0/1/2 = number of A alleles,
anything > 2 = missing (``src/dGenGWAS.cpp:33-39``).
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------
# SNP selection  (CdBaseWorkSpace::Select_SNP_Base, src/dGenGWAS.cpp:361-397;
# Get_AF_MR_perSNP :472-552; thresholds as mapped by R/Internal.R:438-439)
# ---------------------------------------------------------------------------


def snp_stats(geno: np.ndarray):
    """Per-SNP (sum, num): vec_u8_geno_count, src/dVect.cpp:30-117."""
    valid = geno <= 2
    num = valid.sum(axis=1).astype(np.int64)
    s = np.where(valid, geno, 0).sum(axis=1).astype(np.int64)
    return s, num


def select_snp_base(geno: np.ndarray, remove_monosnp=True, maf=np.nan,
                    missing_rate=np.nan) -> np.ndarray:
    """Boolean mask of kept SNPs (src/dGenGWAS.cpp:361-397)."""
    maf_t = -1.0 if not np.isfinite(maf) else float(maf)        # R/Internal.R:438
    mr_t = 2.0 if not np.isfinite(missing_rate) else float(missing_rate)  # :439
    s, num = snp_stats(geno)
    nsamp = geno.shape[1]
    with np.errstate(divide="ignore", invalid="ignore"):
        af = np.where(num > 0, s / (2.0 * num), np.nan)
    mafv = np.minimum(af, 1 - af)
    mr = 1.0 - num / float(nsamp)
    keep = np.isfinite(mafv)
    if remove_monosnp:
        keep &= ~(mafv <= 0)
    keep &= ~(mafv < maf_t)
    keep &= ~(mr > mr_t)
    return keep


# ---------------------------------------------------------------------------
# GRM / PCA / EIGMIX numerators
# ---------------------------------------------------------------------------


def _avg_geno(geno):
    """DivideGeno: avg = sum/num, 0 if num == 0 (src/genPCA.cpp:98-142)."""
    s, num = snp_stats(geno)
    avg = np.where(num > 0, s / np.maximum(num, 1), 0.0)
    return s, num, avg


def _rsqrt_prod(avg):
    """scale = 1/sqrt(p(1-p)), p = avg/2, 0 unless 0<p<1 (src/genPCA.cpp:145-181)."""
    p = avg * 0.5
    ok = (0 < p) & (p < 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        sc = np.where(ok, 1.0 / np.sqrt(p * (1 - p)), 0.0)
    return sc


def _centred(geno, avg, scale=None):
    """TransposeGenotype + GenoSub (+ GenoMul): Z[i,l] = (g - avg_l) * scale_l,
    missing -> exactly 0 (src/genPCA.h:93-108, src/genPCA.cpp:315-368).
    Returns Z as [nsamp, nsnp] float64."""
    valid = geno <= 2
    z = np.where(valid, geno.astype(np.float64) - avg[:, None], 0.0)
    if scale is not None:
        z = z * scale[:, None]
    return np.ascontiguousarray(z.T)


def _zzt(z, block=4096):
    n = z.shape[0]
    c = np.zeros((n, n))
    for s in range(0, z.shape[1], block):
        zb = z[:, s:s + block]
        c += zb @ zb.T
    return c


def cov_eigenstrat(geno, bayesian=False):
    """CExactPCA::Run (src/genPCA.cpp:395-464): un-normalised Z Z^T."""
    s, num, avg = _avg_geno(geno)
    if bayesian:
        p = (s + 1.0) / (2 * num + 2)                # :445-452
        scale = 1.0 / np.sqrt(p * (1 - p))
    else:
        scale = _rsqrt_prod(avg)
    return _zzt(_centred(geno, avg, scale))


def pca_genmat(geno, bayesian=False):
    """gnrPCA normalisation (src/genPCA.cpp:1381-1390).
    Returns (genmat, TraceXTX, TraceVal)."""
    c = cov_eigenstrat(geno, bayesian)
    n = c.shape[0]
    tr = np.trace(c)
    c = c * ((n - 1) / tr)
    return c, tr, np.trace(c)


def pca_randomized(geno, aux_mat, aux_dim, iter_num=10):
    """CRandomPCA::Run (src/genPCA.cpp:684-790), the randomized algorithm of gnrPCA (:1436-1442).
    aux_mat: the R vector rnorm(aux.dim * n.samp) (R/PCA.R:56), read by the reference as the
    column-major nSamp x AuxDim matrix G_0 (AuxMat[j * nSamp + i], :535-563).
      Y[l][g] = (g - avg_l) / sqrt(2 p_l (1 - p_l)), 0 for missing / monomorphic      (:497-517)
      H_it = Y G_it,  G_{it+1} = Y^T H_it / nSNP,  it = 0 .. iter_num                  (:710-742)
      V^T = right singular vectors of MatH [hsize x nSNP] (dgesvd "N","O", :642-662,752)
      T = V^T Y [hsize x nSamp]  (:619-640,763-768),  sigma, V_T^T = svd(T)             (:778-779)
    Returns (sigma padded with zeros to nSamp, V_T^T [min(hsize, nSamp), nSamp], 2 TraceXTX) -- the
    three list elements gnrPCA hands back (:781-793)."""
    geno = np.asarray(geno)
    m, n = geno.shape
    s, num, avg = _avg_geno(geno)
    p = avg * 0.5
    ok = (p > 0) & (p < 1)
    sc = np.where(ok, 1.0 / np.sqrt(np.where(ok, 2 * p * (1 - p), 1.0)), 0.0)
    y = np.where(geno <= 2, (geno.astype(np.float64) - avg[:, None]) * sc[:, None], 0.0)      # [nSNP, nSamp]
    trace = float((y * y).sum())
    g = np.asarray(aux_mat, dtype=np.float64).reshape(aux_dim, n).T.copy()                    # [nSamp, AuxDim]
    hs = []
    for it in range(iter_num + 1):
        h = y @ g
        hs.append(h)
        if it < iter_num:
            g = (y.T @ h) / m
    math = np.hstack(hs)                                                                      # [nSNP, hsize]
    u, _, _ = np.linalg.svd(math, full_matrices=False)        # columns = rows of the reference's V^T
    t = u.T @ y
    _, sig, vt = np.linalg.svd(t, full_matrices=False)
    sigma = np.zeros(n)
    sigma[:len(sig)] = sig
    return sigma, vt, 2.0 * trace


def pca_randomized_result(sigma, vt, trace2, n_samp, eigen_cnt):
    """R/PCA.R:80-89: varprop = 2 d^2 / TraceXTX, eigenval = (n - 1) varprop, eigenvect = t(V[1:eigen.cnt, ])."""
    vp = 2.0 * np.asarray(sigma) ** 2 / trace2
    return dict(eigenval=(n_samp - 1) * vp, eigenvect=np.asarray(vt)[:eigen_cnt].T.copy(), varprop=vp, TraceXTX=trace2)


def pca_eigen(genmat, eigen_cnt):
    """CalcEigen on -C (src/genPCA.cpp:1262-1346): top-k eigenpairs, descending.
    The reference's own tests pin eigenvectors only through rounded loadings,
    so this is pinned against LAPACK (numpy eigh) on the oracle matrix."""
    w, v = np.linalg.eigh(-genmat)
    k = min(eigen_cnt, genmat.shape[0])
    return -w[:k], v[:, :k]


# ---------------------------------------------------------------------------
# SNP loadings, sample loadings (projection) and SNP-PC correlations
# (the tall-skinny products either side of the eigen step)
# ---------------------------------------------------------------------------


def pca_snp_loading(geno, eigenval, eigenvect, trace_xtx, bayesian=False):
    """gnrPCASNPLoading + CPCA_SNPLoad::thread_loading (src/genPCA.cpp:1489-1540,
    938-1040): eigenvectors scaled by sqrt(((n-1)/TraceXTX)/eigenval_k), then
    loading[k, l] = sum_j (g_jl - avg_l) scale_l v_jk over non-missing genotypes.
    Returns (loading [k, nsnp], avgfreq [nsnp], scale [nsnp])."""
    n = geno.shape[1]
    k = eigenvect.shape[1]
    v = eigenvect * np.sqrt(((n - 1) / trace_xtx) / np.asarray(eigenval[:k], dtype=np.float64))[None, :]
    s, num, avg = _avg_geno(geno)
    if bayesian:                                  # :963-966
        p = (s + 1.0) / (2 * num + 2)
        scale = np.where(num > 0, 1.0 / np.sqrt(p * (1 - p)), 0.0)
    else:
        scale = np.where(num > 0, _rsqrt_prod(avg), 0.0)
    z = np.where(geno <= 2, (geno.astype(np.float64) - avg[:, None]) * scale[:, None], 0.0)
    return (z @ v).T, avg, scale


def pca_samp_loading(geno, loadings, avgfreq, scale):
    """gnrPCASampLoading + CPCA_SampleLoad::thread_loading (src/genPCA.cpp:1542-1563,
    1042-1123): out[i, k] = sum_l (g_il - avgfreq_l) scale_l loadings[k, l].  `loadings`
    is what R passes: snploading * sqrt(((n0-1)/TraceXTX)/eigenval) (R/PCA.R:274-281)."""
    z = np.where(geno <= 2, (geno.astype(np.float64) - avgfreq[:, None]) * scale[:, None], 0.0)
    return z.T @ np.asarray(loadings, dtype=np.float64).T


def pca_corr(geno, eigenvect):
    """gnrPCACorr + CPCA_SNPCorr::SNP_PC_Corr (src/genPCA.cpp:1456-1485, 809-936):
    Pearson correlation between each SNP's genotypes and each eigenvector over the
    SNP's non-missing samples; NaN when fewer than 2 samples or a zero variance.
    Returns [k, nsnp]."""
    valid = (geno <= 2).astype(np.float64)
    y = np.where(geno <= 2, geno, 0).astype(np.float64)
    v = np.asarray(eigenvect, dtype=np.float64)
    m = valid.sum(axis=1)[:, None]
    XY, X, XX = y @ v, valid @ v, valid @ (v * v)
    Y, YY = y.sum(axis=1)[:, None], (y * y).sum(axis=1)[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        c1, c2 = XX - X * X / m, YY - Y * Y / m
        val = c1 * c2
        out = np.where((m > 1) & (val > 0), (XY - X * Y / m) / np.sqrt(np.where(val > 0, val, 1.0)), np.nan)
    return out.T


def _eigmix_z(geno, afreq):
    """(g - 2 af_l) / sqrt(sum_l 4 af_l (1 - af_l)), missing -> 0 (src/genEIGMIX.cpp:455-478,508-513)."""
    afreq = np.asarray(afreq, dtype=np.float64)
    sc = 1.0 / np.sqrt(np.sum(4 * afreq * (1 - afreq)))
    return np.where(geno <= 2, (geno.astype(np.float64) - 2 * afreq[:, None]) * sc, 0.0)


def eigmix_snp_loading(geno, eigenval, eigenvect, afreq):
    """gnrEigMixSNPLoading + CEigMix_SNPLoad (src/genEIGMIX.cpp:739-775, 445-530) -> [k, nsnp]."""
    k = eigenvect.shape[1]
    v = eigenvect * np.sqrt(1.0 / np.asarray(eigenval[:k], dtype=np.float64))[None, :]
    return (_eigmix_z(geno, afreq) @ v).T


def eigmix_samp_loading(geno, loadings, afreq):
    """gnrEigMixSampLoading + CEigMix_SampleLoad (src/genEIGMIX.cpp:777-803, 534-640) -> [n, k]."""
    return _eigmix_z(geno, afreq).T @ np.asarray(loadings, dtype=np.float64).T


def _missing_pair_denom(geno, d):
    """Denom[i,j] = sum_l d_l [i missing or j missing]
    (src/genPCA.cpp:1201-1224, src/genEIGMIX.cpp:113-138)."""
    m = (geno > 2).astype(np.float64).T          # [nsamp, nsnp]
    dm = m * d[None, :]
    r = dm.sum(axis=1)
    return r[:, None] + r[None, :] - dm @ m.T


def grm_gcta(geno):
    """CGCTA_AlgArith::Run (src/genPCA.cpp:1148-1237)."""
    s, num, avg = _avg_geno(geno)
    scale = _rsqrt_prod(avg)
    c = _zzt(_centred(geno, avg, scale))
    poly = ((0 < s) & (s < 2 * num)).astype(np.float64)      # :1206
    nlocus = int(poly.sum())
    denom = np.rint(_missing_pair_denom(geno, poly))
    return c / (2.0 * (nlocus - denom))


def grm_eigenstrat(geno):
    """gnrGRM method "Eigenstrat" (src/genPCA.cpp:1636-1647)."""
    return pca_genmat(geno)[0]


def grm_corr(geno):
    """gnrGRM method "Corr" (src/genPCA.cpp:1658-1685)."""
    g = grm_gcta(geno)
    d = np.sqrt(np.diag(g))
    out = g / (d[:, None] * d[None, :])
    np.fill_diagonal(out, 1.0)
    return out


def eigmix_ibd(geno, diagadj=True):
    """CEigMix_AlgArith::Run (src/genEIGMIX.cpp:60-156).
    Returns (ibd, afreq)."""
    s, num, avg = _avg_geno(geno)
    c = _zzt(_centred(geno, avg, None))
    af = 0.5 * avg
    d = 4 * af * (1 - af)
    sum_den = d.sum()
    denom = _missing_pair_denom(geno, d)
    if diagadj:
        het = (geno == 1).sum(axis=0).astype(np.float64)
        c[np.diag_indices_from(c)] -= het
    return c / (sum_den - denom), af


def grm_eigmix(geno):
    """CalcEigMixGRM (src/genEIGMIX.cpp:645-652)."""
    return 2.0 * eigmix_ibd(geno, diagadj=False)[0]


# ---------------------------------------------------------------------------
# Bit-plane packers and packed-bit pair kernels
# ---------------------------------------------------------------------------

_B1 = np.array([0, 1, 1, 0], dtype=np.uint8)   # src/dGenGWAS.cpp:1429-1475
_B2 = np.array([0, 0, 1, 1], dtype=np.uint8)


def pack_geno1b(geno, n_total=None):
    """PackSNPGeno1b for every sample: returns (plane1, plane2) uint8
    [nsamp, n_total/8]; SNP l sits at bit (l % 8) of byte l // 8; padding SNPs
    encode as missing (plane1=0, plane2=1) (src/dGenGWAS.cpp:1467-1472)."""
    nsnp, nsamp = geno.shape
    if n_total is None:
        n_total = (nsnp + 7) // 8 * 8
    g = np.minimum(geno, 3).T                      # [nsamp, nsnp]
    b1 = np.zeros((nsamp, n_total), dtype=np.uint8)
    b2 = np.ones((nsamp, n_total), dtype=np.uint8)
    b1[:, :nsnp] = _B1[g]
    b2[:, :nsnp] = _B2[g]
    p1 = np.packbits(b1, axis=1, bitorder="little")
    p2 = np.packbits(b2, axis=1, bitorder="little")
    return p1, p2


def _popcount_rows(x):
    return np.unpackbits(x, axis=-1).sum(axis=-1, dtype=np.int64)


def ibs_counts_packed(geno):
    """CIBSCount::thread_ibs_num (src/genIBS.cpp:154-273), literally on the
    bit planes.  O(N^2 M/8) numpy work -- small inputs only."""
    p1, p2 = pack_geno1b(geno)
    n = p1.shape[0]
    out = np.zeros((3, n, n), dtype=np.int64)
    for i in range(n):
        a1, a2 = p1[i][None, :], p2[i][None, :]
        mask = (a1 | ~a2) & (p1 | ~p2)
        ibs0 = ~((a1 ^ ~p1) | (a2 ^ ~p2)) & mask
        ibs2 = ~((a1 ^ p1) | (a2 ^ p2)) & mask
        n0 = _popcount_rows(ibs0)
        n2 = _popcount_rows(ibs2)
        out[0, i] = n0
        out[2, i] = n2
        out[1, i] = _popcount_rows(mask) - n0 - n2
    return out


def _channels(geno):
    valid = (geno <= 2)
    x = np.where(valid, geno, 0).astype(np.float64).T     # [nsamp, nsnp]
    a = valid.astype(np.float64).T
    return x, a


def ibs_counts(geno):
    """Same integers as ibs_counts_packed via indicator Grams (exact in f64
    while counts < 2^53).  Returns int64 [3, n, n] = IBS0, IBS1, IBS2."""
    valid = geno <= 2
    e = [((geno == k) & valid).astype(np.float64).T for k in range(3)]
    a = valid.astype(np.float64).T
    nv = a @ a.T
    ibs2 = sum(ek @ ek.T for ek in e)
    ibs0 = e[0] @ e[2].T + e[2] @ e[0].T
    out = np.stack([ibs0, nv - ibs0 - ibs2, ibs2])
    return np.rint(out).astype(np.int64)


def ibs_ave(counts):
    """gnrIBSAve (src/genIBS.cpp:463-490)."""
    c = counts.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return (0.5 * c[1] + c[2]) / (c[0] + c[1] + c[2])


def king_robust_counts(geno):
    """CKINGRobust::thread_ibs_num (src/genKING.cpp:292-426): int64 [5, n, n]
    = IBS0, nLoci, SumSq, N1_Aa, N2_Aa where N1_Aa[i,j] counts het loci of the
    ROW sample i restricted to loci where j is valid (:328,377)."""
    valid = geno <= 2
    x, a = _channels(geno)
    e0 = ((geno == 0) & valid).astype(np.float64).T
    e1 = ((geno == 1) & valid).astype(np.float64).T
    e2 = ((geno == 2) & valid).astype(np.float64).T
    nloci = a @ a.T
    ibs0 = e0 @ e2.T + e2 @ e0.T
    x2 = x * x
    sumsq = x2 @ a.T + a @ x2.T - 2 * (x @ x.T)
    n1 = e1 @ a.T
    n2 = a @ e1.T
    return np.rint(np.stack([ibs0, nloci, sumsq, n1, n2])).astype(np.int64)


def king_robust(counts, family_id=None):
    """gnrIBD_KING_Robust epilogue (src/genKING.cpp:626-641).
    family_id: int array with negative / None entries meaning NA."""
    ibs0, nloci, sumsq, n1, n2 = [c.astype(np.float64) for c in counts]
    n = ibs0.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        r_ibs0 = np.where(nloci > 0, ibs0 / nloci, np.nan)
        between = 0.5 - sumsq / (4.0 * np.minimum(n1, n2))
        within = 0.5 - sumsq / (2.0 * (n1 + n2))
    if family_id is None:
        kin = between
    else:
        f = np.asarray(family_id)
        same = (f[:, None] == f[None, :]) & (f[:, None] >= 0)
        kin = np.where(same, within, between)
    kin = np.where(np.isfinite(kin), kin, np.nan)
    r_ibs0[np.diag_indices(n)] = 0.0
    kin[np.diag_indices(n)] = 0.5
    return r_ibs0, kin


def king_homo(geno):
    """CKINGHomo + gnrIBD_KING_Homo (src/genKING.cpp:69-266,493-570).
    Returns (k0, k1)."""
    s, num = snp_stats(geno)
    p = np.where(num > 0, 0.5 * s / np.maximum(num, 1), 0.0)
    af = p * (1 - p)
    x, a = _channels(geno)
    valid = geno <= 2
    e0 = ((geno == 0) & valid).astype(np.float64).T
    e2 = ((geno == 2) & valid).astype(np.float64).T
    ibs0 = e0 @ e2.T + e2 @ e0.T
    x2 = x * x
    sumsq = x2 @ a.T + a @ x2.T - 2 * (x @ x.T)
    s1 = (a * af[None, :]) @ a.T
    s2 = (a * (af * af)[None, :]) @ a.T
    with np.errstate(divide="ignore", invalid="ignore"):
        theta = 0.5 - sumsq / (8 * s1)
        k0 = ibs0 / (2 * s2)
        k1 = 2 - 2 * k0 - 4 * theta
    k0 = np.where(np.isfinite(k0), k0, np.nan)
    k1 = np.where(np.isfinite(k1), k1, np.nan)
    n = k0.shape[0]
    k0[np.diag_indices(n)] = 0.0
    k1[np.diag_indices(n)] = 0.0
    return k0, k1


def beta_counts(geno):
    """CIndivBeta::thread_ibs_num (src/genBeta.cpp:65-178): int64 [2, n, n]
    = ibscnt, num.  ibscnt = #(either het, both valid) + 2 #(same homozygote)."""
    valid = geno <= 2
    a = valid.astype(np.float64).T
    e0 = ((geno == 0) & valid).astype(np.float64).T
    e1 = ((geno == 1) & valid).astype(np.float64).T
    e2 = ((geno == 2) & valid).astype(np.float64).T
    num = a @ a.T
    anyhet = e1 @ a.T + a @ e1.T - e1 @ e1.T
    samehom = e0 @ e0.T + e2 @ e2.T
    return np.rint(np.stack([anyhet + 2 * samehom, num])).astype(np.int64)


def _offdiag_upper_sum_rowmajor(m):
    """avg += s in the reference's row-major upper-triangle order."""
    n = m.shape[0]
    iu = np.triu_indices(n, 1)
    return m[iu].sum()


def indiv_beta(counts, inbreeding=True):
    """gnrIBD_Beta (src/genBeta.cpp:361-460). Returns (beta, avg)."""
    ibscnt, num = [c.astype(np.float64) for c in counts]
    n = num.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        b = 0.5 * ibscnt / num
        d = np.diag(ibscnt) / np.diag(num) - 1 if inbreeding else np.diag(b).copy()
    avg = _offdiag_upper_sum_rowmajor(b) / (n * (n - 1) // 2)
    b[np.diag_indices(n)] = d
    return (b - avg) * (1.0 / (1 - avg)), avg


def grm_indivbeta(counts):
    """CalcIndivBetaGRM (src/genBeta.cpp:308-357). Returns (grm, avg)."""
    ibscnt, num = [c.astype(np.float64) for c in counts]
    n = num.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        b = 0.5 * ibscnt / num
        b[np.diag_indices(n)] = np.diag(ibscnt) / np.diag(num) - 1
    avg = _offdiag_upper_sum_rowmajor(b) / (n * (n - 1) // 2)
    mn = b.min()
    scale = 2.0 / (1 - mn)
    out = (b - mn) * scale
    out[np.diag_indices(n)] = (np.diag(b) - mn) * scale * 0.5 + 1
    return out, avg


# ---------------------------------------------------------------------------
# Packed triangle helpers  (CdMatTri, src/dGenGWAS.h:511-583)
# ---------------------------------------------------------------------------


def ibd_mom_tables(geno, allele_freq=None):
    """IBD::Init_EPrIBD_IBS (src/genIBD.cpp:253-338): expected P(IBS i | IBD j)
    averaged over SNPs, sequential f64 sums in SNP order.  Without allele_freq
    the frequency is that of the A allele from the genotype counts and PLINK's
    finite-sample correction factor is applied; with allele_freq the counts are
    zero and no correction is used (src/genIBS.cpp:585-587).
    Returns (E[3,3], afreq[nsnp])."""
    nsnp = geno.shape[0]
    if allele_freq is None:
        aa = (geno == 2).sum(axis=1).astype(np.int64)
        ab = (geno == 1).sum(axis=1).astype(np.int64)
        bb = (geno == 0).sum(axis=1).astype(np.int64)
    else:
        aa = ab = bb = np.zeros(nsnp, dtype=np.int64)
    e = np.zeros((3, 3))
    afreq = np.empty(nsnp)
    nvalid = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(nsnp):
            n = 2 * int(aa[i] + ab[i] + bb[i])
            p = (2.0 * aa[i] + ab[i]) / n if n > 0 else np.nan
            if allele_freq is not None:
                p = float(allele_freq[i])
                if np.isfinite(p) and (p < 0 or p > 1):
                    p = np.nan
            afreq[i] = p
            q = 1 - p
            na = np.float64(n)
            x = np.float64(2 * aa[i] + ab[i])
            y = np.float64(2 * bb[i] + ab[i])
            if allele_freq is None:
                f3 = (na / (na - 1)) * (na / (na - 2)) * (na / (na - 3))
                a00 = 2*p*p*q*q * ((x-1)/x * (y-1)/y * f3)
                a01 = (4*p*p*p*q * ((x-1)/x * (x-2)/x * f3)
                       + 4*p*q*q*q * ((y-1)/y * (y-2)/y * f3))
                a02 = (q*q*q*q * ((y-1)/y * (y-2)/y * (y-3)/y * f3)
                       + p*p*p*p * ((x-1)/x * (x-2)/x * (x-3)/x * f3)
                       + 4*p*p*q*q * ((x-1)/x * (y-1)/y * f3))
                f2 = na / (na - 1) * na / (na - 2)
                a11 = (2*p*p*q * ((x-1)/x * na/(na-1) * na/(na-2))
                       + 2*p*q*q * ((y-1)/y * na/(na-1) * na/(na-2)))
                a12 = (p*p*p * ((x-1)/x * (x-2)/x * na/(na-1) * na/(na-2))
                       + q*q*q * ((y-1)/y * (y-2)/y * na/(na-1) * na/(na-2))
                       + p*p*q * ((x-1)/x * na/(na-1) * na/(na-2))
                       + p*q*q * ((y-1)/y * na/(na-1) * na/(na-2)))
                del f2
            else:
                a00 = 2*p*p*q*q
                a01 = 4*p*p*p*q + 4*p*q*q*q
                a02 = q*q*q*q + p*p*p*p + 4*p*p*q*q
                a11 = 2*p*p*q + 2*p*q*q
                a12 = p*p*p + q*q*q + p*p*q + p*q*q
            if all(np.isfinite(v) for v in (a00, a01, a02, a11, a12)):
                e[0, 0] += a00
                e[0, 1] += a01
                e[0, 2] += a02
                e[1, 1] += a11
                e[1, 2] += a12
                nvalid += 1
        nv = np.float64(nvalid)
        e[0, 0] /= nv
        e[0, 1] /= nv
        e[0, 2] /= nv
        e[1, 1] /= nv
        e[1, 2] /= nv
    e[2, 2] = 1.0
    return e, afreq


def ibd_mom(counts, e, kinship_constraint=False):
    """IBD::Est_PLINK_Kinship over all pairs (src/genIBD.cpp:341-383) as
    gnrIBD_PLINK applies it (src/genIBS.cpp:591-607): diagonal k0 = k1 = 0.
    counts = int [3, n, n] IBS0/1/2.  Returns (k0, k1) f64 [n, n]."""
    c = counts.astype(np.float64)
    ntot = (counts[0] + counts[1] + counts[2]).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        e00, e01, e11 = e[0, 0] * ntot, e[0, 1] * ntot, e[1, 1] * ntot
        e02, e12, e22 = e[0, 2] * ntot, e[1, 2] * ntot, e[2, 2] * ntot
        k0 = c[0] / e00
        k1 = (c[1] - k0 * e01) / e11
        k2 = (c[2] - k0 * e02 - k1 * e12) / e22
        # the six sequential clamps, each seeing the previous one's result
        m = k0 > 1
        k0 = np.where(m, 1.0, k0); k1 = np.where(m, 0.0, k1); k2 = np.where(m, 0.0, k2)
        m = k1 > 1
        k1 = np.where(m, 1.0, k1); k0 = np.where(m, 0.0, k0); k2 = np.where(m, 0.0, k2)
        m = k2 > 1
        k2 = np.where(m, 1.0, k2); k0 = np.where(m, 0.0, k0); k1 = np.where(m, 0.0, k1)
        m = k0 < 0
        s = k1 + k2
        k1 = np.where(m, k1 / s, k1); k2 = np.where(m, k2 / s, k2); k0 = np.where(m, 0.0, k0)
        m = k1 < 0
        s = k0 + k2
        k0 = np.where(m, k0 / s, k0); k2 = np.where(m, k2 / s, k2); k1 = np.where(m, 0.0, k1)
        m = k2 < 0
        s = k0 + k1
        k0 = np.where(m, k0 / s, k0); k1 = np.where(m, k1 / s, k1); k2 = np.where(m, 0.0, k2)
        if kinship_constraint:
            k2 = 1 - k0 - k1
            pihat = k1 / 2 + k2
            m = pihat * pihat < k2
            k0 = np.where(m, (1 - pihat) * (1 - pihat), k0)
            k1 = np.where(m, 2 * pihat * (1 - pihat), k1)
    np.fill_diagonal(k0, 0.0)
    np.fill_diagonal(k1, 0.0)
    return k0, k1


def to_packed_upper(m):
    """Row-packed upper triangle: idx(r,c) = c + r(2N-r-1)/2, r <= c."""
    n = m.shape[0]
    return m[np.triu_indices(n)]


def merge_grm(grms, weights):
    """snpgdsMergeGRM weighted merge (src/genPCA.cpp:1834-1855)."""
    w = np.asarray(weights, dtype=np.float64)
    w = w / w.sum()
    return sum(wi * g for wi, g in zip(w, grms))


def merge_grm_indivbeta(grms, avg_vals, weights):
    """gnrGRMMerge, ":method = IndivBeta" branch (src/genPCA.cpp:1737-1822): undo each
    file's min-based normalisation through its off-diagonal mean M_b and stored
    avg_val, average, then re-normalise against the new minimum.
    Returns (grm, grm_avg_value)."""
    n = grms[0].shape[0]
    off = ~np.eye(n, dtype=bool)
    acc = np.zeros((n, n))
    for g, av, w in zip(grms, avg_vals, weights):
        mb = g[off].sum() / (float(n) * (n - 1)) * 0.5
        inv = 1 / (1 - mb)
        m = (g * 0.5 - mb) * inv * (1 - av) + av
        d = (np.diag(g) - 1 - mb) * inv * (1 - av) + av
        m[np.arange(n), np.arange(n)] = d
        acc += m * w
    avg = acc[off].sum() / (float(n) * (n - 1))
    mn = acc.min()
    out = (acc - mn) * (2 / (1 - mn))
    out[np.arange(n), np.arange(n)] = np.diag(out) * 0.5 + 1
    return out, avg


# ---------------------------------------------------------------------------
# Entries of the covariance-type matrices for a SUBSET of the samples, given the per-SNP
# statistics over ALL samples (full-size parity checks: the sub-matrix of scattered samples
# costs O(k^2 M) instead of O(N^2 M)).  Same formulas as cov_eigenstrat / grm_gcta above.
# ---------------------------------------------------------------------------
def subset_entries(sub, afreq, method="GCTA", n_total=None, trace=None):
    """sub: uint8 [nsnp, k] genotypes of k samples of a larger data set; afreq: [nsnp]
    allele frequency of every SNP over ALL samples (sum / (2 num), src/genPCA.cpp:84-142).
    GCTA: G_ij = sum_l z_il z_jl / (2 (nLocus - Denom_ij)) (src/genPCA.cpp:1201-1236);
    Eigenstrat: cov_ij (n_total - 1) / trace with the all-sample trace (src/genPCA.cpp:1381-1390)."""
    af = np.asarray(afreq, dtype=np.float64)
    mu = 2.0 * af
    poly = (af > 0) & (af < 1)
    w = np.where(poly, 1.0 / np.where(poly, af * (1 - af), 1.0), 0.0)
    valid = sub <= 2
    z = np.where(valid, (sub - mu[:, None]) * np.sqrt(w)[:, None], 0.0)
    cov = _zzt(z.T.copy()) if z.shape[1] > 64 else z.T @ z
    if method == "cov":              # un-normalised Z Z^T: shards of a multi-GPU run add these
        return cov
    if method == "Eigenstrat":
        return cov * ((n_total - 1) / trace)
    if method != "GCTA":
        raise ValueError(method)
    mm = (~valid).astype(np.float64)
    miss = mm * poly[:, None]
    den = miss.sum(0)[:, None] + miss.sum(0)[None, :] - miss.T @ mm
    return cov / (2.0 * (poly.sum() - den))


def scattered_samples(n, k, seed=1):
    """k distinct sample indices spanning the first, a middle and the last 256-sample tile row
    (plus the very first and last sample), sorted."""
    rng = np.random.default_rng(seed)
    pick = {0, n - 1}
    zones = [(0, min(256, n)), (max(0, n // 2 - 128), min(n, n // 2 + 128)), (max(0, n - 256), n)]
    while len(pick) < min(k, n):
        lo, hi = zones[len(pick) % 3] if len(pick) < k - k // 4 else (0, n)
        pick.add(int(rng.integers(lo, hi)))
    return np.array(sorted(pick), dtype=np.int64)


# ---------------------------------------------------------------------------
# Synthetic genotypes: counter-based, any shard is
# reproducible without communication.  The CUDA generator uses the same mixer.
# ---------------------------------------------------------------------------

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def synth_geno(nsamp, nsnp, seed=20261017, maf_lo=0.05, maf_hi=0.5,
               miss_rate=0.005, snp_start=0, samples=None):
    """uint8 [nsnp, nsamp] with p_l ~ U(maf_lo, maf_hi), g ~ Binomial(2, p_l),
    missing (code 3) with probability miss_rate.  Bit-identical to the device
    generator in snprelate_b200/csrc (same 64-bit mixer, same thresholds).
    `samples`: optional array of sample indices -- only those columns of the
    (arbitrarily large) data set are generated, in the given order (the generator is
    counter-based per (SNP, sample), so scattered samples cost nothing)."""
    with np.errstate(over="ignore"):
        l = (np.arange(nsnp, dtype=np.uint64) + np.uint64(snp_start))
        seed = np.uint64(seed)
        hp = _splitmix64(seed ^ (l * np.uint64(0xD1342543DE82EF95)))
        u = (hp >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        p = maf_lo + (maf_hi - maf_lo) * u
        t0 = (1 - p) ** 2
        t1 = t0 + 2 * p * (1 - p)
        th0 = np.minimum(np.floor(t0 * 4294967296.0), 4294967295.0).astype(np.uint64)
        th1 = np.minimum(np.floor(t1 * 4294967296.0), 4294967295.0).astype(np.uint64)
        thm = np.uint64(min(int(miss_rate * 4294967296.0), 4294967295))
        i = np.arange(nsamp, dtype=np.uint64) if samples is None else np.asarray(samples, dtype=np.uint64)
        key = _splitmix64((hp[:, None] + i[None, :] * np.uint64(0x9E3779B97F4A7C15)) & _M64)
        r = key >> np.uint64(32)
        rm = key & np.uint64(0xFFFFFFFF)
        g = (r >= th0[:, None]).astype(np.uint8) + (r >= th1[:, None]).astype(np.uint8)
        g = np.where(rm < thm, np.uint8(3), g).astype(np.uint8)
    return g


def check_pair_rows(est_name, idx, n_snp, kept_rows, seed=20261017, miss_rate=0.005):
    """Checker for snprelate_b200.configs.run_pair_config: `kept_rows` = [(a, [values (idx[a], idx[a:]) of
    each result matrix])] against ibs_ave / king_robust of the scattered samples `idx` over ALL n_snp SNPs of
    the synthetic data set.  Returns (max abs error, entries checked).  The results are exact rational
    functions of exact integer counters, so the error is float64 rounding only."""
    sub = synth_geno(0, n_snp, seed=seed, miss_rate=miss_rate, samples=idx)
    refs = [ibs_ave(ibs_counts(sub))] if est_name == "ibs" else list(king_robust(king_robust_counts(sub)))
    err, cnt = 0.0, 0
    for a, vals in kept_rows:
        for v, ref in zip(vals, refs):
            d = np.abs(v - ref[a, a:])
            if d.size:
                err = max(err, float(np.nanmax(d)))
            cnt += d.size
    return err, cnt
