"""Host-side mirror of the reference's R interface for the relatedness path.

Function names, argument names (``.`` -> ``_``), defaults, selection semantics and
error messages follow ``R/IBD.R``, ``R/PCA.R``, ``R/IBS.R`` and ``.InitFile2``
(``R/Internal.R:166-484``); results come back as dicts with the reference's list
element names (``sample.id``, ``snp.id``, ``grm``, ``eigenval`` ...).  All compute
runs in libsnprel_b200.so; this module only selects samples / SNPs, streams
genotype blocks to the device the way ``CGenoReadBySNP`` would
(``src/dGenGWAS.cpp:1218-1397``) and shapes the outputs.
"""
from __future__ import annotations

import math

import numpy as np

from ._lib import Context, SNPRelError, EST_IBS, EST_KING_ROBUST, EST_BETA, NA_INT

_BLOCK_BYTES = 64 << 20   # host block size for push_u8 (the reference reads cache-sized blocks)


class GenotypeData:
    """Stand-in for an open SNP GDS file object (``snpgdsOpen``).

    ``genotype`` is uint8 ``[n_snp, n_samp]`` (SNP-major, sample fastest, the
    ``snp.order``-free layout ``snpRead`` produces; 0/1/2, >2 missing) or, with
    ``packed_2bit=True``, GDS dBit2 rows ``[n_snp, ceil(n_samp/4)]``.
    """

    def __init__(self, genotype, sample_id=None, snp_id=None, chromosome=None,
                 packed_2bit=False, n_samp=None):
        g = np.asarray(genotype)
        if g.dtype != np.uint8 or g.ndim != 2:
            raise SNPRelError("genotype must be a 2-D uint8 array [n_snp, n_samp]")
        self.packed = bool(packed_2bit)
        self.genotype = g
        self.n_snp = g.shape[0]
        if self.packed:
            if n_samp is None:
                raise SNPRelError("n_samp is required for 2-bit packed genotypes")
            self.n_samp = int(n_samp)
        else:
            self.n_samp = g.shape[1]
        self.sample_id = np.arange(1, self.n_samp + 1) if sample_id is None else np.asarray(sample_id)
        self.snp_id = np.arange(1, self.n_snp + 1) if snp_id is None else np.asarray(snp_id)
        self.chromosome = None if chromosome is None else np.asarray(chromosome)
        if len(self.sample_id) != self.n_samp or len(self.snp_id) != self.n_snp:
            raise SNPRelError("sample_id / snp_id length does not match the genotype matrix")

    def block_u8(self, snp_idx, samp_mask):
        """uint8 block [len(snp_idx), n_selected_samples]."""
        if self.packed:
            rows = self.genotype[snp_idx]
            g = np.stack([(rows >> (2 * k)) & 3 for k in range(4)], axis=-1)
            g = g.reshape(rows.shape[0], -1)[:, :self.n_samp]
        else:
            g = self.genotype[snp_idx]
        if samp_mask is not None:
            g = g[:, samp_mask]
        return np.ascontiguousarray(g, dtype=np.uint8)


def _init_file2(gdsobj, sample_id, snp_id, autosome_only, remove_monosnp, maf, missing_rate,
                num_thread, verbose, device=0, ctx=None, allele_freq=None):
    """.InitFile2 (R/Internal.R:166-484): select, load the workspace, QC-filter."""
    if not isinstance(gdsobj, GenotypeData):
        raise SNPRelError("'gdsobj' should be a GenotypeData object")
    if not (isinstance(num_thread, (int, np.integer)) and num_thread >= 1):
        raise SNPRelError("`num.thread' should be a positive value or NA.")
    # samples: membership, file order kept (R/Internal.R:292-303)
    samp_mask = None
    sample_ids = gdsobj.sample_id
    if sample_id is not None:
        sample_id = np.asarray(sample_id)
        samp_mask = np.isin(gdsobj.sample_id, sample_id)
        if samp_mask.sum() != len(sample_id):
            raise SNPRelError("Some of sample.id do not exist!")
        if samp_mask.sum() <= 0:
            raise SNPRelError("No sample in the working dataset.")
        sample_ids = gdsobj.sample_id[samp_mask]
    # SNPs (R/Internal.R:316-422)
    snp_mask = np.ones(gdsobj.n_snp, dtype=bool)
    if allele_freq is not None:
        allele_freq = np.asarray(allele_freq)
        if allele_freq.ndim != 1 or not np.issubdtype(allele_freq.dtype, np.number):
            raise SNPRelError("'allele.freq' should be a numeric vector or NULL.")
        allele_freq = allele_freq.astype(np.float64)
        if snp_id is not None:
            if len(allele_freq) != len(snp_id):
                raise SNPRelError("'length(allele.freq)' should be 'length(snp.id)'.")
            # re-order to file order of the selected SNPs (R/Internal.R:355-356)
            pos = {v: k for k, v in enumerate(np.asarray(snp_id).tolist())}
        elif len(allele_freq) != gdsobj.n_snp:
            raise SNPRelError("'length(allele.freq)' should be the number of SNPs.")
    if snp_id is not None:
        snp_id = np.asarray(snp_id)
        snp_mask = np.isin(gdsobj.snp_id, snp_id)
        if snp_mask.sum() != len(snp_id):
            raise SNPRelError("Some of snp.id do not exist!")
        if snp_mask.sum() <= 0:
            raise SNPRelError("No SNP in the working dataset.")
    if autosome_only is not False and gdsobj.chromosome is not None:
        if autosome_only is True:
            snp_mask &= (gdsobj.chromosome >= 1) & (gdsobj.chromosome <= 22)   # snpgdsOption defaults
        else:
            snp_mask &= (gdsobj.chromosome == autosome_only)
    snp_idx = np.nonzero(snp_mask)[0]
    snp_ids = gdsobj.snp_id[snp_idx]
    if allele_freq is not None:
        if snp_id is not None:
            allele_freq = allele_freq[[pos[v] for v in snp_ids.tolist()]]
        else:
            allele_freq = allele_freq[snp_idx]

    # gnrSetGenoSpace: stream blocks to the device
    ctx = ctx or Context(device)
    n_samp = int(samp_mask.sum()) if samp_mask is not None else gdsobj.n_samp
    ctx.geno_begin(n_samp, len(snp_idx))
    rows_per_block = max(1, _BLOCK_BYTES // max(n_samp, 1))
    for s in range(0, len(snp_idx), rows_per_block):
        ctx.geno_push_u8(gdsobj.block_u8(snp_idx[s:s + rows_per_block], samp_mask))

    # gnrSelSNP_Base (R/Internal.R:434-462)
    if remove_monosnp or math.isfinite(maf) or math.isfinite(missing_rate):
        t_maf = maf if math.isfinite(maf) else -1.0
        t_mr = missing_rate if math.isfinite(missing_rate) else 2.0
        sel, nrm = ctx.select_snp_base(remove_monosnp, t_maf, t_mr, allele_freq)
        snp_ids = snp_ids[sel]
        if allele_freq is not None:
            allele_freq = allele_freq[sel]
        if verbose:
            print(f"Excluding {nrm} SNP{'s' if nrm != 1 else ''} (monomorphic: {remove_monosnp}, "
                  f"MAF: {maf}, missing rate: {missing_rate})")
    n, m = ctx.geno_dim()
    if verbose:
        print(f"    # of samples: {n}\n    # of SNPs: {m}")
    return dict(ctx=ctx, sample_id=sample_ids, snp_id=snp_ids, n_snp=m, n_samp=n, allele_freq=allele_freq)


def _newmat(n, packed_values):
    """R/Internal.R:46-51 wraps the packed vector into Matrix::dspMatrix(uplo='L');
    here the packed row-upper vector is returned as-is with its dimension."""
    return dict(n=n, x=packed_values, uplo="L")


def snpgdsGRM(gdsobj, sample_id=None, snp_id=None, autosome_only=True, remove_monosnp=True,
              maf=float("nan"), missing_rate=0.01,
              method="GCTA", num_thread=1, useMatrix=False, out_fn=None, out_prec="double",
              with_id=True, verbose=False, device=0, window_rows=None):
    """R/IBD.R:543-615 -> gnrGRM (src/genPCA.cpp:1614-1717).  With out_fn the matrix is
    streamed to a SNPRELATE_OUTPUT container (grmfile.py; grm_save_to_gds,
    src/genPCA.cpp:1571-1584) band by band and nothing is returned."""
    methods = ("GCTA", "Eigenstrat", "EIGMIX", "Weighted", "Corr", "IndivBeta")
    if method not in methods:
        raise SNPRelError("'arg' should be one of " + ", ".join(f'"{m}"' for m in methods))
    mtxt = method
    if method == "Weighted":          # R/IBD.R:552-555
        method = "EIGMIX"
    ws = _init_file2(gdsobj, sample_id, snp_id, autosome_only, remove_monosnp, maf, missing_rate,
                     num_thread, verbose, device)
    windowed = method in ("GCTA", "EIGMIX", "Eigenstrat")
    if out_fn is not None:
        if not isinstance(out_fn, str):
            raise SNPRelError("'out.fn' should be a file name")
        from .grmfile import GrmWriter
        writer = GrmWriter(out_fn, ["snpgdsGRM", f":method = {method}"], ws["sample_id"], ws["snp_id"], out_prec)
        with ws["ctx"] as ctx:
            rows = window_rows or (ctx.auto_window_rows(16) if windowed else 0)
            avg = None
            if rows and windowed:
                try:
                    for r0, h in ctx.windows(rows):
                        ctx.set_row_window(r0, h)
                        writer.write_band(r0, ctx.grm(method, packed=True)[0])
                finally:
                    ctx.set_row_window(0, 0)
            else:
                grm, avg = ctx.grm(method, packed=False)
                writer.write_full(grm)
        writer.close(avg if method == "IndivBeta" else None)
        return None
    with ws["ctx"] as ctx:
        rows = (window_rows or ctx.auto_window_rows(16)) if (useMatrix and windowed) else 0
        if rows:      # N^2 int64 planes do not fit: walk row windows, packed slices concatenated
            grm, avg = ctx.packed_by_windows(lambda: ctx.grm(method, packed=True)[0], rows), 0.0
        else:
            grm, avg = ctx.grm(method, packed=bool(useMatrix))
    if useMatrix and method != "Corr":
        grm = _newmat(ws["n_samp"], grm)
    if not with_id:
        return grm
    rv = {"sample.id": ws["sample_id"], "snp.id": ws["snp_id"], "method": method, "grm": grm}   # the mapped method, R/IBD.R:603-604
    if method == "IndivBeta":
        rv["avg_val"] = avg           # R/IBD.R:605-606
    return rv


def snpgdsMergeGRM(filelist, out_fn=None, out_prec="double", weight=None, verbose=False):
    """R/IBD.R:624-748 -> gnrGRMMerge (src/genPCA.cpp:1721-1857) over the containers written
    by snpgdsGRM(out_fn=...)."""
    from .grmfile import merge_grm_files
    return merge_grm_files(filelist, out_fn, out_prec, weight, verbose)


def snpgdsPCA(gdsobj, sample_id=None, snp_id=None, autosome_only=True, remove_monosnp=True,
              maf=float("nan"), missing_rate=0.01, algorithm="exact", eigen_cnt=None, num_thread=1,
              bayesian=False, need_genmat=False, genmat_only=False, eigen_method="DSPEVX",
              aux_dim=None, iter_num=10, aux_mat=None, verbose=False, device=0):
    """R/PCA.R:22-91 -> gnrPCA "exact" / "randomized" (src/genPCA.cpp:1355-1452).  eigen_cnt
    defaults to 32 (exact) or 16 (randomized) when None; aux_dim to 2 * eigen_cnt.  `aux_mat`
    stands in for R's rnorm(aux.dim * n.samp) (R/PCA.R:56): pass it for a reproducible run."""
    if algorithm not in ("exact", "randomized"):
        raise SNPRelError("'arg' should be one of \"exact\", \"randomized\"")      # match.arg
    if eigen_cnt is None:
        eigen_cnt = 16 if algorithm == "randomized" else 32
    if eigen_method not in ("DSPEVX", "DSPEV"):
        raise SNPRelError("Unknown 'eigen.method'.")
    ws = _init_file2(gdsobj, sample_id, snp_id, autosome_only, remove_monosnp, maf, missing_rate,
                     num_thread, verbose, device)
    if genmat_only:
        need_genmat = True
    if eigen_cnt <= 0:
        eigen_cnt = ws["n_samp"]
    if algorithm == "randomized":
        n = ws["n_samp"]
        if aux_dim is None:
            aux_dim = 2 * eigen_cnt
        if aux_mat is None:
            aux_mat = np.random.default_rng().standard_normal(aux_dim * n)
        with ws["ctx"] as ctx:
            d, h, tr2 = ctx.pca_randomized(aux_mat, aux_dim, iter_num)
        vp = 2.0 * d ** 2 / tr2                                            # R/PCA.R:82-87
        return {"sample.id": ws["sample_id"], "snp.id": ws["snp_id"], "eigenval": (n - 1) * vp,
                "eigenvect": np.ascontiguousarray(h[:eigen_cnt].T), "varprop": vp, "TraceXTX": tr2,
                "Bayesian": False, "class": "snpgdsPCAClass"}
    with ws["ctx"] as ctx:
        r = ctx.pca(eigen_cnt, bayesian, need_genmat, genmat_only)
    eigenval = r["eigenval"]
    return {"sample.id": ws["sample_id"], "snp.id": ws["snp_id"], "eigenval": eigenval,
            "eigenvect": r["eigenvect"],
            "varprop": None if eigenval is None else eigenval / r["TraceVal"],
            "TraceXTX": r["TraceXTX"], "Bayesian": bayesian, "genmat": r["genmat"],
            "class": "snpgdsPCAClass"}


def snpgdsEIGMIX(gdsobj, sample_id=None, snp_id=None, autosome_only=True, remove_monosnp=True,
                 maf=float("nan"), missing_rate=0.01, num_thread=1, eigen_cnt=32, diagadj=True,
                 ibdmat=False, verbose=False, device=0):
    """R/PCA.R:311-338 -> gnrEigMix (src/genEIGMIX.cpp:656-735)."""
    ws = _init_file2(gdsobj, sample_id, snp_id, autosome_only, remove_monosnp, maf, missing_rate,
                     num_thread, verbose, device)
    if eigen_cnt < 0:
        eigen_cnt = ws["n_samp"]
    with ws["ctx"] as ctx:
        r = ctx.eigmix(eigen_cnt, diagadj, ibdmat)
    return {"sample.id": ws["sample_id"], "snp.id": ws["snp_id"], "eigenval": r["eigenval"],
            "eigenvect": r["eigenvect"], "afreq": r["afreq"], "ibd": r["ibd"], "diagadj": diagadj,
            "class": "snpgdsEigMixClass"}


def _init_file(gdsobj, sample_id, snp_id, device=0):
    """.InitFile (R/Internal.R:56-163): selection only, no QC filter."""
    return _init_file2(gdsobj, sample_id, snp_id, False, False, float("nan"), float("nan"), 1, False, device)


def snpgdsPCACorr(pcaobj, gdsobj, snp_id=None, eig_which=None, num_thread=1, with_id=True, outgds=None,
                  verbose=False, device=0):
    """R/PCA.R:99-175 -> gnrPCACorr (src/genPCA.cpp:1456-1485).  `eig_which` is 1-based like
    the reference's eig.which.  pcaobj: a snpgdsPCA / snpgdsEIGMIX result, or a
    (sample_id, eigenvect) pair standing in for a matrix with sample-id row names."""
    if isinstance(pcaobj, dict) and pcaobj.get("class") in ("snpgdsPCAClass", "snpgdsEigMixClass"):
        sampid, eigenvect = pcaobj["sample.id"], pcaobj["eigenvect"]
    elif isinstance(pcaobj, (tuple, list)) and len(pcaobj) == 2:
        sampid, eigenvect = pcaobj
        if sampid is None:
            raise SNPRelError("'rownames(pcaobj)' should be sample id.")
    else:
        raise SNPRelError("'pcaobj' should be a snpgdsPCA / snpgdsEIGMIX result or (sample.id, eigenvect)")
    if outgds is not None:
        raise SNPRelError("'outgds' (packedreal16 GDS output) is not supported; use the returned matrix")
    eigenvect = np.asarray(eigenvect, dtype=np.float64)
    ws = _init_file(gdsobj, sampid, snp_id, device)
    if len(sampid) != eigenvect.shape[0]:
        raise SNPRelError("Internal error: the number of samples should be equal to the number of rows in 'eigenvect'.")
    if eig_which is None:
        idx = np.arange(eigenvect.shape[1])
    else:
        idx = np.asarray(eig_which, dtype=np.int64).reshape(-1) - 1
        if idx.size == 0 or idx.min() < 0 or idx.max() >= eigenvect.shape[1]:
            raise SNPRelError("'eig.which' is out of range")
    with ws["ctx"] as ctx:
        corr = ctx.pca_corr(eigenvect[:, idx])
    if with_id:
        return {"sample.id": ws["sample_id"], "snp.id": ws["snp_id"], "snpcorr": corr}
    return corr


def snpgdsPCASNPLoading(pcaobj, gdsobj, num_thread=1, verbose=False, device=0):
    """R/PCA.R:184-229 -> gnrPCASNPLoading (src/genPCA.cpp:1489-1540) or, for a snpgdsEIGMIX
    result, gnrEigMixSNPLoading (src/genEIGMIX.cpp:739-775)."""
    cls = pcaobj.get("class") if isinstance(pcaobj, dict) else None
    if cls not in ("snpgdsPCAClass", "snpgdsEigMixClass"):
        raise SNPRelError("'pcaobj' should be a snpgdsPCA or snpgdsEIGMIX result")
    if pcaobj.get("eigenval") is None or pcaobj.get("eigenvect") is None:
        raise SNPRelError("'pcaobj' has no eigenvalues / eigenvectors")
    ws = _init_file(gdsobj, pcaobj["sample.id"], pcaobj["snp.id"], device)
    k = pcaobj["eigenvect"].shape[1]
    with ws["ctx"] as ctx:
        if cls == "snpgdsPCAClass":
            load, avg, scale = ctx.pca_snp_loading(pcaobj["eigenval"], pcaobj["eigenvect"], pcaobj["TraceXTX"],
                                                   pcaobj["Bayesian"])
            return {"sample.id": pcaobj["sample.id"], "snp.id": pcaobj["snp.id"], "eigenval": pcaobj["eigenval"],
                    "snploading": load, "TraceXTX": pcaobj["TraceXTX"], "Bayesian": pcaobj["Bayesian"],
                    "avgfreq": avg, "scale": scale, "class": "snpgdsPCASNPLoadingClass"}
        if pcaobj.get("diagadj"):
            raise SNPRelError("Please run `snpgdsEIGMIX(, diagadj=FALSE)` for projecting new samples.")
        load = ctx.eigmix_snp_loading(pcaobj["eigenval"][:k], pcaobj["eigenvect"], pcaobj["afreq"])
        return {"sample.id": pcaobj["sample.id"], "snp.id": pcaobj["snp.id"], "eigenval": pcaobj["eigenval"],
                "snploading": load, "afreq": pcaobj["afreq"], "class": "snpgdsEigMixSNPLoadingClass"}


def snpgdsPCASampLoading(loadobj, gdsobj, sample_id=None, num_thread=1, verbose=False, device=0):
    """R/PCA.R:238-303 -> gnrPCASampLoading (src/genPCA.cpp:1542-1563) / gnrEigMixSampLoading
    (src/genEIGMIX.cpp:777-803): project (new) samples onto the components of `loadobj`."""
    cls = loadobj.get("class") if isinstance(loadobj, dict) else None
    if cls not in ("snpgdsPCASNPLoadingClass", "snpgdsEigMixSNPLoadingClass"):
        raise SNPRelError("'loadobj' should be a snpgdsPCASNPLoading result")
    ws = _init_file(gdsobj, sample_id, loadobj["snp.id"], device)
    load = np.asarray(loadobj["snploading"], dtype=np.float64)
    eigcnt = load.shape[0]
    nan = np.full(ws["n_samp"], np.nan)
    with ws["ctx"] as ctx:
        if cls == "snpgdsPCASNPLoadingClass":
            ss = (len(loadobj["sample.id"]) - 1) / loadobj["TraceXTX"]              # R/PCA.R:274-277
            sload = load * np.sqrt(ss / np.asarray(loadobj["eigenval"][:eigcnt]))[:, None]
            mm = ctx.pca_samp_loading(sload, loadobj["avgfreq"], loadobj["scale"])
            return {"sample.id": ws["sample_id"], "snp.id": loadobj["snp.id"], "eigenval": nan, "eigenvect": mm,
                    "varprop": nan.copy(), "TraceXTX": loadobj["TraceXTX"], "Bayesian": loadobj["Bayesian"],
                    "genmat": None, "class": "snpgdsPCAClass"}
        sload = load * np.sqrt(1.0 / np.asarray(loadobj["eigenval"][:eigcnt]))[:, None]       # R/PCA.R:296-297
        mm = ctx.eigmix_samp_loading(sload, loadobj["afreq"])
        return {"sample.id": ws["sample_id"], "snp.id": loadobj["snp.id"], "eigenval": nan, "eigenvect": mm,
                "afreq": loadobj["afreq"], "class": "snpgdsEigMixClass"}


def snpgdsIBS(gdsobj, sample_id=None, snp_id=None, autosome_only=True, remove_monosnp=True,
              maf=float("nan"), missing_rate=0.01, num_thread=1, useMatrix=False, verbose=False,
              device=0):
    """R/IBS.R:22-46 -> gnrIBSAve (src/genIBS.cpp:441-497)."""
    ws = _init_file2(gdsobj, sample_id, snp_id, autosome_only, remove_monosnp, maf, missing_rate,
                     num_thread, verbose, device)
    with ws["ctx"] as ctx:
        rows = ctx.auto_window_rows(12) if useMatrix else 0      # TIBS: 12 bytes per pair
        ibs = (ctx.packed_by_windows(lambda: ctx.ibs_ave(packed=True), rows) if rows
               else ctx.ibs_ave(packed=bool(useMatrix)))
    if useMatrix:
        ibs = _newmat(ws["n_samp"], ibs)
    return {"sample.id": ws["sample_id"], "snp.id": ws["snp_id"], "ibs": ibs}


def snpgdsIBDMoM(gdsobj, sample_id=None, snp_id=None, autosome_only=True, remove_monosnp=True,
                 maf=float("nan"), missing_rate=0.01, allele_freq=None, kinship=False,
                 kinship_constraint=False, num_thread=1, useMatrix=False, verbose=False, device=0):
    """R/IBD.R:22-70 -> gnrIBD_PLINK (src/genIBS.cpp:558-639): PLINK method of moments."""
    ws = _init_file2(gdsobj, sample_id, snp_id, autosome_only, remove_monosnp, maf, missing_rate,
                     num_thread, verbose, device, allele_freq=allele_freq)
    for name, v in (("kinship", kinship), ("kinship.constraint", kinship_constraint), ("useMatrix", useMatrix)):
        if not isinstance(v, (bool, np.bool_)):
            raise SNPRelError(f"'{name}' should be a logical value.")
    with ws["ctx"] as ctx:
        rows = ctx.auto_window_rows(12) if useMatrix else 0
        if rows:
            sums, afreq = ctx.ibd_mom_sums(ws["allele_freq"])
            k0, k1 = ctx.packed_by_windows(
                lambda: ctx.ibd_mom_from_sums(sums, kinship_constraint, packed=True), rows)
        else:
            k0, k1, afreq = ctx.ibd_mom(ws["allele_freq"], kinship_constraint, packed=bool(useMatrix))
    afreq = afreq.copy()
    afreq[afreq < 0] = np.nan
    ans = {"sample.id": ws["sample_id"], "snp.id": ws["snp_id"], "afreq": afreq, "k0": k0, "k1": k1}
    if kinship:
        ans["kinship"] = 0.5 * (1 - k0 - k1) + 0.25 * k1
    if useMatrix:
        for key in ("k0", "k1") + (("kinship",) if kinship else ()):
            ans[key] = _newmat(ws["n_samp"], ans[key])
    return ans


def snpgdsIBSNum(gdsobj, sample_id=None, snp_id=None, autosome_only=True, remove_monosnp=True,
                 maf=float("nan"), missing_rate=0.01, num_thread=1, verbose=False, device=0):
    """R/IBS.R:55-73 -> gnrIBSNum (src/genIBS.cpp:500-550)."""
    ws = _init_file2(gdsobj, sample_id, snp_id, autosome_only, remove_monosnp, maf, missing_rate,
                     num_thread, verbose, device)
    with ws["ctx"] as ctx:
        i0, i1, i2 = ctx.ibs_num()
    return {"sample.id": ws["sample_id"], "snp.id": ws["snp_id"], "ibs0": i0, "ibs1": i1, "ibs2": i2}


def _family_codes(family_id, gdsobj, sample_id, ws):
    """R/IBD.R:348-372: family.id -> integer factor codes, "" / None -> NA."""
    if family_id is None:
        return None
    family_id = list(family_id)
    if len(family_id) != ws["n_samp"]:
        raise SNPRelError("'length(family.id)' should be the number of samples.")
    if sample_id is not None:
        # exactly the reference's re-indexing, family.id[match(sample.id, ws$sample.id)] (R/IBD.R:356-357):
        # entry k becomes the family of the sample at the WORKSPACE position of sample.id[k].  (For a
        # sample.id that is not in file order this is the inverse of the permutation one would expect;
        # a drop-in replacement has to pick the same pairs as "within family", so it is mirrored as is.)
        pos = {s: k for k, s in enumerate(ws["sample_id"].tolist())}
        family_id = [family_id[pos[s]] for s in np.asarray(sample_id).tolist()]
    levels = sorted({f for f in family_id if f is not None and f != "" and f == f})
    code = {f: k + 1 for k, f in enumerate(levels)}
    return np.array([code.get(f, NA_INT) if (f is not None and f != "" and f == f) else NA_INT
                     for f in family_id], dtype=np.int32)


def snpgdsIBDKING(gdsobj, sample_id=None, snp_id=None, autosome_only=True, remove_monosnp=True,
                  maf=float("nan"), missing_rate=0.01, type="KING-robust", family_id=None,
                  num_thread=1, useMatrix=False, verbose=False, device=0):
    """R/IBD.R:333-419 -> gnrIBD_KING_Robust / gnrIBD_KING_Homo."""
    if type not in ("KING-robust", "KING-homo"):
        raise SNPRelError("Invalid 'type'.")
    ws = _init_file2(gdsobj, sample_id, snp_id, autosome_only, remove_monosnp, maf, missing_rate,
                     num_thread, verbose, device)
    fam = _family_codes(family_id, gdsobj, sample_id, ws)
    rv = {"sample.id": ws["sample_id"], "snp.id": ws["snp_id"], "afreq": None}
    with ws["ctx"] as ctx:
        if type == "KING-homo":
            a, b = ctx.king_homo(packed=bool(useMatrix))
            names = ("k0", "k1")
        else:
            rows = ctx.auto_window_rows(20) if useMatrix else 0  # TS_KINGRobust: 20 bytes per pair
            a, b = (ctx.packed_by_windows(lambda: ctx.king_robust(fam, packed=True), rows) if rows
                    else ctx.king_robust(fam, packed=bool(useMatrix)))
            names = ("IBS0", "kinship")
    if useMatrix:
        a, b = _newmat(ws["n_samp"], a), _newmat(ws["n_samp"], b)
    rv[names[0]], rv[names[1]] = a, b
    return rv


def snpgdsIndivBeta(gdsobj, sample_id=None, snp_id=None, autosome_only=True, remove_monosnp=True,
                    maf=float("nan"), missing_rate=0.01, method="weighted", inbreeding=True,
                    num_thread=1, with_id=True, useMatrix=False, verbose=False, device=0):
    """R/IBD.R:838-866 -> gnrIBD_Beta (src/genBeta.cpp:361-460)."""
    if method != "weighted":
        raise SNPRelError("'arg' should be \"weighted\"")
    if not isinstance(inbreeding, (bool, np.bool_)):
        raise SNPRelError("'inbreeding' must be TRUE or FALSE.")
    ws = _init_file2(gdsobj, sample_id, snp_id, autosome_only, remove_monosnp, maf, missing_rate,
                     num_thread, verbose, device)
    with ws["ctx"] as ctx:
        beta, avg = ctx.indiv_beta(inbreeding, packed=bool(useMatrix))
    if useMatrix:
        beta = _newmat(ws["n_samp"], beta)
    if not with_id:
        return beta
    return {"sample.id": ws["sample_id"], "snp.id": ws["snp_id"], "inbreeding": inbreeding,
            "beta": beta, "avg_val": avg}


def snpgdsSNPRateFreq(gdsobj, sample_id=None, snp_id=None, with_id=False, device=0):
    """R/AllUtilities.R snpgdsSNPRateFreq -> gnrSNPRateFreq (src/SNPRelate.cpp:243)."""
    ws = _init_file2(gdsobj, sample_id, snp_id, False, False, float("nan"), float("nan"), 1, False,
                     device)
    with ws["ctx"] as ctx:
        af, maf, mr = ctx.snp_ratefreq()
    rv = {"AlleleFreq": af, "MinorFreq": maf, "MissingRate": mr}
    if with_id:
        rv["sample.id"], rv["snp.id"] = ws["sample_id"], ws["snp_id"]
    return rv
