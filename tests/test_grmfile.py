"""SNPRELATE_OUTPUT container + snpgdsMergeGRM host logic (no GPU): streamed band writes
reproduce the matrix, and merging per-SNP-set GRMs reproduces the GRM of the union
(inst/unitTests/test_GRM.R:15-90) for the plain and the IndivBeta transforms."""
import numpy as np
import pytest

from oracle import snprel_oracle as O
from snprelate_b200 import grmfile as F


def _write(path, mat, snp_id, method, avg=None, band=None, prec="double"):
    n = mat.shape[0]
    w = F.GrmWriter(str(path), ["snpgdsGRM", f":method = {method}"], np.array([f"s{i}" for i in range(n)]),
                    snp_id, prec)
    if band:
        packed = O.to_packed_upper(mat)
        off = 0
        for r0 in range(0, n, band):
            h = min(band, n - r0)
            cnt = sum(n - r for r in range(r0, r0 + h))
            w.write_band(r0, packed[off:off + cnt])
            off += cnt
    else:
        w.write_full(mat)
    w.close(avg)


def test_band_writes_and_readback(tmp_path):
    rng = np.random.default_rng(3)
    a = rng.normal(size=(37, 37))
    a = a + a.T
    _write(tmp_path / "a.grm", a, np.arange(5), "GCTA", band=8)
    _write(tmp_path / "b.grm", a, np.arange(5), "GCTA")
    _write(tmp_path / "c.grm", a, np.arange(5), "GCTA", band=16, prec="single")
    fa, fb, fc = (F.GrmFile(str(tmp_path / k)) for k in ("a.grm", "b.grm", "c.grm"))
    assert np.array_equal(fa.grm, a) and np.array_equal(fb.grm, a)
    assert fc.grm.dtype == np.float32 and np.allclose(fc.grm, a, rtol=1e-6)
    assert fa.command == ["snpgdsGRM", ":method = GCTA"] and np.isnan(fa.avg_val)
    assert list(fa.sample_id[:2]) == ["s0", "s1"] and np.array_equal(fa.snp_id, np.arange(5))
    (tmp_path / "bad").write_bytes(b"not a grm file at all, really")
    with pytest.raises(F.GrmFileError, match="not valid"):
        F.GrmFile(str(tmp_path / "bad"))


def test_merge_identity_gcta_and_beta(tmp_path):
    g = O.synth_geno(60, 3000, seed=4, miss_rate=0.0)
    g = g[O.select_snp_base(g)]
    m = g.shape[0]
    parts = [np.arange(0, 700), np.arange(700, 1900), np.arange(1900, m)]
    for method in ("GCTA", "IndivBeta"):
        names = []
        for k, idx in enumerate(parts):
            if method == "GCTA":
                mat, avg = O.grm_gcta(g[idx]), None
            else:
                mat, avg = O.grm_indivbeta(O.beta_counts(g[idx]))
            names.append(str(tmp_path / f"{method}{k}.grm"))
            _write(names[-1], mat, idx, method, avg, band=25)
        whole = O.grm_gcta(g) if method == "GCTA" else O.grm_indivbeta(O.beta_counts(g))[0]
        rv = F.merge_grm_files(names)
        assert np.max(np.abs(rv["grm"] - whole)) < 1e-12, method
        assert np.array_equal(rv["snp.id"], np.arange(m))
        out = str(tmp_path / f"{method}_merged.grm")
        assert F.merge_grm_files(names, out) is None
        fm = F.GrmFile(out)
        assert np.array_equal(fm.grm, rv["grm"]) and fm.command[1] == f":method = {method}"
        if method == "IndivBeta":
            w = np.array([len(p) for p in parts], dtype=float)
            grms = [np.array(F.GrmFile(nm).grm) for nm in names]
            avgs = [F.GrmFile(nm).avg_val for nm in names]
            ref, ravg = O.merge_grm_indivbeta(grms, avgs, w / w.sum())
            assert np.max(np.abs(rv["grm"] - ref)) < 1e-13 and abs(rv["avg_val"] - ravg) < 1e-14
            assert abs(fm.avg_val - ravg) < 1e-14
        else:
            w = np.array([len(p) for p in parts], dtype=float)
            ref = O.merge_grm([O.grm_gcta(g[idx]) for idx in parts], w)
            assert np.max(np.abs(rv["grm"] - ref)) < 1e-14
    # logical weights subtract a SNP set again (R/IBD.R:683-690)
    allf = str(tmp_path / "GCTA_merged.grm")
    a = F.merge_grm_files([allf, str(tmp_path / "GCTA0.grm")], weight=[True, False])
    assert np.array_equal(a["snp.id"], np.arange(700, m))
    assert np.max(np.abs(a["grm"] - O.grm_gcta(g[700:]))) < 1e-12
    with pytest.raises(F.GrmFileError, match="different command"):
        F.merge_grm_files([names[0], str(tmp_path / "GCTA0.grm")])
