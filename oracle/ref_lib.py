"""ORACLE / TEST INFRASTRUCTURE ONLY.

ctypes driver for oracle/_ref/libsnprelate_ref.so: the reference's OWN hot-path
sources (src/genPCA.cpp, genEIGMIX.cpp, genIBS.cpp, genKING.cpp, genBeta.cpp,
dGenGWAS.cpp, dVect.cpp, ThreadPool.cpp) compiled unmodified by oracle/Makefile
against the shim in oracle/ref_shim.  Used to validate the numpy restatement and
as the CPU baseline (`cpu_baseline.kind = "reference"`)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_PATH = os.path.join(_DIR, "libsnprelate_ref.so")
# the product's R binding (r_shim.cpp) behind the same driver functions, see oracle/Makefile
RSHIM_PATH = os.path.join(_DIR, "libsnprelate_b200_rshim.so")
_LIBS = {}
_CURRENT = None


def available() -> bool:
    return os.path.exists(_PATH)


def rshim_available() -> bool:
    return os.path.exists(RSHIM_PATH)


def _load(path):
    if path not in _LIBS:
        lib = C.CDLL(path)
        lib.ref_error.restype = C.c_char_p
        _LIBS[path] = lib
    return _LIBS[path]


def _lib():
    return _CURRENT if _CURRENT is not None else _load(_PATH)


def _ck(rc):
    if rc != 0:
        raise RuntimeError(_lib().ref_error().decode())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RefWorkspace:
    """Drives the reference like R does: set the genotype space, optionally
    gnrSelSNP_Base, then the gnr* estimators."""

    def __init__(self, geno: np.ndarray, lib_path: str = None):
        """lib_path: the shared library whose gnr* entry points are driven (default: the reference's
        own sources; RSHIM_PATH: the product's R binding).  One library is current at a time, like
        the process-global workspace it owns."""
        global _CURRENT
        _CURRENT = _load(lib_path or _PATH)
        g = np.ascontiguousarray(geno, dtype=np.uint8)
        self.nsnp, self.nsamp = g.shape
        _ck(_lib().ref_set_geno(_p(g), C.c_int(g.shape[0]), C.c_int(g.shape[1])))

    def select_snp_base(self, remove_mono=True, maf=-1.0, missrate=2.0):
        flags = np.zeros(self.nsnp, dtype=np.uint8)
        n = C.c_int()
        _ck(_lib().ref_select_snp_base(int(remove_mono), C.c_double(maf), C.c_double(missrate), _p(flags), C.byref(n)))
        return flags.astype(bool), n.value

    def dims(self):
        a, b = C.c_int(), C.c_int()
        _ck(_lib().ref_dims(C.byref(a), C.byref(b)))
        return a.value, b.value

    def grm(self, method="GCTA", nthread=1):
        n = self.dims()[1]
        out = np.empty((n, n))
        _ck(_lib().ref_grm(method.encode(), int(nthread), _p(out)))
        return out

    def grm_gds(self, method="GCTA", nthread=1):
        """gnrGRM with a GDS output node: the float64 stream appended to "grm" (n rows of n values,
        grm_save_to_gds, src/genPCA.cpp:1571-1584) as an [n, n] array."""
        n = self.dims()[1]
        out = np.empty((n, n))
        cnt = C.c_longlong()
        _ck(_lib().ref_grm_gds(method.encode(), int(nthread), _p(out), C.byref(cnt)))
        if cnt.value != n * n:
            raise RuntimeError(f"GDS stream has {cnt.value} values, expected {n * n}")
        return out

    def pca(self, nthread=1, bayesian=False, eigen_cnt=0):
        n = self.dims()[1]
        genmat = np.empty((n, n))
        tr = C.c_double()
        ev = np.empty(n) if eigen_cnt > 0 else None
        evec = np.empty((eigen_cnt, n)) if eigen_cnt > 0 else None
        _ck(_lib().ref_pca(int(nthread), int(bayesian), int(eigen_cnt), _p(genmat), C.byref(tr), _p(ev), _p(evec)))
        return dict(genmat=genmat, TraceXTX=tr.value, eigenval=ev, eigenvect=None if evec is None else evec.T)

    def pca_randomized(self, aux_mat, aux_dim, iter_num=10, nthread=1):
        """gnrPCA algorithm "randomized" (src/genPCA.cpp:1436-1442, CRandomPCA :469-796)
        -> (sigma [nsamp], V^T [hsize, nsamp], 2 TraceXTX)."""
        n = self.dims()[1]
        hsize = aux_dim * (iter_num + 1)
        aux = np.ascontiguousarray(aux_mat, dtype=np.float64).copy()     # overwritten by the reference
        sigma, vt = np.empty(n), np.empty((n, hsize))                    # R matrix hsize x n, column major
        tr = C.c_double()
        _ck(_lib().ref_pca_randomized(int(nthread), _p(aux), int(aux_dim), int(iter_num), _p(sigma), _p(vt), C.byref(tr)))
        return sigma, vt.T.copy(), tr.value

    def eigmix(self, nthread=1, diagadj=True):
        nsnp, n = self.dims()
        ibd, af = np.empty((n, n)), np.empty(nsnp)
        _ck(_lib().ref_eigmix(int(nthread), int(diagadj), _p(ibd), _p(af)))
        return ibd, af

    def ibs_num(self, nthread=1):
        n = self.dims()[1]
        o = [np.empty((n, n), dtype=np.int32) for _ in range(3)]
        _ck(_lib().ref_ibs_num(int(nthread), _p(o[0]), _p(o[1]), _p(o[2])))
        return o

    def ibs_ave(self, nthread=1):
        n = self.dims()[1]
        o = np.empty((n, n))
        _ck(_lib().ref_ibs_ave(int(nthread), _p(o)))
        return o

    def king_robust(self, nthread=1, family=None):
        n = self.dims()[1]
        a, b = np.empty((n, n)), np.empty((n, n))
        fam = None if family is None else np.ascontiguousarray(family, dtype=np.int32)
        _ck(_lib().ref_king_robust(int(nthread), _p(fam), _p(a), _p(b)))
        return a, b

    def king_homo(self, nthread=1):
        n = self.dims()[1]
        a, b = np.empty((n, n)), np.empty((n, n))
        _ck(_lib().ref_king_homo(int(nthread), _p(a), _p(b)))
        return a, b

    def ibd_mom(self, nthread=1, allele_freq=None, kinship_constraint=False):
        """gnrIBD_PLINK (src/genIBS.cpp:558-639) -> (k0, k1, afreq)."""
        m, n = self.dims()
        k0, k1, af = np.empty((n, n)), np.empty((n, n)), np.empty(m)
        afin = None
        if allele_freq is not None:
            afin = np.ascontiguousarray(allele_freq, dtype=np.float64)
            assert afin.shape == (m,)
        _ck(_lib().ref_ibd_mom(int(nthread), _p(afin) if afin is not None else None,
                               int(kinship_constraint), _p(k0), _p(k1), _p(af)))
        return k0, k1, af

    def indiv_beta(self, nthread=1, inbreeding=True):
        n = self.dims()[1]
        o = np.empty((n, n))
        _ck(_lib().ref_indiv_beta(int(nthread), int(inbreeding), _p(o)))
        return o

    # ---- PCA / EIGMIX loadings and correlations ----
    def pca_corr(self, eigenvect, nthread=1):
        """gnrPCACorr (src/genPCA.cpp:1456-1485) -> [k, nsnp]."""
        m, n = self.dims()
        v = np.asfortranarray(eigenvect, dtype=np.float64)
        k = v.shape[1]
        out = np.empty((m, k))
        _ck(_lib().ref_pca_corr(int(nthread), k, _p(v), _p(out)))
        return out.T

    def pca_snp_loading(self, eigenval, eigenvect, trace_xtx, bayesian=False, nthread=1):
        """gnrPCASNPLoading (src/genPCA.cpp:1489-1540) -> (loading [k, nsnp], avgfreq, scale)."""
        m, n = self.dims()
        v = np.asfortranarray(eigenvect, dtype=np.float64)
        k = v.shape[1]
        ev = np.ascontiguousarray(eigenval[:k], dtype=np.float64)
        load, af, sc = np.empty((m, k)), np.empty(m), np.empty(m)
        _ck(_lib().ref_pca_snp_loading(int(nthread), k, _p(ev), _p(v), C.c_double(trace_xtx), int(bayesian),
                                       _p(load), _p(af), _p(sc)))
        return load.T, af, sc

    def pca_samp_loading(self, loadings, avgfreq, scale, nthread=1):
        """gnrPCASampLoading (src/genPCA.cpp:1542-1563): loadings [k, nsnp] -> [n, k]."""
        m, n = self.dims()
        ld = np.ascontiguousarray(np.asarray(loadings, dtype=np.float64).T)     # [nsnp][k] == k x nsnp column-major
        k = ld.shape[1]
        out = np.empty((k, n))
        _ck(_lib().ref_pca_samp_loading(int(nthread), k, _p(ld), _p(np.ascontiguousarray(avgfreq, dtype=np.float64)),
                                        _p(np.ascontiguousarray(scale, dtype=np.float64)), _p(out)))
        return out.T

    def eigmix_snp_loading(self, eigenval, eigenvect, afreq, nthread=1):
        """gnrEigMixSNPLoading (src/genEIGMIX.cpp:739-775) -> [k, nsnp]."""
        m, n = self.dims()
        v = np.asfortranarray(eigenvect, dtype=np.float64)
        k = v.shape[1]
        ev = np.ascontiguousarray(eigenval[:k], dtype=np.float64)
        load = np.empty((m, k))
        _ck(_lib().ref_eigmix_snp_loading(int(nthread), k, _p(ev), _p(v),
                                          _p(np.ascontiguousarray(afreq, dtype=np.float64)), _p(load)))
        return load.T

    def eigmix_samp_loading(self, loadings, afreq, nthread=1):
        """gnrEigMixSampLoading (src/genEIGMIX.cpp:777-803): loadings [k, nsnp] -> [n, k]."""
        m, n = self.dims()
        ld = np.ascontiguousarray(np.asarray(loadings, dtype=np.float64).T)
        k = ld.shape[1]
        out = np.empty((k, n))
        _ck(_lib().ref_eigmix_samp_loading(int(nthread), k, _p(ld), _p(np.ascontiguousarray(afreq, dtype=np.float64)),
                                           _p(out)))
        return out.T
