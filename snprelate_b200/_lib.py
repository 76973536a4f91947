"""ctypes binding of libsnprel_b200.so (the C ABI declared in include/snprel_b200.h).

This is the same binding an R ``.Call`` shim would use (INTEGRATION.md shows it);
Python is only the test / benchmark host here because the image has no R.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GRM_METHODS = {"Eigenstrat": 0, "GCTA": 1, "Corr": 2, "EIGMIX": 3, "IndivBeta": 4}
EST_IBS, EST_KING_ROBUST, EST_BETA, EST_KING_HOMO = 10, 11, 12, 13
NA_INT = -2147483648


class SNPRelError(RuntimeError):
    """Raised for every failure reported by the CUDA library (the analogue of the
    R error raised through COREARRAY_CATCH / gnrErrMsg)."""


class Plan(C.Structure):
    _fields_ = [("max_abs", C.c_double), ("max_abs_w", C.c_double), ("sum_bound", C.c_double),
                ("err_weight", C.c_double), ("scale", C.c_double), ("tol", C.c_double),
                ("total_missing", C.c_int64), ("max_missing", C.c_int64), ("n_snp", C.c_int64),
                ("frac_bits", C.c_int32), ("frac_bits_w", C.c_int32), ("frac_bits_d", C.c_int32),
                ("digits", C.c_int32), ("digits_w", C.c_int32), ("digits_d", C.c_int32),
                ("bayesian", C.c_int32), ("frac_bits_v", C.c_int32), ("rounding", C.c_int32),
                ("diag_bound", C.c_double), ("sum_rest", C.c_double), ("err_weight2", C.c_double)]


ROUNDING_MODES = {"nearest": 0, "random": 1, "auto": 2}


def plan_format(est, plan, mode="auto", n_samp=0):
    """Host-only: the fixed-point format the library chooses for (merged) plan statistics; fills
    frac_bits / digits / rounding of `plan` and returns the number of tensor passes."""
    lib = load_library()
    rc = lib.snprel_plan_format(int(est), C.byref(plan), int(ROUNDING_MODES.get(mode, mode)), int(n_samp))
    if rc < 0:
        raise SNPRelError(lib.snprel_last_error(None).decode())
    return rc


def library_path() -> str:
    return os.path.join(_HERE, "libsnprel_b200.so")


def load_library():
    """Load the CUDA library; fails loudly when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise SNPRelError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(path)
    p, i32, i64, u64, dbl, u32 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double, C.c_uint32
    sig = {
        "snprel_create": [C.POINTER(p), i32],
        "snprel_geno_begin": [p, i64, i64],
        "snprel_geno_push_u8": [p, p, i64],
        "snprel_geno_push_2b": [p, p, i64, i64],
        "snprel_geno_push_bitstream": [p, p, i64, i64],
        "snprel_geno_synth": [p, i64, u64, dbl, dbl, dbl, i64],
        "snprel_geno_dim": [p, C.POINTER(i64), C.POINTER(i64)],
        "snprel_geno_copy_u8": [p, p],
        "snprel_geno_copy_2b": [p, p, i64],
        "snprel_snp_ratefreq": [p, p, p, p],
        "snprel_select_snp_base": [p, i32, dbl, dbl, p, C.POINTER(i64)],
        "snprel_select_snp_base_ex": [p, p, i32, dbl, dbl, p, C.POINTER(i64)],
        "snprel_ibs_num": [p, p, p, p],
        "snprel_ibs_ave": [p, p, i32],
        "snprel_ibd_mom": [p, p, i32, i32, p, p, p],
        "snprel_ibd_mom_sums": [p, p, p, p],
        "snprel_ibd_mom_from_sums": [p, p, i32, i32, p, p],
        "snprel_king_robust": [p, p, p, p, i32],
        "snprel_king_robust_counts": [p, p],
        "snprel_king_homo": [p, p, p, i32],
        "snprel_indiv_beta": [p, i32, p, i32, C.POINTER(dbl)],
        "snprel_indiv_beta_counts": [p, p],
        "snprel_grm": [p, i32, p, i32, C.POINTER(dbl)],
        "snprel_pca": [p, i32, i32, p, C.POINTER(dbl), C.POINTER(dbl), p, p],
        "snprel_eigmix": [p, i32, i32, p, p, p, p],
        "snprel_pca_snp_loading": [p, i32, p, p, dbl, i32, p, p, p],
        "snprel_pca_samp_loading": [p, i32, p, p, p, p],
        "snprel_pca_corr": [p, i32, p, p],
        "snprel_pca_randomized": [p, p, i32, i32, p, p, p],
        "snprel_eigmix_snp_loading": [p, i32, p, p, p, p],
        "snprel_eigmix_samp_loading": [p, i32, p, p, p],
        "snprel_plan_local": [p, i32, C.POINTER(Plan)],
        "snprel_accumulate": [p, i32, C.POINTER(Plan)],
        "snprel_reduce_buffer_count": [p],
        "snprel_reduce_buffer": [p, i32, C.POINTER(p), C.POINTER(i64), C.POINTER(i32)],
        "snprel_mark_reduced": [p],
        "snprel_last_plan": [p, C.POINTER(Plan)],
        "snprel_set_row_window": [p, i64, i64],
        "snprel_window_count": [p, C.POINTER(i64)],
        "snprel_mem_info": [p, C.POINTER(i64), C.POINTER(i64)],
        "snprel_last_hot_kernel": [p, C.POINTER(dbl), C.POINTER(i64), C.POINTER(dbl)],
        "snprel_time_accumulate": [p, i32, i32, C.POINTER(dbl)],
        "snprel_time_finish": [p, i32, C.POINTER(dbl)],
        "snprel_last_step_ms": [p, C.POINTER(dbl)],
        "snprel_invalidate": [p],
        "snprel_table_gram": [p, p, p, p],
        "snprel_debug_flags": [p, u32],
        "snprel_set_count_engine": [p, i32],
        "snprel_set_rounding": [p, i32],
        "snprel_set_snp_origin": [p, i64],
        "snprel_plan_format": [i32, p, i32, i64],
        "snprel_set_async_output": [p, i32],
        "snprel_output_wait": [p],
        "snprel_geno_push_2b_async": [p, p, i64, i64],
        "snprel_geno_wait": [p],
        "snprel_stream_stats": [p, C.POINTER(i64), C.POINTER(i64)],
        "snprel_stream_last_copy_ms": [p, C.POINTER(dbl)],
        "snprel_k1_trace": [p, p, i64, C.POINTER(i64)],
        "snprel_geno_seek": [p, i64],
        "snprel_geno_commit": [p, i64],
        "snprel_geno_device_rows": [p, C.POINTER(p), C.POINTER(i64), C.POINTER(i64)],
        "snprel_multi_geno_begin_replicated": [p, i64, i64],
        "snprel_multi_geno_gather": [p],
        "snprel_multi_grm_tiled": [p, i32, i64, p, p],
        "snprel_reduce_ipc_export": [p, i32, p, C.POINTER(i64)],
        "snprel_peer_reduce_open": [p, i32, i32, p, p],
        "snprel_peer_reduce_phase": [p, i32, i32, C.POINTER(i64)],
        "snprel_multi_create": [p, i32, C.POINTER(p)],
        "snprel_multi_geno_begin": [p, i64, i64],
        "snprel_multi_geno_push_u8": [p, p, i64],
        "snprel_multi_geno_push_2b": [p, p, i64, i64],
        "snprel_multi_geno_synth": [p, i64, u64, dbl, dbl, dbl, i64],
        "snprel_multi_set_row_window": [p, i64, i64],
        "snprel_multi_set_count_engine": [p, i32],
        "snprel_multi_set_rounding": [p, i32],
        "snprel_multi_accumulate": [p, i32, i32, i32],
        "snprel_multi_last_reduce": [p, C.POINTER(dbl), C.POINTER(i64)],
        "snprel_multi_device_count": [p],
        "snprel_last_eigen_info": [p, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(dbl)],
    }
    for name, args in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = i32
    lib.snprel_destroy.argtypes = [p]
    lib.snprel_destroy.restype = None
    lib.snprel_abi_sizeof_plan.argtypes = []
    lib.snprel_abi_sizeof_plan.restype = C.c_int64
    if lib.snprel_abi_sizeof_plan() != C.sizeof(Plan):
        raise SNPRelError(f"{path}: snprel_plan is {lib.snprel_abi_sizeof_plan()} bytes in the library but "
                          f"{C.sizeof(Plan)} in this binding: rebuild the library (or update _lib.Plan)")
    lib.snprel_last_error.argtypes = [p]
    lib.snprel_last_error.restype = C.c_char_p
    lib.snprel_version.argtypes = []
    lib.snprel_version.restype = C.c_char_p
    lib.snprel_kernel_launches.argtypes = [p]
    lib.snprel_kernel_launches.restype = i64
    lib.snprel_device_count.argtypes = []
    lib.snprel_device_count.restype = i32
    lib.snprel_peer_reduce_close.argtypes = [p]
    lib.snprel_peer_reduce_close.restype = None
    lib.snprel_multi_destroy.argtypes = [p]
    lib.snprel_multi_destroy.restype = None
    lib.snprel_multi_last_error.argtypes = [p]
    lib.snprel_multi_last_error.restype = C.c_char_p
    lib.snprel_multi_ctx.argtypes = [p, i32]
    lib.snprel_multi_ctx.restype = p
    _LIB = lib
    return lib


EXPORTED_SYMBOLS = [
    "snprel_create", "snprel_destroy", "snprel_last_error", "snprel_version", "snprel_abi_sizeof_plan",
    "snprel_geno_begin", "snprel_geno_push_u8", "snprel_geno_push_2b", "snprel_geno_push_bitstream", "snprel_geno_synth",
    "snprel_geno_dim", "snprel_geno_copy_u8", "snprel_geno_copy_2b", "snprel_snp_ratefreq", "snprel_select_snp_base", "snprel_select_snp_base_ex",
    "snprel_ibs_num", "snprel_ibs_ave", "snprel_ibd_mom", "snprel_ibd_mom_sums", "snprel_ibd_mom_from_sums", "snprel_king_robust", "snprel_king_robust_counts",
    "snprel_king_homo", "snprel_indiv_beta", "snprel_indiv_beta_counts", "snprel_grm",
    "snprel_pca", "snprel_eigmix", "snprel_pca_snp_loading", "snprel_pca_samp_loading", "snprel_pca_corr", "snprel_pca_randomized",
    "snprel_eigmix_snp_loading", "snprel_eigmix_samp_loading", "snprel_plan_local", "snprel_accumulate",
    "snprel_reduce_buffer_count", "snprel_reduce_buffer", "snprel_mark_reduced", "snprel_last_plan", "snprel_set_row_window", "snprel_window_count", "snprel_mem_info",
    "snprel_kernel_launches", "snprel_last_hot_kernel", "snprel_time_accumulate", "snprel_time_finish", "snprel_last_step_ms", "snprel_invalidate",
    "snprel_table_gram", "snprel_debug_flags", "snprel_set_count_engine", "snprel_last_eigen_info", "snprel_set_rounding",
    "snprel_set_snp_origin", "snprel_plan_format", "snprel_multi_set_rounding",
    "snprel_device_count", "snprel_multi_create", "snprel_multi_destroy", "snprel_multi_last_error", "snprel_multi_device_count", "snprel_multi_ctx",
    "snprel_multi_geno_begin", "snprel_multi_geno_push_u8", "snprel_multi_geno_push_2b", "snprel_multi_geno_synth",
    "snprel_multi_set_row_window", "snprel_multi_set_count_engine", "snprel_multi_accumulate", "snprel_multi_last_reduce",
    "snprel_geno_seek", "snprel_geno_device_rows", "snprel_geno_commit",
    "snprel_multi_geno_begin_replicated", "snprel_multi_geno_gather", "snprel_multi_grm_tiled",
    "snprel_set_async_output", "snprel_output_wait", "snprel_geno_push_2b_async", "snprel_geno_wait", "snprel_stream_stats", "snprel_stream_last_copy_ms", "snprel_k1_trace",
    "snprel_reduce_ipc_export", "snprel_peer_reduce_open", "snprel_peer_reduce_phase", "snprel_peer_reduce_close",
]


def window_owner(k: int, world: int) -> int:
    """Rank that computes row window k when the N x N output is tiled across `world` GPUs.
    Windows shrink with k (upper triangle), so they are dealt in boustrophedon order
    (0..w-1, w-1..0, ...): every rank's share of the pairs is within a fraction of a window of
    1/world, where plain round-robin always hands rank 0 the largest window of each round."""
    r = k % (2 * world)
    return r if r < world else 2 * world - 1 - r


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One CUDA device + one genotype workspace (the reference's process-global
    MCWorkingGeno, src/dGenGWAS.cpp:2000)."""

    def __init__(self, device: int = 0, _borrowed=None):
        self.lib = load_library()
        self.device = device
        self._owned = _borrowed is None
        if _borrowed is not None:      # a context that belongs to a MultiContext
            self.h = C.c_void_p(_borrowed)
            return
        h = C.c_void_p()
        rc = self.lib.snprel_create(C.byref(h), int(device))
        if rc != 0:
            raise SNPRelError(self.lib.snprel_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            if self._owned:
                self.lib.snprel_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise SNPRelError(self.lib.snprel_last_error(self.h).decode())

    # ---- workspace ----
    def geno_begin(self, n_samp, snp_capacity):
        self._ck(self.lib.snprel_geno_begin(self.h, int(n_samp), int(snp_capacity)))

    def geno_push_u8(self, block):
        block = np.ascontiguousarray(block, dtype=np.uint8)
        if block.ndim != 2:
            raise SNPRelError("geno_push_u8: block must be [cnt, n_samp]")
        self._ck(self.lib.snprel_geno_push_u8(self.h, _ptr(block), block.shape[0]))

    def geno_push_2b(self, packed):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        self._ck(self.lib.snprel_geno_push_2b(self.h, _ptr(packed), packed.shape[0], packed.shape[1]))

    def geno_push_bitstream(self, stream, first_snp, cnt):
        """GDS dBit2 payload (continuous 2-bit stream, no row padding) -> `cnt` SNPs from `first_snp`."""
        stream = np.ascontiguousarray(stream, dtype=np.uint8).reshape(-1)
        n, _ = self.geno_dim()
        if (int(first_snp) + int(cnt)) * n * 2 > stream.size * 8:
            raise SNPRelError("geno_push_bitstream: the stream is shorter than the requested SNP range")
        self._ck(self.lib.snprel_geno_push_bitstream(self.h, _ptr(stream), int(first_snp) * n, int(cnt)))

    def set_async_output(self, on=True):
        """Row-window pipelines: ibs_ave / king_robust / grm return once their device-to-host copy is queued;
        call output_wait() before reading (or reusing) the host buffer."""
        self._ck(self.lib.snprel_set_async_output(self.h, int(bool(on))))

    def output_wait(self):
        self._ck(self.lib.snprel_output_wait(self.h))

    def geno_push_2b_async(self, packed):
        """Queue the host-to-device copy and return; `packed` (ideally pinned) must stay alive until the next
        call on this context returns.  A following pca() / grm() / eigmix() overlaps with the copy."""
        if packed.dtype != np.uint8 or not packed.flags.c_contiguous or packed.ndim != 2:
            raise SNPRelError("geno_push_2b_async: a C-contiguous uint8 [cnt, row_bytes] array is required (no copy is made)")
        self._async_keepalive = packed
        self._ck(self.lib.snprel_geno_push_2b_async(self.h, _ptr(packed), packed.shape[0], packed.shape[1]))

    def geno_wait(self):
        self._ck(self.lib.snprel_geno_wait(self.h))
        self._async_keepalive = None

    def stream_stats(self):
        a, b = C.c_int64(), C.c_int64()
        self._ck(self.lib.snprel_stream_stats(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def k1_trace(self):
        """[items, 8] clock64 stamps of the last table-Gram launch (after debug_flags(1))."""
        n = C.c_int64()
        self._ck(self.lib.snprel_k1_trace(self.h, None, 0, C.byref(n)))
        out = np.zeros((max(n.value, 1), 8), dtype=np.int64)
        self._ck(self.lib.snprel_k1_trace(self.h, _ptr(out), n.value, C.byref(n)))
        return out[: n.value]

    def stream_last_copy_ms(self):
        ms = C.c_double()
        self._ck(self.lib.snprel_stream_last_copy_ms(self.h, C.byref(ms)))
        return ms.value

    def geno_seek(self, snp_index):
        self._ck(self.lib.snprel_geno_seek(self.h, int(snp_index)))

    def geno_commit(self, n_snp):
        self._ck(self.lib.snprel_geno_commit(self.h, int(n_snp)))

    def geno_device_rows(self):
        """(device pointer of SNP row 0, row pitch in bytes, reserved rows) of the 2-bit workspace."""
        ptr, rb, cap = C.c_void_p(), C.c_int64(), C.c_int64()
        self._ck(self.lib.snprel_geno_device_rows(self.h, C.byref(ptr), C.byref(rb), C.byref(cap)))
        return ptr.value, rb.value, cap.value

    def geno_synth(self, n_snp, seed=20261017, maf_lo=0.05, maf_hi=0.5, miss_rate=0.005, snp_start=0):
        self._ck(self.lib.snprel_geno_synth(self.h, int(n_snp), int(seed), float(maf_lo), float(maf_hi),
                                            float(miss_rate), int(snp_start)))

    def geno_dim(self):
        a, b = C.c_int64(), C.c_int64()
        self._ck(self.lib.snprel_geno_dim(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def geno_copy_u8(self):
        n, m = self.geno_dim()
        out = np.empty((m, n), dtype=np.uint8)
        self._ck(self.lib.snprel_geno_copy_u8(self.h, _ptr(out)))
        return out

    def geno_copy_2b(self, out=None):
        n, m = self.geno_dim()
        rb = (n + 3) // 4
        if out is None:
            out = np.empty((m, rb), dtype=np.uint8)
        self._ck(self.lib.snprel_geno_copy_2b(self.h, _ptr(out), out.shape[1]))
        return out

    def snp_ratefreq(self):
        _, m = self.geno_dim()
        af, maf, mr = (np.empty(m) for _ in range(3))
        self._ck(self.lib.snprel_snp_ratefreq(self.h, _ptr(af), _ptr(maf), _ptr(mr)))
        return af, maf, mr

    def select_snp_base(self, remove_mono=True, maf=-1.0, missrate=2.0, allele_freq=None):
        _, m = self.geno_dim()
        sel = np.empty(m, dtype=np.uint8)
        nrm = C.c_int64()
        if allele_freq is None:
            self._ck(self.lib.snprel_select_snp_base(self.h, int(bool(remove_mono)), float(maf), float(missrate),
                                                     _ptr(sel), C.byref(nrm)))
        else:
            af = self._afreq_in(allele_freq)
            self._ck(self.lib.snprel_select_snp_base_ex(self.h, _ptr(af), int(bool(remove_mono)), float(maf),
                                                        float(missrate), _ptr(sel), C.byref(nrm)))
        return sel.astype(bool), nrm.value

    # ---- estimators ----
    # ---- row windows ----
    def set_row_window(self, row0=0, rows=0):
        self._ck(self.lib.snprel_set_row_window(self.h, int(row0), int(rows)))
        self._win = rows > 0

    def window_count(self):
        cnt = C.c_int64()
        self._ck(self.lib.snprel_window_count(self.h, C.byref(cnt)))
        return cnt.value

    def mem_info(self):
        f, t = C.c_int64(), C.c_int64()
        self._ck(self.lib.snprel_mem_info(self.h, C.byref(f), C.byref(t)))
        return f.value, t.value

    def auto_window_rows(self, bytes_per_pair):
        """0 when whole-matrix accumulators (bytes_per_pair per pair, padded square) fit
        comfortably in free HBM, else a window height (multiple of 256) that uses about a
        quarter of it."""
        n, _ = self.geno_dim()
        npad = (n + 255) // 256 * 256
        free, _ = self.mem_info()
        if bytes_per_pair * npad * npad <= 0.5 * free:
            return 0
        rows = int(0.25 * free / (bytes_per_pair * npad)) // 256 * 256
        return max(256, rows)

    def packed_by_windows(self, fn, rows, rank=0, world=1):
        """Run fn() (an estimator call with packed=True) over row windows of height `rows`
        and concatenate the packed slices.  With world > 1 this rank only computes the windows
        window_owner() deals to it and returns a list of (packed_offset, slice[s]) -- the N x N
        output tiled across GPUs that all hold the same genotypes, no collective."""
        n, _ = self.geno_dim()
        parts, offs, off = [], [], 0
        try:
            for k, (r0, h) in enumerate(self.windows(rows)):
                self.set_row_window(r0, h)
                cnt = self.window_count()
                if window_owner(k, world) == rank:
                    parts.append(fn())
                    offs.append(off)
                off += cnt
        finally:
            self.set_row_window(0, 0)
        if world > 1:
            return list(zip(offs, parts))
        if isinstance(parts[0], (tuple, list)):
            return tuple(np.concatenate([p[i] for p in parts]) for i in range(len(parts[0])))
        return np.concatenate(parts)

    def windows(self, rows):
        """Iterate (row0, rows) windows of height `rows` (multiple of 256) over all samples."""
        n, _ = self.geno_dim()
        for r0 in range(0, n, rows):
            yield r0, rows

    def _out(self, packed, out=None):
        """Result buffer: a fresh array, or a view of the caller's `out` (C-contiguous float64,
        e.g. pinned host memory re-used across row windows)."""
        win = getattr(self, "_win", False)
        if win and not packed:
            raise SNPRelError("a row window returns the packed upper triangle only (useMatrix)")
        n, _ = self.geno_dim()
        need = self.window_count() if win else (n * (n + 1) // 2 if packed else n * n)
        if out is None:
            return np.empty(need) if (win or packed) else np.empty((n, n))
        if out.dtype != np.float64 or not out.flags.c_contiguous or out.size < need:
            raise SNPRelError("'out' must be a C-contiguous float64 buffer of at least the result's size")
        o = out.reshape(-1)[:need]
        return o if (win or packed) else o.reshape(n, n)

    def ibs_num(self):
        n, _ = self.geno_dim()
        if getattr(self, "_win", False):
            o = [np.empty(self.window_count(), dtype=np.int32) for _ in range(3)]
        else:
            o = [np.empty((n, n), dtype=np.int32) for _ in range(3)]
        self._ck(self.lib.snprel_ibs_num(self.h, _ptr(o[0]), _ptr(o[1]), _ptr(o[2])))
        return o

    def ibs_ave(self, packed=False, out=None):
        o = self._out(packed, out)
        self._ck(self.lib.snprel_ibs_ave(self.h, _ptr(o), int(packed)))
        return o

    def _afreq_in(self, allele_freq):
        if allele_freq is None:
            return None
        af = np.ascontiguousarray(allele_freq, dtype=np.float64)
        if af.shape != (self.geno_dim()[1],):
            raise SNPRelError("allele_freq must have one entry per SNP")
        return af

    def ibd_mom(self, allele_freq=None, kinship_constraint=False, packed=False):
        """PLINK method of moments -> (k0, k1, afreq)."""
        k0, k1 = self._out(packed), self._out(packed)
        af = np.empty(self.geno_dim()[1], dtype=np.float64)
        afin = self._afreq_in(allele_freq)
        self._ck(self.lib.snprel_ibd_mom(self.h, _ptr(afin), int(bool(kinship_constraint)), int(packed),
                                         _ptr(k0), _ptr(k1), _ptr(af)))
        return k0, k1, af

    def ibd_mom_sums(self, allele_freq=None):
        sums = np.empty(6, dtype=np.float64)
        af = np.empty(self.geno_dim()[1], dtype=np.float64)
        self._ck(self.lib.snprel_ibd_mom_sums(self.h, _ptr(self._afreq_in(allele_freq)), _ptr(sums), _ptr(af)))
        return sums, af

    def ibd_mom_from_sums(self, sums, kinship_constraint=False, packed=False):
        k0, k1 = self._out(packed), self._out(packed)
        sums = np.ascontiguousarray(sums, dtype=np.float64)
        self._ck(self.lib.snprel_ibd_mom_from_sums(self.h, _ptr(sums), int(bool(kinship_constraint)),
                                                   int(packed), _ptr(k0), _ptr(k1)))
        return k0, k1

    def king_robust(self, family_id=None, packed=False, out=None):
        """`out`: optional pair of caller buffers (IBS0, kinship)."""
        a, b = self._out(packed, None if out is None else out[0]), self._out(packed, None if out is None else out[1])
        fam = None if family_id is None else np.ascontiguousarray(family_id, dtype=np.int32)
        self._ck(self.lib.snprel_king_robust(self.h, _ptr(fam), _ptr(a), _ptr(b), int(packed)))
        return a, b

    def king_robust_counts(self):
        n, _ = self.geno_dim()
        o = np.empty((5, n, n), dtype=np.int32)
        self._ck(self.lib.snprel_king_robust_counts(self.h, _ptr(o)))
        return o

    def king_homo(self, packed=False):
        a, b = self._out(packed), self._out(packed)
        self._ck(self.lib.snprel_king_homo(self.h, _ptr(a), _ptr(b), int(packed)))
        return a, b

    def indiv_beta(self, inbreeding=True, packed=False):
        o = self._out(packed)
        avg = C.c_double()
        self._ck(self.lib.snprel_indiv_beta(self.h, int(inbreeding), _ptr(o), int(packed), C.byref(avg)))
        return o, avg.value

    def indiv_beta_counts(self):
        n, _ = self.geno_dim()
        o = np.empty((2, n, n), dtype=np.int32)
        self._ck(self.lib.snprel_indiv_beta_counts(self.h, _ptr(o)))
        return o

    def grm(self, method="GCTA", packed=False, out=None):
        """`out`: caller buffer (e.g. pinned host memory) of at least the result's size."""
        if method not in GRM_METHODS:
            raise SNPRelError("Invalid 'method'!")
        if method == "Corr":
            packed = False
        o = self._out(packed, out)
        avg = C.c_double()
        self._ck(self.lib.snprel_grm(self.h, GRM_METHODS[method], _ptr(o), int(packed), C.byref(avg)))
        return o, avg.value

    def pca(self, eigen_cnt=32, bayesian=False, need_genmat=False, genmat_only=False, genmat_out=None):
        n, _ = self.geno_dim()
        k = min(int(eigen_cnt), n)
        genmat = genmat_out
        if genmat is None and (need_genmat or genmat_only):
            genmat = np.empty((n, n))
        tx, tv = C.c_double(), C.c_double()
        eigval = eigvec = None
        if not genmat_only:
            eigval = np.empty(n)
            eigvec = np.empty((k, n))      # column-major n x k == C-order k x n
        self._ck(self.lib.snprel_pca(self.h, k, int(bool(bayesian)), _ptr(genmat), C.byref(tx), C.byref(tv),
                                     _ptr(eigval), _ptr(eigvec)))
        return dict(TraceXTX=tx.value, TraceVal=tv.value, genmat=genmat, eigenval=eigval,
                    eigenvect=None if eigvec is None else eigvec.T)

    def eigmix(self, eigen_cnt=32, diagadj=True, ibdmat=False):
        n, m = self.geno_dim()
        k = n if (eigen_cnt < 0 or eigen_cnt > n) else int(eigen_cnt)
        ibd = np.empty((n, n)) if ibdmat else None
        af = np.empty(m)
        eigval = np.empty(n) if k > 0 else None
        eigvec = np.empty((k, n)) if k > 0 else None
        self._ck(self.lib.snprel_eigmix(self.h, k, int(bool(diagadj)), _ptr(ibd), _ptr(af), _ptr(eigval),
                                        _ptr(eigvec)))
        return dict(eigenval=eigval, eigenvect=None if eigvec is None else eigvec.T, afreq=af, ibd=ibd)

    # ---- loadings / projection / correlation (matrices in the reference's R layouts) ----
    @staticmethod
    def _colmajor(a, name):
        a = np.asarray(a, dtype=np.float64)
        if a.ndim != 2:
            raise SNPRelError(f"{name} must be a matrix")
        return np.asfortranarray(a)

    def pca_snp_loading(self, eigenval, eigenvect, trace_xtx, bayesian=False):
        """-> (snploading [k, n_snp], avgfreq [n_snp], scale [n_snp])."""
        n, m = self.geno_dim()
        v = self._colmajor(eigenvect, "eigenvect")
        if v.shape[0] != n:
            raise SNPRelError("the number of samples should be equal to the number of rows in 'eigenvect'.")
        k = v.shape[1]
        ev = np.ascontiguousarray(np.asarray(eigenval, dtype=np.float64)[:k])
        load, af, sc = np.empty((m, k)), np.empty(m), np.empty(m)
        self._ck(self.lib.snprel_pca_snp_loading(self.h, k, _ptr(ev), _ptr(v), float(trace_xtx), int(bool(bayesian)),
                                                 _ptr(load), _ptr(af), _ptr(sc)))
        return load.T, af, sc

    def pca_samp_loading(self, loadings, avgfreq, scale):
        """loadings [k, n_snp] (pre-scaled) -> eigenvect [n_samp, k]."""
        n, m = self.geno_dim()
        ld = self._colmajor(loadings, "loadings")          # k x n_snp column-major == [n_snp][k]
        if ld.shape[1] != m:
            raise SNPRelError("the number of SNPs should be equal to the number of columns in 'snploading'.")
        k = ld.shape[0]
        af = np.ascontiguousarray(avgfreq, dtype=np.float64)
        sc = np.ascontiguousarray(scale, dtype=np.float64)
        if af.shape != (m,) or sc.shape != (m,):
            raise SNPRelError("'avgfreq' and 'scale' must have one entry per SNP")
        out = np.empty((k, n))
        self._ck(self.lib.snprel_pca_samp_loading(self.h, k, _ptr(ld), _ptr(af), _ptr(sc), _ptr(out)))
        return out.T

    def pca_randomized(self, aux_mat, aux_dim, iter_num=10):
        """gnrPCA "randomized" -> (sigma [n], V^T [aux_dim * (iter_num + 1), n], 2 TraceXTX)."""
        n, _ = self.geno_dim()
        aux = np.ascontiguousarray(aux_mat, dtype=np.float64).reshape(-1)
        if aux.size != aux_dim * n:
            raise SNPRelError("aux_mat must hold aux_dim * n_samp values")
        hsize = aux_dim * (iter_num + 1)
        sigma, vt = np.empty(n), np.empty((hsize, n))
        tr = C.c_double()
        self._ck(self.lib.snprel_pca_randomized(self.h, _ptr(aux), int(aux_dim), int(iter_num), _ptr(sigma), _ptr(vt),
                                                C.byref(tr)))
        return sigma, vt, tr.value

    def pca_corr(self, eigenvect):
        """-> snpcorr [k, n_snp]."""
        n, m = self.geno_dim()
        v = self._colmajor(eigenvect, "eigenvect")
        if v.shape[0] != n:
            raise SNPRelError("the number of samples should be equal to the number of rows in 'eigenvect'.")
        k = v.shape[1]
        out = np.empty((m, k))
        self._ck(self.lib.snprel_pca_corr(self.h, k, _ptr(v), _ptr(out)))
        return out.T

    def eigmix_snp_loading(self, eigenval, eigenvect, afreq):
        n, m = self.geno_dim()
        v = self._colmajor(eigenvect, "eigenvect")
        if v.shape[0] != n:
            raise SNPRelError("the number of samples should be equal to the number of rows in 'eigenvect'.")
        k = v.shape[1]
        ev = np.ascontiguousarray(np.asarray(eigenval, dtype=np.float64)[:k])
        load = np.empty((m, k))
        self._ck(self.lib.snprel_eigmix_snp_loading(self.h, k, _ptr(ev), _ptr(v), _ptr(self._afreq_in(afreq)), _ptr(load)))
        return load.T

    def eigmix_samp_loading(self, loadings, afreq):
        n, m = self.geno_dim()
        ld = self._colmajor(loadings, "loadings")
        if ld.shape[1] != m:
            raise SNPRelError("the number of SNPs should be equal to the number of columns in 'snploading'.")
        k = ld.shape[0]
        out = np.empty((k, n))
        self._ck(self.lib.snprel_eigmix_samp_loading(self.h, k, _ptr(ld), _ptr(self._afreq_in(afreq)), _ptr(out)))
        return out.T

    # ---- split accumulate / reduce / finish ----
    def plan_local(self, est, bayesian=False, tol=0.0):
        pl = Plan()
        pl.frac_bits = -1
        pl.frac_bits_w = -1
        pl.frac_bits_d = -1
        pl.tol = float(tol)
        pl.bayesian = int(bool(bayesian))
        self._ck(self.lib.snprel_plan_local(self.h, int(est), C.byref(pl)))
        return pl

    def accumulate(self, est, plan=None):
        self._ck(self.lib.snprel_accumulate(self.h, int(est), None if plan is None else C.byref(plan)))

    def reduce_buffers(self):
        out = []
        for i in range(self.lib.snprel_reduce_buffer_count(self.h)):
            ptr, cnt, kind = C.c_void_p(), C.c_int64(), C.c_int()
            self._ck(self.lib.snprel_reduce_buffer(self.h, i, C.byref(ptr), C.byref(cnt), C.byref(kind)))
            out.append((ptr.value, cnt.value, kind.value))
        return out

    def reduce_ipc_handles(self):
        """(handles bytes [n_buffers * 64], offsets int64 [n_buffers]) of this context's reduce buffers."""
        nb = self.lib.snprel_reduce_buffer_count(self.h)
        hb = np.zeros(nb * 64, dtype=np.uint8)
        off = np.zeros(nb, dtype=np.int64)
        for k in range(nb):
            o = C.c_int64()
            self._ck(self.lib.snprel_reduce_ipc_export(self.h, k, hb[k * 64:].ctypes.data_as(C.c_void_p), C.byref(o)))
            off[k] = o.value
        return hb, off

    def peer_reduce_open(self, rank, world, handles, offsets):
        handles = np.ascontiguousarray(handles, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self._ck(self.lib.snprel_peer_reduce_open(self.h, int(rank), int(world), _ptr(handles), _ptr(offsets)))

    def peer_reduce_phase(self, phase, root=-1):
        b = C.c_int64()
        self._ck(self.lib.snprel_peer_reduce_phase(self.h, int(phase), int(root), C.byref(b)))
        return b.value

    def last_plan(self):
        pl = Plan()
        self._ck(self.lib.snprel_last_plan(self.h, C.byref(pl)))
        return pl

    def mark_reduced(self):
        self._ck(self.lib.snprel_mark_reduced(self.h))

    # ---- introspection ----
    def kernel_launches(self):
        return int(self.lib.snprel_kernel_launches(self.h))

    def last_hot_kernel(self):
        ms, n, u = C.c_double(), C.c_int64(), C.c_double()
        self._ck(self.lib.snprel_last_hot_kernel(self.h, C.byref(ms), C.byref(n), C.byref(u)))
        return ms.value, n.value, u.value

    def time_accumulate(self, est, reps=1):
        ms = C.c_double()
        self._ck(self.lib.snprel_time_accumulate(self.h, int(est), int(reps), C.byref(ms)))
        return ms.value

    def time_finish(self, est):
        """Device epilogue of the accumulators at hand (result stays on the device); ms."""
        ms = C.c_double()
        self._ck(self.lib.snprel_time_finish(self.h, int(est), C.byref(ms)))
        return ms.value

    def last_step_ms(self):
        ms = C.c_double()
        self._ck(self.lib.snprel_last_step_ms(self.h, C.byref(ms)))
        return ms.value

    def invalidate(self):
        self._ck(self.lib.snprel_invalidate(self.h))

    def table_gram(self, tabA, tabB):
        n, m = self.geno_dim()
        tabA = np.ascontiguousarray(tabA, dtype=np.int8)
        tabB = np.ascontiguousarray(tabB, dtype=np.int8)
        if tabA.shape != (m, 4) or tabB.shape != (4,):
            raise SNPRelError("table_gram: tabA must be [n_snp, 4] and tabB [4] int8")
        out = np.empty((n, n), dtype=np.int64)
        self._ck(self.lib.snprel_table_gram(self.h, _ptr(tabA), _ptr(tabB), _ptr(out)))
        return out

    def last_eigen_info(self):
        """(solver, filter rounds, block products) of the last eigen step; solver 1 = Chebyshev-filtered
        subspace iteration, 0 = dense cusolverDnXsyevd."""
        a, b, g = C.c_int(), C.c_int(), C.c_int()
        ph = (C.c_double * 3)()
        self._ck(self.lib.snprel_last_eigen_info(self.h, C.byref(a), C.byref(b), C.byref(g), ph))
        self.eigen_phase_ms = {"filter": ph[0], "orthonormalise": ph[1], "rayleigh_ritz": ph[2]}
        return a.value, b.value, g.value

    def set_rounding(self, mode):
        """Rounding of the main row table of the covariance estimators: 'nearest' (worst-case error
        bound), 'random' (unbiased randomised rounding + Hoeffding bound, failure probability 1e-12)
        or 'auto' (default: whichever needs fewer tensor passes for the tolerance)."""
        code = ROUNDING_MODES.get(mode, mode)
        self._ck(self.lib.snprel_set_rounding(self.h, int(code)))

    def set_snp_origin(self, origin):
        """Global index of this context's first SNP row (SNP shards: the shard offset); keys the
        rounding draws so that a sharded run reproduces the one-context result bit for bit."""
        self._ck(self.lib.snprel_set_snp_origin(self.h, int(origin)))

    def set_count_engine(self, engine):
        """'bits' (default: packed-bit XOR/AND/popcount kernels) or 'tensor' (same exact counters
        from the tensor pipe) for IBS / KING-robust / IndivBeta / PLINK-MoM."""
        code = {"bits": 0, "tensor": 1}.get(engine, engine)
        self._ck(self.lib.snprel_set_count_engine(self.h, int(code)))

    def debug_flags(self, flags):
        self._ck(self.lib.snprel_debug_flags(self.h, int(flags)))


class MultiContext:
    """Several GPUs of one box behind one handle, one host process (snprel_multi_*, csrc/multi.cu):
    SNP-block sharding, per-device accumulation on library threads, peer-memory reduction over
    NVLink.  After accumulate() the finishing calls are made on `ctx(root)`.  `devices` may repeat
    a GPU (the whole path then runs on one device)."""

    def __init__(self, devices):
        self.lib = load_library()
        self.devices = [int(d) for d in devices]
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        if self.lib.snprel_multi_create(arr, len(self.devices), C.byref(h)) != 0:
            raise SNPRelError(self.lib.snprel_multi_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.snprel_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise SNPRelError(self.lib.snprel_multi_last_error(self.h).decode())

    def ctx(self, i=0) -> Context:
        p = self.lib.snprel_multi_ctx(self.h, int(i))
        if not p:
            raise SNPRelError("MultiContext.ctx: device index out of range")
        c = Context(self.devices[i], _borrowed=p)
        c._win = getattr(self, "_win", False)
        return c

    def geno_begin(self, n_samp, snp_capacity):
        self._ck(self.lib.snprel_multi_geno_begin(self.h, int(n_samp), int(snp_capacity)))

    def geno_push_u8(self, block):
        block = np.ascontiguousarray(block, dtype=np.uint8)
        self._ck(self.lib.snprel_multi_geno_push_u8(self.h, _ptr(block), block.shape[0]))

    def geno_push_2b(self, packed):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        self._ck(self.lib.snprel_multi_geno_push_2b(self.h, _ptr(packed), packed.shape[0], packed.shape[1]))

    def geno_synth(self, n_snp, seed=20261017, maf_lo=0.05, maf_hi=0.5, miss_rate=0.005, snp_start=0):
        self._ck(self.lib.snprel_multi_geno_synth(self.h, int(n_snp), int(seed), float(maf_lo), float(maf_hi),
                                                  float(miss_rate), int(snp_start)))

    def set_row_window(self, row0=0, rows=0):
        self._ck(self.lib.snprel_multi_set_row_window(self.h, int(row0), int(rows)))
        self._win = rows > 0

    def set_count_engine(self, engine):
        code = {"bits": 0, "tensor": 1}.get(engine, engine)
        self._ck(self.lib.snprel_multi_set_count_engine(self.h, int(code)))

    def set_rounding(self, mode):
        self._ck(self.lib.snprel_multi_set_rounding(self.h, int(ROUNDING_MODES.get(mode, mode))))

    def accumulate(self, est, bayesian=False, root=0):
        """est: a GRM method name, or EST_IBS / EST_KING_ROBUST / EST_BETA.  root = -1: all-reduce."""
        if isinstance(est, str):
            est = GRM_METHODS[est]
        self._ck(self.lib.snprel_multi_accumulate(self.h, int(est), int(bool(bayesian)), int(root)))

    def geno_begin_replicated(self, n_samp, snp_capacity):
        self._ck(self.lib.snprel_multi_geno_begin_replicated(self.h, int(n_samp), int(snp_capacity)))

    def geno_gather(self):
        self._ck(self.lib.snprel_multi_geno_gather(self.h))

    def grm_tiled(self, method, window_rows=0, out=None):
        """Packed upper triangle (CdMatTri order) of the GRM computed window by window, window w on device
        w mod n.  `out`: float64 buffer of n (n + 1) / 2 values (allocated when None)."""
        n, _ = self.ctx(0).geno_dim()
        need = n * (n + 1) // 2
        if out is None:
            out = np.empty(need)
        flat = out.reshape(-1)
        calls = []
        SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.c_int64)

        def sink(_user, first, vals, count):
            flat[first: first + count] = np.ctypeslib.as_array(vals, shape=(count,))
            calls.append((first, count))
            return 0
        cb = SINK(sink)
        self._ck(self.lib.snprel_multi_grm_tiled(self.h, GRM_METHODS[method], int(window_rows), C.cast(cb, C.c_void_p), None))
        self.tiled_calls = calls
        return out

    def last_reduce(self):
        ms, b = C.c_double(), C.c_int64()
        self._ck(self.lib.snprel_multi_last_reduce(self.h, C.byref(ms), C.byref(b)))
        return ms.value, b.value
