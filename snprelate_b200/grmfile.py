"""On-disk GRM output and merging.

The reference writes GRMs to a gdsfmt container tagged FileFormat=SNPRELATE_OUTPUT
with nodes `command`, `sample.id`, `snp.id`, `grm` (n x n, float64 or float32,
appended row by row: grm_save_to_gds, src/genPCA.cpp:1571-1584) and `avg_val`
(R/IBD.R:570-612), and merges such files with snpgdsMergeGRM -> gnrGRMMerge
(R/IBD.R:624-748, src/genPCA.cpp:1721-1857).  gdsfmt is a third-party storage
engine that is not part of this build, so the same logical nodes live in a flat
binary file that can be appended and memory-mapped row by row:

    bytes 0..7    magic  b"SNPRELO1"
    bytes 8..15   uint64 length H of the JSON header
    bytes 16..23  float64 avg_val (NaN when absent; patched after the rows)
    bytes 24..    JSON header (FileFormat, version, command, n_samp, n_snp, prec,
                  dtype / shape of sample.id and snp.id), then the two id arrays
                  (.npy payloads, 64-byte aligned), then the n x n matrix, row-major
                  (= column-major, it is symmetric), 64-byte aligned.

Rows are streamed: a writer receives row bands of the packed upper triangle as the
device finishes them (Context.packed_by_windows) and scatters each band into the
memory-mapped n x n region, so N^2 never has to sit in host memory twice.
"""
from __future__ import annotations

import io
import json
import struct

import numpy as np

MAGIC = b"SNPRELO1"
FILE_FORMAT = "SNPRELATE_OUTPUT"
_PREC = {"double": np.float64, "float64": np.float64, "single": np.float32, "float32": np.float32}


class GrmFileError(RuntimeError):
    pass


def _align(x, a=64):
    return (x + a - 1) // a * a


def _npy_bytes(a):
    b = io.BytesIO()
    np.save(b, np.asarray(a), allow_pickle=False)
    return b.getvalue()


class GrmWriter:
    """createfn.gds + add.gdsn(...) of R/IBD.R:570-590."""

    def __init__(self, path, command, sample_id, snp_id, out_prec="double", version="snprelate_b200"):
        if out_prec not in _PREC:
            raise GrmFileError("'out.prec' should be one of \"double\", \"single\"")
        self.path = path
        self.dtype = np.dtype(_PREC[out_prec])
        self.n = int(len(sample_id))
        ids = [_npy_bytes(np.asarray(sample_id)), _npy_bytes(np.asarray(snp_id))]
        hdr = {"FileFormat": FILE_FORMAT, "version": version, "command": list(command),
               "n_samp": self.n, "n_snp": int(len(snp_id)), "prec": self.dtype.name,
               "sample_id": None, "snp_id": None, "grm": None}
        # two passes: offsets depend on the header length, which depends on the offsets' digits
        for _ in range(3):
            h = json.dumps(hdr).encode()
            off = _align(24 + len(h) + 64)
            hdr["sample_id"] = [off, len(ids[0])]
            off = _align(off + len(ids[0]))
            hdr["snp_id"] = [off, len(ids[1])]
            off = _align(off + len(ids[1]))
            hdr["grm"] = off
        h = json.dumps(hdr).encode()
        assert 24 + len(h) <= hdr["sample_id"][0]
        self.grm_off = hdr["grm"]
        with open(path, "wb") as f:
            f.write(MAGIC + struct.pack("<Q", len(h)) + struct.pack("<d", float("nan")) + h)
            f.seek(hdr["sample_id"][0])
            f.write(ids[0])
            f.seek(hdr["snp_id"][0])
            f.write(ids[1])
            f.truncate(self.grm_off + self.n * self.n * self.dtype.itemsize)
        self.mm = np.memmap(path, dtype=self.dtype, mode="r+", offset=self.grm_off, shape=(self.n, self.n))

    def write_full(self, mat):
        self.mm[:, :] = mat

    def write_band(self, row0, packed_slice):
        """Rows [row0, row0 + h) of the packed upper triangle (CdMatTri order,
        src/dGenGWAS.h:556-561), h implied by the slice length: fill the band and its
        transpose (CdMatTri::GetRow semantics, src/dGenGWAS.h:563-572)."""
        n = self.n
        cnt, h = 0, 0
        while cnt < len(packed_slice):
            cnt += n - (row0 + h)
            h += 1
        if cnt != len(packed_slice):
            raise GrmFileError("band slice does not end on a row boundary")
        band = np.zeros((h, n - row0), dtype=np.float64)
        iu = np.triu_indices(h, 0, n - row0)
        band[iu] = packed_slice
        sq = band[:, :h]
        band[:, :h] = sq + np.triu(sq, 1).T          # square part of the band: symmetrise
        self.mm[row0:row0 + h, row0:] = band
        if row0 + h < n:
            self.mm[row0 + h:, row0:row0 + h] = band[:, h:].T

    def close(self, avg_val=None):
        self.mm.flush()
        del self.mm
        if avg_val is not None:
            with open(self.path, "r+b") as f:
                f.seek(16)
                f.write(struct.pack("<d", float(avg_val)))


class GrmFile:
    """openfn.gds + the node reads snpgdsMergeGRM performs (R/IBD.R:646-676)."""

    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            head = f.read(24)
            if len(head) < 24 or head[:8] != MAGIC:
                raise GrmFileError(f"'{path}' is not valid.")
            hlen, = struct.unpack("<Q", head[8:16])
            self.avg_val, = struct.unpack("<d", head[16:24])
            hdr = json.loads(f.read(hlen).decode())
            if hdr.get("FileFormat") != FILE_FORMAT:
                raise GrmFileError(f"'{path}' is not valid.")
            self.header = hdr
            self.command = hdr["command"]
            f.seek(hdr["sample_id"][0])
            self.sample_id = np.load(io.BytesIO(f.read(hdr["sample_id"][1])), allow_pickle=False)
            f.seek(hdr["snp_id"][0])
            self.snp_id = np.load(io.BytesIO(f.read(hdr["snp_id"][1])), allow_pickle=False)
        n = hdr["n_samp"]
        self.n = n
        self.grm = np.memmap(path, dtype=np.dtype(hdr["prec"]), mode="r", offset=hdr["grm"], shape=(n, n))

    def close(self):
        del self.grm


def merge_grm_files(filelist, out_fn=None, out_prec="double", weight=None, verbose=False,
                    block_rows=1024):
    """snpgdsMergeGRM (R/IBD.R:624-748) over GrmFile containers; the arithmetic is
    gnrGRMMerge (src/genPCA.cpp:1721-1857): a weighted row-by-row sum, or for
    ":method = IndivBeta" the un-normalise / re-normalise transform of :1737-1822."""
    if isinstance(filelist, str) or len(filelist) == 0:
        raise GrmFileError("'filelist' should be a non-empty list of file names")
    if out_prec not in _PREC:
        raise GrmFileError("'out.prec' should be one of \"double\", \"single\"")
    files = [GrmFile(p) for p in filelist]
    try:
        f0 = files[0]
        cmd = f0.command
        if cmd[0] != "snpgdsGRM":
            raise GrmFileError("The GDS files should be created by snpgdsGRM()")
        for p, f in zip(filelist, files):
            if f.command != cmd:
                raise GrmFileError(f"'{p}' has a different command.")
            if f.n != f0.n:
                raise GrmFileError(f"'{p}' has a different GRM matrix.")
        num = np.array([len(f.snp_id) for f in files], dtype=np.float64)
        if weight is None or np.asarray(weight).dtype == bool:
            if weight is not None:
                num = np.where(np.asarray(weight), num, -num)
            w = num / num.sum()
        else:
            w = np.asarray(weight, dtype=np.float64)
            if len(w) != len(files):
                raise GrmFileError("'weight' should have one entry per file")
        if verbose:
            print("Weight: " + ", ".join(f"{x:g}" for x in w))
        sid = None
        for f, wi in zip(files, w):
            s = f.snp_id
            if wi >= 0:
                sid = s if sid is None else np.concatenate([sid, s])
            else:
                sid = np.setdiff1d(sid, s) if sid is not None else sid
        n = f0.n
        beta = len(cmd) > 1 and cmd[1] == ":method = IndivBeta"
        writer = GrmWriter(out_fn, cmd, f0.sample_id, sid, out_prec) if out_fn else None
        out = writer.mm if writer else np.empty((n, n), dtype=np.float64)
        avg_out = None
        if not beta:
            for r0 in range(0, n, block_rows):
                r1 = min(n, r0 + block_rows)
                acc = np.zeros((r1 - r0, n), dtype=np.float64)
                for f, wi in zip(files, w):
                    acc += np.asarray(f.grm[r0:r1], dtype=np.float64) * wi
                out[r0:r1] = acc
        else:
            # baseline M_b of each file = half the mean off-diagonal value (:1750-1765)
            mb = []
            for f in files:
                tot = 0.0
                for r0 in range(0, n, block_rows):
                    r1 = min(n, r0 + block_rows)
                    blk = np.asarray(f.grm[r0:r1], dtype=np.float64)
                    tot += blk.sum() - blk[np.arange(r1 - r0), np.arange(r0, r1)].sum()
                mb.append(tot / (float(n) * (n - 1)) * 0.5)
            work = np.empty((n, n), dtype=np.float64) if writer else out
            for r0 in range(0, n, block_rows):
                r1 = min(n, r0 + block_rows)
                acc = np.zeros((r1 - r0, n), dtype=np.float64)
                di = (np.arange(r1 - r0), np.arange(r0, r1))
                for f, wi, b in zip(files, w, mb):
                    blk = np.asarray(f.grm[r0:r1], dtype=np.float64)
                    inv = 1.0 / (1.0 - b)
                    m = (blk * 0.5 - b) * inv * (1 - f.avg_val) + f.avg_val
                    m[di] = (blk[di] - 1 - b) * inv * (1 - f.avg_val) + f.avg_val
                    acc += m * wi
                work[r0:r1] = acc
            mn = work.min()
            avg_out = (work.sum() - np.trace(work)) / (float(n) * (n - 1))
            scale = 2 / (1 - mn)
            work -= mn
            work *= scale
            d = np.arange(n)
            work[d, d] = work[d, d] * 0.5 + 1
            if writer:
                out[:, :] = work
        if writer:
            writer.close(avg_out)
            return None
        rv = {"sample.id": f0.sample_id, "snp.id": sid, "grm": out}
        if beta:
            rv["avg_val"] = avg_out
        return rv
    finally:
        for f in files:
            f.close()
