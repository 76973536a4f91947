"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Decoders for the reference's own fixture and golden vectors:

* ``load_hapmap_gds``  -- inst/extdata/hapmap_geno.gds (CoreArray container; the
  genotype node is an uncompressed dBit2 stream, SNP-major / sample-fastest,
  4 genotypes per byte LSB first, 3 = missing).
* ``load_rdata``       -- inst/unitTests/valid/*.RData (xz + "RDX2" XDR
  serialization v2).

These run only in the build container (where /root/reference is mounted) from
``oracle/make_golden.py``; the decoded arrays are committed under
``tests/golden/`` so nothing on the GPU box needs /root/reference.
"""
from __future__ import annotations

import hashlib
import lzma
import re
import struct
import zlib

import numpy as np

HAPMAP_SHA256 = "42350a8885a10c5b804474e99faff7077203b76040c68b1f6c5c56d9c02974e0"
HAPMAP_NSNP = 9088
HAPMAP_NSAMP = 279


def _zlib_streams(buf: bytes):
    """Yield (offset, decompressed bytes) for every complete raw zlib stream."""
    out = []
    for m in re.finditer(b"\x78[\x01\x5e\x9c\xda]", buf):
        o = m.start()
        try:
            d = zlib.decompressobj()
            data = d.decompress(buf[o:o + 400000])
            if d.eof and len(data) > 8:
                out.append((o, data))
        except zlib.error:
            pass
    return out


def unpack_bit2(raw: np.ndarray, count: int) -> np.ndarray:
    """Continuous 2-bit stream (LSB first) -> uint8 codes 0..3."""
    raw = np.asarray(raw, dtype=np.uint8)
    g = np.empty((raw.size, 4), dtype=np.uint8)
    for k in range(4):
        g[:, k] = (raw >> (2 * k)) & 3
    return g.reshape(-1)[:count]


def load_hapmap_gds(path: str) -> dict:
    """Return dict(geno[u8 nsnp x nsamp, 0/1/2/3], sample_id, snp_id, chromosome, position)."""
    buf = open(path, "rb").read()
    if hashlib.sha256(buf).hexdigest() != HAPMAP_SHA256:
        raise ValueError("unexpected hapmap_geno.gds content")
    tag = buf.find(b"sample.order")
    if tag < 0:
        raise ValueError("genotype node attribute 'sample.order' not found")
    nbytes = (HAPMAP_NSNP * HAPMAP_NSAMP * 2 + 7) // 8
    # the payload starts right after the attribute block; locate it by the
    # known genotype histogram rather than a hard-coded offset
    want = (920248, 657267, 947899, 10138)
    start = None
    for cand in range(tag + 12, tag + 128):
        raw = np.frombuffer(buf, dtype=np.uint8, count=nbytes, offset=cand)
        g = unpack_bit2(raw, HAPMAP_NSNP * HAPMAP_NSAMP)
        if tuple(np.bincount(g, minlength=4)) == want:
            start = cand
            break
    if start is None:
        raise ValueError("genotype payload not located")
    geno = g.reshape(HAPMAP_NSNP, HAPMAP_NSAMP).copy()

    streams = _zlib_streams(buf[:tag])
    sample_id = snp_id = chrom = pos = None
    for _, data in streams:
        if sample_id is None and data.count(b"\x00") == HAPMAP_NSAMP and len(data) < 4000:
            sample_id = [s.decode() for s in data.split(b"\x00")[:HAPMAP_NSAMP]]
        elif len(data) == 4 * HAPMAP_NSNP:
            arr = np.frombuffer(data, dtype="<i4")
            if snp_id is None and np.array_equal(arr, np.arange(1, HAPMAP_NSNP + 1)):
                snp_id = arr.copy()
            elif pos is None:
                pos = arr.copy()
        elif len(data) == HAPMAP_NSNP and chrom is None:
            chrom = np.frombuffer(data, dtype=np.uint8).copy()
    if sample_id is None or snp_id is None or chrom is None:
        raise ValueError("annotation nodes not decoded")
    return dict(geno=geno, sample_id=sample_id, snp_id=snp_id, chromosome=chrom,
                position=pos, payload_offset=start)


# ---------------------------------------------------------------------------
# R serialization (version 2, XDR) reader -- the subset the goldens use
# ---------------------------------------------------------------------------

class _RReader:
    def __init__(self, data: bytes):
        self.b = data
        self.o = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if t == 254:      # NILVALUE
            return None
        if t == 253:      # GLOBALENV
            return "<globalenv>"
        if t == 255:      # REFSXP
            return self.refs[(flags >> 8) - 1]
        if t == 1:        # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 2:        # LISTSXP (pairlist) -> dict / list of (tag, value)
            out = []
            while True:
                if has_attr:
                    self.item()
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                nflags = struct.unpack_from(">i", self.b, self.o)[0]
                if (nflags & 0xFF) != 2:
                    self.item()  # terminating NILVALUE (or whatever CDR is)
                    break
                flags = self.i32()
                has_attr = bool(flags & 0x200)
                has_tag = bool(flags & 0x400)
            return out
        if t == 9:        # CHARSXP
            n = self.i32()
            if n < 0:
                return None
            s = self.b[self.o:self.o + n].decode("latin-1")
            self.o += n
            return s
        if t in (10, 13):  # LGLSXP / INTSXP
            n = self.i32()
            v = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.o).astype(np.int32)
            self.o += 4 * n
        elif t == 14:      # REALSXP
            n = self.i32()
            v = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.o).astype(np.float64)
            self.o += 8 * n
        elif t == 16:      # STRSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
        elif t == 19:      # VECSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
        else:
            raise NotImplementedError(f"SEXP type {t} at {self.o}")
        if has_attr:
            attrs = dict(self.item())
            if "dim" in attrs and isinstance(v, np.ndarray):
                d = tuple(int(x) for x in attrs["dim"])
                v = v.reshape(d, order="F")
            if "names" in attrs and isinstance(v, list):
                v = dict(zip(attrs["names"], v))
        return v


def load_rdata(path: str) -> dict:
    data = lzma.decompress(open(path, "rb").read())
    if not data.startswith(b"RDX2\nX\n"):
        raise ValueError("not an RDX2/XDR file")
    r = _RReader(data)
    r.o = 7
    r.i32(); r.i32(); r.i32()
    top = r.item()
    return dict(top)
