"""CPU model of the packed-bit pair kernels' arithmetic (snprelate_b200/csrc/bitcount.cu): the
three-input truth tables of pair_streams<EST>, the 3:2 carry-save step and the counter mapping,
evaluated with Python integers on bit planes built like planes_kernel (incl. the all-missing padding
words) and compared with the oracle's counters.  It pins the LOGIC of the kernel on the CPU (the
GPU tests pin the kernel itself); the LUT expressions below mirror the lop3<> template arguments."""
import numpy as np

from oracle import snprel_oracle as O

M32 = 0xFFFFFFFF
TA, TB, TC = 0xF0, 0xCC, 0xAA


def N8(x):
    return (~x) & 0xFF


def lop3(lut, a, b, c):
    """PTX lop3.b32: bit (a<<2 | b<<1 | c) of the immediate selects the result"""
    r = 0
    for idx in range(8):
        if (lut >> idx) & 1:
            r |= (a if (idx >> 2) & 1 else ~a) & (b if (idx >> 1) & 1 else ~b) & (c if idx & 1 else ~c)
    return r & M32


def streams(est, a1, a2, b1, b2):
    vb = (b1 | ~b2) & M32
    hb = b1 & ~b2 & M32
    mask = lop3((TA | N8(TB)) & TC, a1, a2, vb)
    if est == "ibs":
        t = lop3((TA ^ TB) & TC, a1, b1, mask)
        u = lop3(N8(TA ^ TB) & TC, a1, b1, mask)
        return [lop3(TA & (TB ^ TC), t, a2, b2), lop3(TA & N8(TB ^ TC), u, a2, b2), mask]
    if est == "king":
        t = lop3((TA ^ TB) & TC, a1, b1, mask)
        pa = lop3(TA ^ TB ^ TC, a1, a2, b1)
        return [lop3(TA & (TB ^ TC), t, a2, b2), mask, lop3((TA ^ TB) & TC, pa, b2, mask),
                lop3(TA & N8(TB) & TC, a1, a2, vb), lop3(TA & (TB | N8(TC)), hb, a1, a2)]
    h1 = lop3((TA ^ TB) | TC, a1, a2, hb)
    v = mask & ~h1 & M32
    return [h1 & mask, lop3(TA & N8(TB ^ TC), v, a1, b1), mask]


def test_pair_kernel_logic_reproduces_the_oracle_counters():
    n, m = 11, 203                                   # 7 words of 32 SNPs: 21 padding SNPs, ragged carry-save group
    g = O.synth_geno(n, m, seed=9, miss_rate=0.15, maf_lo=0.05)
    nw = (m + 31) // 32
    enc = {0: (0, 0), 1: (1, 0), 2: (1, 1), 3: (0, 1)}          # PackSNPGeno1b, src/dGenGWAS.cpp:1429-1475
    p1 = [[0] * nw for _ in range(n)]
    p2 = [[0] * nw for _ in range(n)]
    for i in range(n):
        for l in range(nw * 32):
            b1, b2 = enc[int(g[l, i]) if l < m else 3]
            p1[i][l // 32] |= b1 << (l % 32)
            p2[i][l // 32] |= b2 << (l % 32)
    pop = lambda v: bin(v).count("1")
    iu = np.triu_indices(n)
    for est, nc in (("ibs", 3), ("king", 5), ("beta", 3)):
        cnt = np.zeros((nc, n, n), dtype=np.int64)
        for i in range(n):
            for j in range(i, n):
                for w0 in range(0, nw, 3):
                    st = [streams(est, p1[i][w], p2[i][w], p1[j][w], p2[j][w]) for w in range(w0, min(w0 + 3, nw))]
                    st += [[0] * nc] * (3 - len(st))
                    for k in range(nc):
                        s = st[0][k] ^ st[1][k] ^ st[2][k]
                        c = (st[0][k] & st[1][k]) | (st[2][k] & (st[0][k] ^ st[1][k]))
                        cnt[k, i, j] += pop(s) + 2 * pop(c)
        if est == "ibs":
            got = [cnt[0], cnt[2] - cnt[0] - cnt[1], cnt[1]]
            ref = O.ibs_counts(g)
        elif est == "king":
            got = [cnt[0], cnt[1], cnt[2] + 4 * cnt[0], cnt[3], cnt[4]]
            ref = O.king_robust_counts(g)
        else:
            got = [cnt[0] + 2 * cnt[1], cnt[2]]
            ref = O.beta_counts(g)
        for a, b in zip(got, ref):
            assert np.array_equal(a[iu], b[iu]), est
