"""Bring-up probe for the GPU box: exercises every kernel once on tiny inputs and
prints enough structure to debug layout problems remotely.  Not a test."""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import snprel_oracle as O          # noqa: E402  (checker only)
import snprelate_b200 as S                      # noqa: E402


def table_gram_ref(g, tabA, tabB):
    gi = np.minimum(g, 3).astype(np.int64)              # [M, N]
    a = np.take_along_axis(tabA.astype(np.int64), gi, axis=1)   # [M, N]
    b = tabB.astype(np.int64)[gi]
    return a.T @ b


def section(name):
    print(f"\n===== {name} =====", flush=True)


def main():
    ctx = S.Context(0)
    rng = np.random.default_rng(1)

    section("workspace roundtrip / synth / stats")
    N, M = 301, 517
    g = O.synth_geno(N, M, seed=11, miss_rate=0.02)
    ctx.geno_begin(N, M)
    ctx.geno_push_u8(g[:200])
    ctx.geno_push_u8(g[200:])
    back = ctx.geno_copy_u8()
    print("push/copy roundtrip equal:", np.array_equal(back, g))
    ctx.geno_begin(N, M)
    ctx.geno_synth(M, seed=11, miss_rate=0.02)
    print("device synth == oracle synth:", np.array_equal(ctx.geno_copy_u8(), g))
    af, maf, mr = ctx.snp_ratefreq()
    s, num = O.snp_stats(g)
    print("af max err:", np.nanmax(np.abs(af - s / (2.0 * num))), "mr max err:", np.max(np.abs(mr - (1 - num / N))))

    section("packed-bit kernels")
    try:
        i0, i1, i2 = ctx.ibs_num()
        ref = O.ibs_counts(g)
        print("IBS exact:", np.array_equal(np.stack([i0, i1, i2]), ref))
        kc = ctx.king_robust_counts()
        print("KING counts exact:", np.array_equal(kc, O.king_robust_counts(g)))
        bc = ctx.indiv_beta_counts()
        print("Beta counts exact:", np.array_equal(bc, O.beta_counts(g)))
    except Exception:
        traceback.print_exc()

    section("tcgen05 table gram: one-hot layout probes")
    for flags in (0, 1):
        try:
            ctx2 = S.Context(0)
            ctx2.debug_flags(flags)
            Np, Mp = 300, 40
            gp = np.zeros((Mp, Np), dtype=np.uint8)
            i0_, j0_, l0_ = 37, 150, 13
            gp[l0_, i0_] = 1
            gp[l0_, j0_] = 2
            tabA = np.zeros((Mp, 4), dtype=np.int8)
            tabA[l0_, 1] = 5          # A side sees sample i0 (code 1)
            tabB = np.array([0, 0, 3, 0], dtype=np.int8)   # B side sees sample j0 (code 2)
            ctx2.geno_begin(Np, Mp)
            ctx2.geno_push_u8(gp)
            out = ctx2.table_gram(tabA, tabB)
            nz = np.argwhere(out != 0)
            print(f"flags={flags}: expect single nonzero 15 at ({i0_},{j0_}); got {len(nz)} nonzeros:",
                  [(int(a), int(b), int(out[a, b])) for a, b in nz[:12]])
            # SNP-position probe: every SNP l contributes l+1 at (i0, j0)
            gp = np.zeros((Mp, Np), dtype=np.uint8)
            gp[:, i0_] = 1
            gp[:, j0_] = 2
            tabA = np.zeros((Mp, 4), dtype=np.int8)
            tabA[:, 1] = np.arange(1, Mp + 1)
            ctx2.geno_begin(Np, Mp)
            ctx2.geno_push_u8(gp)
            out = ctx2.table_gram(tabA, tabB)
            print(f"flags={flags}: K-sum probe expect {3 * Mp * (Mp + 1) // 2} at ({i0_},{j0_}): got",
                  int(out[i0_, j0_]), "nonzeros:", int((out != 0).sum()))
            # random exactness
            Nr, Mr = 301, 517
            gr = O.synth_geno(Nr, Mr, seed=5, miss_rate=0.03)
            tA = rng.integers(-128, 128, size=(Mr, 4)).astype(np.int8)
            tB = np.array([0, 1, 2, 0], dtype=np.int8)
            ctx2.geno_begin(Nr, Mr)
            ctx2.geno_push_u8(gr)
            t0 = time.time()
            out = ctx2.table_gram(tA, tB)
            ref = table_gram_ref(gr, tA, tB)
            bad = int((out != ref).sum())
            print(f"flags={flags}: random table gram mismatches: {bad} / {out.size}  ({time.time() - t0:.3f}s)")
            if bad and bad < out.size:
                w = np.argwhere(out != ref)[:5]
                print("   first mismatches:", [(int(a), int(b), int(out[a, b]), int(ref[a, b])) for a, b in w])
            ctx2.close()
        except Exception:
            traceback.print_exc()

    section("GRM family vs oracle (tcgen05 path)")
    try:
        ctx.geno_begin(N, M)
        ctx.geno_push_u8(g)
        for method, ref in (("GCTA", O.grm_gcta(g)), ("Eigenstrat", O.grm_eigenstrat(g)),
                            ("EIGMIX", O.grm_eigmix(g)), ("Corr", O.grm_corr(g))):
            got, _ = ctx.grm(method)
            err = np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1))
            print(f"{method}: max err {err:.3e}")
        r = ctx.pca(eigen_cnt=4, need_genmat=True)
        ev, evec = O.pca_eigen(r["genmat"], 4)
        print("PCA eigenval err:", np.max(np.abs(ev - r["eigenval"][:4])))
        k0, k1 = ctx.king_homo()
        rk0, rk1 = O.king_homo(g)
        print("KING-homo err:", np.nanmax(np.abs(k0 - rk0)), np.nanmax(np.abs(k1 - rk1)))
    except Exception:
        traceback.print_exc()
    print("kernel launches:", ctx.kernel_launches())


if __name__ == "__main__":
    main()
