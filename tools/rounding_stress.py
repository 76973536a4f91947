"""The default ('auto') rounding mode away from the bench's data: rare alleles, heavy missingness, few / many
samples -- format, passes and the error of GCTA and Eigenstrat entries at scattered samples against the oracle,
for round-to-nearest and the default.

    python tools/rounding_stress.py
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import snprelate_b200 as S
from oracle import snprel_oracle as O       # checker only

CASES = [   # n_samp, n_snp, maf_lo, maf_hi, miss
    (4096, 300000, 0.005, 0.5, 0.02),
    (2048, 600000, 0.01, 0.1, 0.001),
    (8192, 150000, 0.2, 0.5, 0.05),
    (1500, 1000000, 0.05, 0.5, 0.0),
]
K = 40
worst = 0.0
for n, m, lo, hi, miss in CASES:
    idx = O.scattered_samples(n, K, seed=9)
    sub = O.synth_geno(0, m, seed=77, maf_lo=lo, maf_hi=hi, miss_rate=miss, samples=idx)
    ix = np.ix_(idx, idx)
    with S.Context(0) as c:
        c.geno_begin(n, m)
        c.geno_synth(m, seed=77, maf_lo=lo, maf_hi=hi, miss_rate=miss)
        sel, _ = c.select_snp_base(remove_mono=True)  # monomorphic SNPs carry no weight in the reference either
        af, _, _ = c.snp_ratefreq()
        kept = c.geno_dim()[1]
        sub_k = sub[sel]
        gref = O.subset_entries(sub_k, af, "GCTA")
        for mode in ("nearest", "auto"):
            c.set_rounding(mode)
            r = c.pca(genmat_only=True)
            pl = c.last_plan()
            eref = O.subset_entries(sub_k, af, "Eigenstrat", n_total=n, trace=r["TraceXTX"])
            e1 = float(np.max(np.abs(r["genmat"][ix] - eref) / np.maximum(np.abs(eref), 1.0)))
            g = c.grm("GCTA")[0]
            pg = c.last_plan()
            e2 = float(np.max(np.abs(g[ix] - gref) / np.maximum(np.abs(gref), 1.0)))
            worst = max(worst, e1, e2)
            print(f"n {n:5d} m {kept:7d} maf [{lo}, {hi}] miss {miss}: {mode:8s} Eigenstrat T{pl.digits} R{pl.digits_w} f {pl.frac_bits}/{pl.frac_bits_w} "
                  f"rounding {pl.rounding} err {e1:.2e} | GCTA T{pg.digits} R{pg.digits_w} D{pg.digits_d} rounding {pg.rounding} err {e2:.2e}", flush=True)
print(f"worst {worst:.3e} (tolerance 1e-10)")
assert worst < 1e-10
