"""Pair-counter variants: debug flags 0x400 / 0x800 / 0xC00 send the last 1 / 2 / 3 streams of every estimator
straight to POPC (no carry-save adder).  Checks bit-exactness on a small problem, then times 16384 x 262144."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import snprelate_b200 as S
from snprelate_b200._lib import EST_IBS, EST_KING_ROBUST, EST_BETA
from oracle import snprel_oracle as O

g = O.synth_geno(500, 3000, seed=4, miss_rate=0.02)
refs = {"ibs": O.ibs_counts(g), "king": O.king_robust_counts(g), "beta": O.beta_counts(g)}
big = S.Context(0)
big.geno_begin(16384, 262144)
big.geno_synth(262144)
for flags in (0, 0x400, 0x800, 0xC00):
    with S.Context(0) as c:
        c.debug_flags(flags)
        c.geno_begin(500, 3000)
        c.geno_push_u8(g)
        ok = (np.array_equal(np.stack(c.ibs_num()), refs["ibs"]) and np.array_equal(c.king_robust_counts(), refs["king"])
              and np.array_equal(c.indiv_beta_counts(), refs["beta"]))
    big.debug_flags(flags)
    line = f"direct streams {flags >> 10}: exact {ok} |"
    for name, est in (("ibs", EST_IBS), ("king", EST_KING_ROBUST), ("beta", EST_BETA)):
        big.invalidate()
        ms = min(big.time_accumulate(est, 1) for _ in range(2))
        hot, _, units = big.last_hot_kernel()
        line += f" {name} {hot:7.1f} ms {units / hot * 1e3:.3e}"
    print(line, flush=True)
