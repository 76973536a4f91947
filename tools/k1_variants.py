"""K1 experiment knobs (snprel_debug_flags): 16 = the round-1 16-byte genotype boxes (default: 32-byte swizzled), 32 = no L2 promotion,
64 = 256-byte L2 promotion, 128 = group-major item order, 256 = no SNP segments (round-2 first half), 0xN000 = N segments.  Checks each variant against the oracle on a
small problem, then times it at the bench size (or the size given).  `--one FLAGS N M` runs a single
accumulate (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import snprelate_b200 as S

if len(sys.argv) > 2 and sys.argv[1] == "--one":
    flags = int(sys.argv[2]); n = int(sys.argv[3]); m = int(sys.argv[4])
    c = S.Context(0); c.debug_flags(flags); c.geno_begin(n, m); c.geno_synth(m, miss_rate=0.005)
    print(flags, c.time_accumulate(0, 1), c.last_hot_kernel())
    sys.exit(0)

from oracle import snprel_oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 16, 32, 64, 16 | 32, 16 | 64, 128, 128 | 16, 128 | 16 | 32]
g = O.synth_geno(700, 5000, seed=3, miss_rate=0.01)
ref = O.grm_gcta(g)
big = S.Context(0)
big.geno_begin(n, m)
big.geno_synth(m, miss_rate=0.005)
for f in variants:
    with S.Context(0) as c:
        c.debug_flags(f)
        c.geno_begin(700, 5000)
        c.geno_push_u8(g)
        got = c.grm("GCTA")[0]
        err = float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1)))
    big.debug_flags(f)
    big.invalidate()
    hot = []
    for _ in range(3):
        big.time_accumulate(0, 1)
        hot.append(big.last_hot_kernel()[0])
    print(f"flags {f:3d}: small GCTA err {err:.2e}  hot ms {min(hot):.1f} (all {', '.join(f'{h:.1f}' for h in hot)})", flush=True)
