"""One accumulate step for ncu (launch list / full capture)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import snprelate_b200 as S
n = int(sys.argv[1]); m = int(sys.argv[2]); est = int(sys.argv[3]); miss = float(sys.argv[4]) if len(sys.argv) > 4 else 0.005
ctx = S.Context(0)
ctx.geno_begin(n, m)
ctx.geno_synth(m, miss_rate=miss)
ms = ctx.time_accumulate(est, 1)
print("step ms", ms, "hot", ctx.last_hot_kernel())
