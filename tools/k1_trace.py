"""Where an item of K1 spends its cycles (snprel_debug_flags 1): median / mean clocks per phase."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import snprelate_b200 as S
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
c = S.Context(0); c.geno_begin(n, m); c.geno_synth(m)
c.time_accumulate(0, 1)
c.debug_flags(1)
ms = c.time_accumulate(0, 1)
hot = c.last_hot_kernel()[0]
t = c.k1_trace()
t = t[(t[:, 0] > 0) & (t[:, 6] > 0)]
mode = (t[:, 7] >> 32) & 0xFF
cols = t[:, 7] >> 40
t = t.copy()
t[:, 7] &= 0xFFFFFFFF
for md in (0, 1):
    for nc in sorted(set(cols[mode == md].tolist())):
        sel = (mode == md) & (cols == nc)
        if sel.sum() == 0:
            continue
        per_stage = (t[sel, 3] - t[sel, 2]) / np.maximum(t[sel, 7], 1)
        print(f"mode {md} (B columns {nc:3d}): {int(sel.sum()):6d} items, main loop {np.median(per_stage):7.1f} clocks per stage (median), "
              f"p90 {np.percentile(per_stage, 90):7.1f}, epilogue {np.median(t[sel, 5] - t[sel, 4]):8.0f}")
full = t[t[:, 7] >= np.median(t[:, 7])]
names = ["set-up (entry -> cluster sync)", "ramp (-> first MMA)", "main loop (-> last commit issued)", "drain (-> MMAs complete)",
         "epilogue", "exit (fences, cluster sync, dealloc)"]
d = np.diff(full[:, :7], axis=1)
tot = full[:, 6] - full[:, 0]
print(f"K1 {hot:.1f} ms, {len(t)} items traced, {len(full)} full-length ({int(np.median(full[:, 7]))} stages); clocks per item:")
for k, nm in enumerate(names):
    print(f"  {nm:38s} median {np.median(d[:, k]):10.0f}  mean {d[:, k].mean():10.0f}  ({100 * d[:, k].mean() / tot.mean():5.2f} %)")
print(f"  {'whole item':38s} median {np.median(tot):10.0f}  mean {tot.mean():10.0f}")
clk = 1.0e-3 * hot * 1.69e9
slots = 74
print(f"  sum of item times / (kernel clocks x {slots} slots) = {t[:, 6].astype(float).sum() - t[:, 0].astype(float).sum():.3e} / {clk * slots:.3e}"
      f" = {(t[:, 6] - t[:, 0]).sum() / (clk * slots):.3f} (the rest: gaps between items on a slot, at 1.69 GHz)")
