"""Per-call wall time of the R entry points of the R-shim test library (diagnostic)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ref_lib as R, snprel_oracle as O
data = O.synth_geno(211, 1777, seed=31, miss_rate=0.03, maf_lo=0.01)
t = time.time(); w = R.RefWorkspace(data, R.RSHIM_PATH); print("set_geno %.2f" % (time.time() - t), flush=True)
for name, fn in [("grm GCTA", lambda: w.grm("GCTA")), ("grm GCTA again", lambda: w.grm("GCTA")), ("ibs_num", lambda: w.ibs_num()),
                 ("ibs_ave", lambda: w.ibs_ave()), ("king_robust", lambda: w.king_robust()), ("king_homo", lambda: w.king_homo()),
                 ("indiv_beta", lambda: w.indiv_beta()), ("ibd_mom", lambda: w.ibd_mom()), ("eigmix", lambda: w.eigmix()),
                 ("pca k=5", lambda: w.pca(eigen_cnt=5)), ("pca k=5 again", lambda: w.pca(eigen_cnt=5))]:
    t = time.time(); fn(); print("%s %.2f s" % (name, time.time() - t), flush=True)
