"""Shared helpers for the test-suite (imported as a top-level module)."""
import numpy as np


def hapmap_subset(h, nsamp_sel, autosome_only=True, remove_monosnp=True,
                  maf=float("nan"), missing_rate=float("nan")):
    """The reference tests' selection: first nsamp_sel samples in file order,
    autosomes, then gnrSelSNP_Base over the selected samples
    (R/Internal.R:292-447).  Returns (geno[nsnp_sel, nsamp_sel], kept snp index)."""
    from oracle import snprel_oracle as O
    g = h["geno"][:, :nsamp_sel]
    keep = np.ones(g.shape[0], dtype=bool)
    if autosome_only:
        keep &= (h["chromosome"] >= 1) & (h["chromosome"] <= 22)
    idx = np.nonzero(keep)[0]
    sel = O.select_snp_base(g[idx], remove_monosnp, maf, missing_rate)
    idx = idx[sel]
    return np.ascontiguousarray(g[idx]), idx
