"""snprelate_b200 -- B200-native (sm_100a) implementation of SNPRelate's pairwise
N x N relatedness-matrix path (snpgdsGRM / snpgdsPCA / snpgdsEIGMIX / snpgdsIBS /
snpgdsIBSNum / snpgdsIBDMoM / snpgdsIBDKING / snpgdsIndivBeta, plus the loadings /
projection / SNP-PC correlation steps around the eigen-decomposition).

The product is ``libsnprel_b200.so`` (hand-written CUDA behind a C ABI, see
``include/snprel_b200.h``).  This package is the thin host layer that mirrors the
reference's R interface (same function names, argument meaning and error
behaviour; ``R/IBD.R``, ``R/PCA.R``, ``R/IBS.R``, ``R/Internal.R``) on top of that
ABI through ctypes.  There is no CPU fallback: importing works anywhere, but
every compute call raises ``SNPRelError`` without the CUDA library and a B200.
"""
from ._lib import SNPRelError, Context, MultiContext, load_library, library_path  # noqa: F401
from .api import (  # noqa: F401
    GenotypeData,
    snpgdsGRM,
    snpgdsPCA,
    snpgdsEIGMIX,
    snpgdsPCACorr,
    snpgdsPCASNPLoading,
    snpgdsPCASampLoading,
    snpgdsIBS,
    snpgdsIBSNum,
    snpgdsIBDMoM,
    snpgdsMergeGRM,
    snpgdsIBDKING,
    snpgdsIndivBeta,
    snpgdsSNPRateFreq,
)

__all__ = [
    "SNPRelError", "Context", "MultiContext", "load_library", "library_path", "GenotypeData",
    "snpgdsGRM", "snpgdsPCA", "snpgdsEIGMIX", "snpgdsPCACorr", "snpgdsPCASNPLoading", "snpgdsPCASampLoading", "snpgdsIBS", "snpgdsIBSNum", "snpgdsIBDMoM", "snpgdsMergeGRM",
    "snpgdsIBDKING", "snpgdsIndivBeta", "snpgdsSNPRateFreq",
]
