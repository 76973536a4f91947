import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def unpack2(geno2b, nsamp):
    """row-padded 2-bit [nsnp, ceil(nsamp/4)] -> uint8 codes [nsnp, nsamp]."""
    g = np.stack([(geno2b >> (2 * k)) & 3 for k in range(4)], axis=-1)
    return g.reshape(geno2b.shape[0], -1)[:, :nsamp].astype(np.uint8)


@pytest.fixture(scope="session")
def hapmap():
    z = np.load(os.path.join(GOLDEN, "hapmap_geno.npz"))
    nsamp = int(z["nsamp"])
    return dict(geno=unpack2(z["geno2b"], nsamp), geno2b=z["geno2b"], nsamp=nsamp,
                sample_id=z["sample_id"], snp_id=z["snp_id"],
                chromosome=z["chromosome"])


@pytest.fixture(scope="session")
def goldens():
    return dict(np.load(os.path.join(GOLDEN, "reference_goldens.npz")))
