"""Several GPUs behind ONE handle and one host process (snprel_multi_*, csrc/multi.cu): SNP-block
sharding, per-device accumulation on library threads, hand-written peer-memory reduction, finishing
on the root device.  On a one-GPU box the same path runs with the device named twice ([0, 0]: two
contexts, two SNP shards, the reduction reads the "peer" through ordinary device pointers); with two
or more GPUs it also runs over NVLink peer access.  Integers bit exact, GRM entries to 1e-10."""
import os

import numpy as np
import pytest
import torch

import snprelate_b200 as S
from oracle import ref_lib as R
from oracle import snprel_oracle as O
from snprelate_b200._lib import EST_IBS, EST_KING_ROBUST, EST_BETA

pytestmark = pytest.mark.gpu
TOL = 1e-10


def relerr(got, ref):
    return float(np.nanmax(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)))


def device_sets():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    sets = [[0, 0], [0, 0, 0]]
    if n >= 2:
        sets.append(list(range(min(n, 4))))
    return sets


@pytest.fixture(scope="module")
def data():
    return O.synth_geno(700, 5000, seed=17, miss_rate=0.02, maf_lo=0.01)


@pytest.mark.parametrize("devices", device_sets(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_sharded_estimators_match_the_oracle(data, devices):
    g = data
    with S.MultiContext(devices) as m:
        m.geno_begin(g.shape[1], g.shape[0])
        m.geno_push_u8(g[:1234])                     # pushes straddle the shard boundaries
        m.geno_push_u8(g[1234:])
        c = m.ctx(0)
        shard_snps = [m.ctx(i).geno_dim()[1] for i in range(len(devices))]
        assert sum(shard_snps) == g.shape[0] and all(s > 0 for s in shard_snps)
        for method, ref in (("GCTA", O.grm_gcta(g)), ("Eigenstrat", O.grm_eigenstrat(g)), ("EIGMIX", O.grm_eigmix(g))):
            m.accumulate(method)
            assert relerr(c.grm(method)[0], ref) < TOL, method
        m.accumulate("Eigenstrat", bayesian=True)
        r = c.pca(eigen_cnt=4, bayesian=True, need_genmat=True)
        assert relerr(r["genmat"], O.pca_genmat(g, bayesian=True)[0]) < TOL
        m.accumulate(EST_IBS)
        assert np.array_equal(np.stack(c.ibs_num()), O.ibs_counts(g))
        ms, nbytes = m.last_reduce()
        assert nbytes > 0 and ms >= 0
        m.accumulate(EST_KING_ROBUST, root=len(devices) - 1)          # any device can be the root
        assert np.array_equal(m.ctx(len(devices) - 1).king_robust_counts(), O.king_robust_counts(g))
        m.accumulate(EST_BETA, root=-1)                                # all-reduce: every device holds the result
        for i in range(len(devices)):
            assert np.array_equal(m.ctx(i).indiv_beta_counts(), O.beta_counts(g))
        m.set_count_engine("tensor")
        m.accumulate(EST_IBS)
        assert np.array_equal(np.stack(c.ibs_num()), O.ibs_counts(g))


def test_row_windows_over_several_devices(data):
    """Tiled N x N output: every window is accumulated by all devices and reduced to its finisher."""
    g = data
    n = g.shape[1]
    ref = O.grm_gcta(g)
    packed_ref = ref[np.triu_indices(n)]
    with S.MultiContext([0, 0]) as m:
        m.geno_begin(n, g.shape[0])
        m.geno_push_u8(g)
        out = np.empty(n * (n + 1) // 2)
        pos = 0
        for w, r0 in enumerate(range(0, n, 256)):
            m.set_row_window(r0, 256)
            root = w % 2
            m.accumulate("GCTA", root=root)
            part = m.ctx(root).grm("GCTA", packed=True)[0]
            out[pos: pos + part.size] = part
            pos += part.size
        assert pos == out.size
        assert relerr(out, packed_ref) < TOL


def test_tiny_workspace_uses_fewer_devices():
    g = O.synth_geno(50, 100, seed=2, miss_rate=0.05)        # one 128-SNP block: a single active shard
    with S.MultiContext([0, 0, 0]) as m:
        m.geno_begin(50, 100)
        m.geno_push_u8(g)
        m.accumulate("GCTA")
        assert relerr(m.ctx(0).grm("GCTA")[0], O.grm_gcta(g)) < TOL


@pytest.mark.skipif(not R.rshim_available(), reason="oracle/_ref/libsnprelate_b200_rshim.so not built")
def test_r_entry_points_on_several_devices(data):
    """The R binding with SNPREL_DEVICES set: gnrGRM / gnrPCA / gnrEigMix / gnrIBSNum / gnrIBD_KING_Robust
    shard over the devices inside one process.  Runs in a subprocess: the binding reads the variable once."""
    import subprocess
    import sys
    code = r'''
import numpy as np
from oracle import ref_lib as R, snprel_oracle as O
g = O.synth_geno(700, 5000, seed=17, miss_rate=0.02, maf_lo=0.01)
w = R.RefWorkspace(g, R.RSHIM_PATH)
rel = lambda a, b: float(np.nanmax(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))
assert rel(w.grm("GCTA"), O.grm_gcta(g)) < 1e-10
assert rel(w.grm("IndivBeta"), O.grm_indivbeta(O.beta_counts(g))[0]) < 1e-12
assert np.array_equal(np.stack(w.ibs_num()), O.ibs_counts(g))
p = w.pca(eigen_cnt=0)            # genmat only: no cuSOLVER start-up in this short-lived process
assert rel(p["genmat"], O.pca_genmat(g)[0]) < 1e-10
ibd, af = w.eigmix(diagadj=True)            # (gnrEigMix with eigen.cnt = 0)
oibd, oaf = O.eigmix_ibd(g, diagadj=True)
assert rel(ibd, oibd) < 1e-10 and np.max(np.abs(af - oaf)) < 1e-15
a, b = w.king_robust()
ra, rb = O.king_robust(O.king_robust_counts(g))
assert np.array_equal(a, ra) and np.allclose(b, rb, rtol=0, atol=0, equal_nan=True)
print("multi-device R shim ok")
'''
    env = dict(os.environ, SNPREL_DEVICES="0,0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "multi-device R shim ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("devices", device_sets()[::2], ids=lambda d: "dev" + "".join(map(str, d)))
def test_tiled_grm_with_sharded_loading_and_gather(data, devices):
    """N x N output tiled across devices: every device reserves the whole SNP range, only the
    owner receives a pushed block, the gather completes the copies device to device, windows are dealt
    w mod n and delivered to the sink in order."""
    g = data
    n = g.shape[1]
    with S.MultiContext(devices) as m:
        m.geno_begin_replicated(n, g.shape[0])
        m.geno_push_u8(g[:3000])
        m.geno_push_u8(g[3000:])
        own = [m.ctx(i).geno_dim()[1] for i in range(len(devices))]          # before the gather: the append positions
        m.geno_gather()
        for i in range(len(devices)):
            assert np.array_equal(m.ctx(i).geno_copy_u8(), g)                 # every device holds all SNPs
        for method, ref in (("GCTA", O.grm_gcta(g)), ("Eigenstrat", O.grm_eigenstrat(g)), ("EIGMIX", O.grm_eigmix(g))):
            out = m.grm_tiled(method, window_rows=256)
            assert relerr(out, ref[np.triu_indices(n)]) < TOL, method
            firsts = [f for f, _ in m.tiled_calls]
            assert firsts == sorted(firsts) and len(firsts) == 3 and sum(c for _, c in m.tiled_calls) == out.size
        out = m.grm_tiled("GCTA")                                             # automatic window height
        assert relerr(out, O.grm_gcta(g)[np.triu_indices(n)]) < TOL
        with pytest.raises(S.SNPRelError):
            m.accumulate("GCTA")
    assert len(own) == len(devices)
