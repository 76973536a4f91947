"""BASELINE config-2 size (10 000 samples x 1 000 000 SNPs, the bench workload) through the C ABI:
entries of the first samples against the oracle evaluated on those samples' columns (the per-SNP
statistics over ALL samples come from the device, themselves checked against the generator), plus
size-independent properties of the whole matrix.  The same checks run at configs 3-5 in
tools/config_run.py and tools/c5_tiled.py (results in profiles/).  Sorted last: ~1 minute."""
import numpy as np
import pytest

import snprelate_b200 as S
from oracle import snprel_oracle as O

pytestmark = pytest.mark.gpu

N, M, K, SEED, MISS = 10000, 1000000, 12, 20261017, 0.005


@pytest.fixture(scope="module")
def ws():
    ctx = S.Context(0)
    ctx.geno_begin(N, M)
    ctx.geno_synth(M, seed=SEED, miss_rate=MISS)
    sub = O.synth_geno(K, M, seed=SEED, miss_rate=MISS)          # samples 0..K-1 of the same data set
    yield ctx, sub
    ctx.close()


def test_generator_and_snp_statistics(ws):
    ctx, sub = ws
    back = ctx.geno_copy_2b()[:, : (K + 3) // 4]
    codes = np.stack([(back >> (2 * k)) & 3 for k in range(4)], axis=-1).reshape(M, -1)[:, :K]
    assert np.array_equal(codes, np.minimum(sub, 3))              # device generator == oracle generator
    af, maf, mr = ctx.snp_ratefreq()
    assert np.all((af > 0.02) & (af < 0.55)) and abs(float(mr.mean()) - MISS) < 2e-4
    assert abs(float(af.mean()) - 0.275) < 2e-3                   # MAF ~ U(0.05, 0.5)


def test_gcta_grm_full_size(ws):
    ctx, sub = ws
    grm, _ = ctx.grm("GCTA")
    assert grm.shape == (N, N) and np.array_equal(grm, grm.T)
    d = np.diag(grm)
    assert np.all(d > 0.9) and np.all(d < 1.1) and abs(float(d.mean()) - 1.0) < 1e-3    # E[G_ii] = 1 under HWE
    off_mean = (float(grm.sum()) - float(d.sum())) / (N * (N - 1.0))
    assert abs(off_mean + 1.0 / (N - 1)) < 1e-5                   # centred genotypes: rows sum to ~0
    af, _, _ = ctx.snp_ratefreq()
    mu = 2 * af
    poly = (af > 0) & (af < 1)
    w = np.where(poly, 1.0 / np.where(poly, af * (1 - af), 1.0), 0.0)
    z = np.where(sub <= 2, (sub - mu[:, None]) * np.sqrt(w)[:, None], 0.0)
    mm = (sub > 2).astype(np.float64)
    miss = mm * poly[:, None]
    den = miss.sum(0)[:, None] + miss.sum(0)[None, :] - miss.T @ mm
    ref = (z.T @ z) / (2.0 * (poly.sum() - den))
    got = grm[:K, :K]
    assert float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0))) < 1e-10
    pl = ctx.last_plan()
    assert pl.digits + pl.digits_w + pl.digits_d <= 10            # the fixed-point plan of DESIGN.md section 3


def test_ibs_counts_full_size(ws):
    ctx, sub = ws
    i0, i1, i2 = ctx.ibs_num()
    ref = O.ibs_counts(sub)
    assert np.array_equal(np.stack([i0[:K, :K], i1[:K, :K], i2[:K, :K]]), ref)          # bit exact
    tot = (i0.astype(np.int64) + i1 + i2)
    assert np.array_equal(tot, tot.T) and int(tot.max()) <= M
    assert np.array_equal(np.diag(i2), np.diag(tot))              # a sample is IBS2 with itself wherever it is valid
