"""CPU model of the covariance path's fixed-point arithmetic under randomised rounding, checked against
the bound the library promises (no GPU needed).  The per-SNP integer column tables (grm.cu:coltab_kernel),
the counter-based draws (grm.cu:dither_u01) and the digit split are restated in numpy; every digit pass is
an exact integer Gram, so the error measured here is the error the tensor-core path produces.  The
format comes from the library's own host-only entry point (snprel_plan_format) fed with the statistics the
device would measure, and the measured error must stay below BOTH the tolerance and the Hoeffding bound
that format was chosen for -- for several independent draws."""
import math

import numpy as np
import pytest

from oracle import snprel_oracle as O
from snprelate_b200._lib import Plan, plan_format

N, M, SEED, MISS = 64, 200000, 424242, 0.005
N_FORMAT = 10000      # the format is chosen as for a 10 000-sample matrix of the same SNPs (more pairs in the union
                      # bound than the 64 samples modelled: conservative; and above the work threshold of 'auto')


def dither_u01(origin, m):
    """grm.cu:dither_u01 for SNP indices origin .. origin+m-1 and genotypes 0..2"""
    l = (np.arange(m, dtype=np.uint64) + np.uint64(origin))[:, None]
    g = np.arange(3, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        x = np.uint64(0xD17E5 | 1) ^ ((l * np.uint64(4) + g) * np.uint64(0xD1342543DE82EF95))
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return (x >> np.uint64(11)).astype(np.float64) / 9007199254740992.0


def build_model():
    g = O.synth_geno(N, M, seed=SEED, miss_rate=MISS)                  # [M, N]
    valid = g <= 2
    x = np.where(valid, g, 0).astype(np.float64)
    num = valid.sum(axis=1)
    # the real workloads take the SNP statistics over thousands of samples; with the 64 modelled here the sample
    # frequencies would scatter far beyond that, so mu / w come from the generator's population frequencies
    # (oracle.synth_geno's formula) -- what a 10 000-sample estimate converges to.  The model is self-consistent:
    # the float64 reference below uses the same mu / w.
    with np.errstate(over="ignore"):
        hp = O._splitmix64(np.uint64(SEED) ^ (np.arange(M, dtype=np.uint64) * np.uint64(0xD1342543DE82EF95)))
    p = 0.05 + 0.45 * ((hp >> np.uint64(11)).astype(np.float64) / 9007199254740992.0)
    mu = 2 * p
    w = 1.0 / (p * (1 - p))                                            # Eigenstrat weight, src/genPCA.cpp:145-181
    U = w[:, None] * (np.arange(3)[None, :] - mu[:, None])            # [M, 3]
    umax = w * np.maximum(mu, 2 - mu)
    # coltab_kernel: s follows |U|, t / s is the best approximation of mu among 31 candidates
    s_hi = np.minimum(np.floor(127.0 / np.maximum(2.0 - mu, 1e-9)), 127.0)
    s_tgt = np.minimum(s_hi, np.maximum(24.0, 127.0 * umax / 40.0))
    bs, bt, be = np.ones(M), np.rint(mu), np.full(M, 1e300)
    for q in range(31):
        sc = np.maximum(1.0, np.floor(s_tgt * (0.7 + 0.01 * q)))
        tc = np.rint(sc * mu)
        ok = ~((2 * sc - tc > 127.0) | (tc > 127.0))
        e = np.where(ok, np.abs(mu - tc / sc), np.inf)
        better = e < be
        bs, bt, be = np.where(better, sc, bs), np.where(better, tc, bt), np.where(better, e, be)
    live = (w > 0) & (num > 0)
    s, t = np.where(live, bs, 1.0), np.where(live, bt, 0.0)
    delta = mu - t / s
    T3, R3 = U / s[:, None], delta[:, None] * U
    code = np.where(valid, g, 3).astype(np.int64)
    B = np.where(valid, s[:, None] * x - t[:, None], 0.0)              # integer column channel
    mis = (~valid).astype(np.float64)
    z = np.where(valid, (x - mu[:, None]) * np.sqrt(w)[:, None], 0.0)
    C = z.T @ z
    trace = float(np.trace(C))
    stats = dict(max_abs=float(np.abs(T3).max()), max_abs_w=float(np.abs(R3).max()),
                 err_weight=float(np.abs(B).sum(axis=0).max()),
                 err_weight2=float((np.ceil(B * B / 193.0) * 193.0).sum(axis=0).max()),   # as sample_stats_kernel measures it
                 scale=trace / (N - 1), sum_bound=float((np.abs(T3).max(axis=1) * np.abs(B).max(axis=1)).sum() * 1.1),
                 total_missing=int(mis.sum()), max_missing=int(mis.sum(axis=0).max()), n_snp=M)
    return dict(T3=T3, R3=R3, code=code, B=B, mis=mis, C=C, stats=stats)


@pytest.fixture(scope="module")
def model():
    return build_model()


def _plan(stats):
    p = Plan()
    p.frac_bits = p.frac_bits_w = p.frac_bits_d = -1
    for k, v in stats.items():
        setattr(p, k, v)
    return p


def _digit_gram(q3, nd, code, right):
    """exact sum_l q[l][g_il] right[l, j] through nd balanced base-256 digit passes; (python ints, overflow flag)"""
    q4 = np.concatenate([q3, np.zeros((q3.shape[0], 1), dtype=np.int64)], axis=1)
    q = np.take_along_axis(q4, code, axis=1)
    total = np.zeros((code.shape[1], right.shape[1]), dtype=object)
    for k in range(nd):
        d = ((q + 128) & 255) - 128
        q = (q - d) >> 8
        total = total + (np.rint(d.astype(np.float64).T @ right).astype(np.int64).astype(object) << (8 * k))
    return total, bool(np.any(q != 0))


def _evaluate(m, plan, origin):
    f, fw = plan.frac_bits, plan.frac_bits_w
    if plan.rounding:
        qT = np.floor(m["T3"] * 2.0 ** f + dither_u01(origin, M)).astype(np.int64)
    else:
        qT = np.rint(m["T3"] * 2.0 ** f).astype(np.int64)
    qR = np.rint(m["R3"] * 2.0 ** fw).astype(np.int64)
    main, o1 = _digit_gram(qT, plan.digits, m["code"], m["B"])
    corr, o2 = _digit_gram(qR, plan.digits_w, m["code"], m["mis"])
    assert not (o1 or o2), "digit overflow"
    r4 = np.concatenate([m["R3"], np.zeros((M, 1))], axis=1)
    vec = np.take_along_axis(r4, m["code"], axis=1).sum(axis=0)
    C = main.astype(np.float64) / 2.0 ** f + corr.astype(np.float64) / 2.0 ** fw - vec[:, None]
    return float(np.max(np.abs(C - m["C"]))) / m["stats"]["scale"]


def test_round_to_nearest_meets_its_worst_case_bound(model):
    p = _plan(model["stats"])
    plan_format(0, p, "nearest", N_FORMAT)
    err = _evaluate(model, p, 0)
    bound = (2.0 ** -(p.frac_bits + 1) * p.err_weight + 2.0 ** -(p.frac_bits_w + 1) * p.max_missing) / p.scale
    assert p.rounding == 0 and err <= bound <= 1e-10, (err, bound)


def test_randomised_rounding_meets_the_hoeffding_bound_over_draws(model):
    p = _plan(model["stats"])
    n_near = plan_format(0, _plan(model["stats"]), "nearest", N_FORMAT)
    assert plan_format(0, p, "auto", N_FORMAT) == n_near - 1 and p.rounding == 1       # 200 000 SNPs: one digit of T saved
    small = _plan(model["stats"])
    assert plan_format(0, small, "auto", N) == n_near and small.rounding == 0    # 64^2 x 200 000 < 2^36: not worth it
    pairs = 0.5 * N_FORMAT * (N_FORMAT + 1)
    bound = (2.0 ** -p.frac_bits * math.sqrt(0.5 * p.err_weight2 * math.log(2 * pairs / 1e-12))
             + 2.0 ** -(p.frac_bits_w + 1) * p.max_missing) / p.scale
    assert bound <= 0.9e-10 * 1.0000001
    errs = [_evaluate(model, p, origin) for origin in (0, 200000, 99999989, 2 ** 41 + 5)]
    assert max(errs) <= bound, (errs, bound)
    assert len(set(errs)) == len(errs)          # different origins are different draws
    # unbiased: the mean error over the draws shrinks (a biased rounding rule would not)
    assert max(errs) < 1e-10
