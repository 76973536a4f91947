"""Host-side mirror of .InitFile2 / family.id handling (R/Internal.R:166-484,
R/IBD.R:348-372): argument checking and selection logic that runs before any
device call."""
import numpy as np
import pytest

import snprelate_b200 as S
from snprelate_b200 import api


def _gds(nsnp=12, nsamp=6):
    g = (np.arange(nsnp * nsamp).reshape(nsnp, nsamp) % 3).astype(np.uint8)
    return S.GenotypeData(g, sample_id=[f"s{i}" for i in range(nsamp)],
                          snp_id=np.arange(1, nsnp + 1), chromosome=np.r_[np.ones(nsnp - 2), 23, 23])


def test_unknown_sample_id_message():
    with pytest.raises(S.SNPRelError, match="Some of sample.id do not exist!"):
        S.snpgdsIBS(_gds(), sample_id=["s0", "nope"])


def test_unknown_snp_id_message():
    with pytest.raises(S.SNPRelError, match="Some of snp.id do not exist!"):
        S.snpgdsGRM(_gds(), snp_id=[1, 99])


def test_bad_method_and_type():
    with pytest.raises(S.SNPRelError):
        S.snpgdsGRM(_gds(), method="nope")
    with pytest.raises(S.SNPRelError, match="Invalid 'type'"):
        S.snpgdsIBDKING(_gds(), type="KING-x")
    with pytest.raises(S.SNPRelError, match="num.thread"):
        S.snpgdsIBS(_gds(), num_thread=0)


def test_packed_block_decoding():
    rng = np.random.default_rng(0)
    g = rng.integers(0, 4, size=(9, 11)).astype(np.uint8)
    nb = (11 + 3) // 4
    pad = np.full((9, nb * 4), 3, dtype=np.uint8)
    pad[:, :11] = g
    p = pad.reshape(9, nb, 4)
    packed = (p[:, :, 0] | (p[:, :, 1] << 2) | (p[:, :, 2] << 4) | (p[:, :, 3] << 6)).astype(np.uint8)
    gd = S.GenotypeData(packed, packed_2bit=True, n_samp=11)
    mask = np.array([True, False] * 5 + [True])
    assert np.array_equal(gd.block_u8(np.array([0, 3, 8]), mask), g[[0, 3, 8]][:, mask])


def test_family_codes():
    ws = dict(n_samp=5, sample_id=np.array(["a", "b", "c", "d", "e"]))
    fam = api._family_codes(["F2", "", "F1", None, "F2"], None, None, ws)
    assert fam[0] == fam[4] and fam[0] != fam[2]
    assert fam[1] == api.NA_INT and fam[3] == api.NA_INT
    with pytest.raises(S.SNPRelError, match="length"):
        api._family_codes(["x"], None, None, ws)


def test_family_codes_follow_the_reference_reindexing():
    """R/IBD.R:356-357: family.id <- family.id[match(sample.id, ws$sample.id)].  With a sample.id that is
    a 3-cycle of the file order the reference picks family.id at the WORKSPACE position of each requested
    sample (not the inverse permutation); the mirror must pick the same entries."""
    ws = dict(n_samp=3, sample_id=np.array(["a", "b", "c"]))           # workspace = file order
    sample_id = ["b", "c", "a"]                                        # a 3-cycle, not an involution
    family = ["F1", "F2", "F3"]
    # R: match(c("b","c","a"), c("a","b","c")) = (2, 3, 1)  ->  family.id[c(2, 3, 1)] = F2, F3, F1
    fam = api._family_codes(family, None, sample_id, ws)
    lv = {"F1": 1, "F2": 2, "F3": 3}
    assert fam.tolist() == [lv["F2"], lv["F3"], lv["F1"]]


def test_loading_and_correlation_argument_checks():
    """snpgdsPCACorr / snpgdsPCASNPLoading / snpgdsPCASampLoading refuse malformed inputs before any
    device call (R/PCA.R:99-303)."""
    gds = _gds()
    with pytest.raises(S.SNPRelError, match="pcaobj"):
        S.snpgdsPCASNPLoading({"eigenvect": np.zeros((6, 2))}, gds)
    with pytest.raises(S.SNPRelError, match="pcaobj"):
        S.snpgdsPCACorr(np.zeros((6, 2)), gds)
    with pytest.raises(S.SNPRelError, match="sample id"):
        S.snpgdsPCACorr((None, np.zeros((6, 2))), gds)
    pca = {"class": "snpgdsPCAClass", "sample.id": [f"s{i}" for i in range(6)], "snp.id": np.arange(1, 13),
           "eigenval": None, "eigenvect": None}
    with pytest.raises(S.SNPRelError, match="eigenvalues"):
        S.snpgdsPCASNPLoading(pca, gds)
    with pytest.raises(S.SNPRelError, match="outgds"):
        S.snpgdsPCACorr(dict(pca, eigenvect=np.zeros((6, 2))), gds, outgds="x.gds")
    with pytest.raises(S.SNPRelError, match="loadobj"):
        S.snpgdsPCASampLoading({"class": "snpgdsPCAClass"}, gds)
    # unknown sample ids are reported by the shared selection code before the device is touched
    with pytest.raises(S.SNPRelError, match="Some of sample.id do not exist!"):
        S.snpgdsPCACorr((["s0", "zz"], np.zeros((2, 1))), gds)


def test_window_owner_is_a_permutation_per_round():
    from snprelate_b200._lib import window_owner
    for world in (1, 2, 5, 8):
        for base in range(0, 6 * world, 2 * world):
            fwd = [window_owner(base + i, world) for i in range(world)]
            back = [window_owner(base + world + i, world) for i in range(world)]
            assert fwd == list(range(world)) and back == list(range(world))[::-1]


def _plan(**kw):
    from snprelate_b200._lib import Plan
    p = Plan()
    p.frac_bits = p.frac_bits_w = p.frac_bits_d = -1
    for k, v in kw.items():
        setattr(p, k, v)
    return p


# plan statistics of BASELINE config 2 (10 000 samples x 1 000 000 SNPs, 0.5 % missing) as the device measures them
CONFIG2 = dict(max_abs=0.807, max_abs_w=0.085, err_weight=1.264e7, err_weight2=3.15e8, scale=1.99e6, sum_bound=1.651e7,
               diag_bound=2.365e6, sum_rest=2.804e4, total_missing=50000343, max_missing=5300, n_snp=1000000)


def test_format_choice_at_config2_statistics():
    """snprel_plan_format (host only): round-to-nearest needs T5 + R3 = 8 tensor passes for the 1e-10 bound,
    randomised rounding T4 + R3 = 7, 'auto' takes the cheaper one and reports which rounding it used."""
    from snprelate_b200._lib import plan_format
    p = _plan(**CONFIG2)
    assert plan_format(0, p, "nearest", 10000) == 8 and (p.digits, p.digits_w, p.digits_d, p.rounding) == (5, 3, 0, 0)
    near_f = p.frac_bits
    p = _plan(**CONFIG2)
    assert plan_format(0, p, "random", 10000) == 7 and (p.digits, p.digits_w, p.rounding) == (4, 3, 1)
    assert p.frac_bits < near_f
    p = _plan(**CONFIG2)
    assert plan_format(0, p, "auto", 10000) == 7 and p.rounding == 1
    # GCTA: one more exact pass for the missing-pair denominators
    p = _plan(**dict(CONFIG2, scale=2e6))
    assert plan_format(1, p, "auto", 10000) == 8 and (p.digits, p.digits_w, p.digits_d) == (4, 3, 1)
    # the Hoeffding bound the library promises: 2^-f sqrt(1/2 sum B^2 ln(2 pairs / 1e-12)) + the R term <= 0.9e-10 scale
    import math
    p = _plan(**CONFIG2)
    plan_format(0, p, "random", 10000)
    t = 2.0 ** -p.frac_bits * math.sqrt(0.5 * p.err_weight2 * math.log(2 * (0.5 * 1e4 * (1e4 + 1)) / 1e-12))
    r = 2.0 ** -(p.frac_bits_w + 1) * p.max_missing
    assert t + r <= 0.9e-10 * p.scale
    # without the measured sum of squares the bound falls back to sum B^2 <= 127 sum |B|
    p = _plan(**dict(CONFIG2, err_weight2=0.0))
    assert plan_format(0, p, "auto", 10000) == 7


def test_format_choice_keeps_round_to_nearest_where_nothing_is_saved():
    from snprelate_b200._lib import plan_format
    for frac in (0.009, 0.02, 0.05):          # 9 000 ... 50 000 SNPs: the sqrt(M) advantage is below one digit
        st = {k: (v * frac if k not in ("max_abs", "max_abs_w") else v) for k, v in CONFIG2.items()}
        st.update(total_missing=int(st["total_missing"]), max_missing=int(st["max_missing"]) + 20, n_snp=int(st["n_snp"]), diag_bound=0.0)
        a, b = _plan(**st), _plan(**st)
        assert plan_format(0, a, "nearest", 700) == plan_format(0, b, "auto", 700)
        assert b.rounding == 0 and (a.frac_bits, a.digits) == (b.frac_bits, b.digits)
    # a caller-fixed format brings its rounding along
    p = _plan(**CONFIG2)
    p.frac_bits, p.frac_bits_w, p.rounding = 31, 26, 1
    assert plan_format(0, p, "nearest", 10000) == 7 and p.rounding == 1
    # KING-homo has no real row table: never randomised
    p = _plan(**dict(CONFIG2, err_weight=0.0, err_weight2=0.0))
    plan_format(13, p, "random", 10000)
    assert p.rounding == 0


def test_plan_format_rejects_bad_arguments():
    import snprelate_b200 as S
    from snprelate_b200._lib import plan_format
    with pytest.raises(S.SNPRelError):
        plan_format(10, _plan(**CONFIG2), "auto", 100)       # IBS is not a covariance estimator
    with pytest.raises(S.SNPRelError):
        plan_format(0, _plan(**CONFIG2), 7, 100)


def test_rounding_environment_variable_is_validated(monkeypatch):
    """SNPREL_ROUNDING gives the initial rounding mode of every context; a bad value fails snprel_create
    before any device is touched."""
    import snprelate_b200 as S
    monkeypatch.setenv("SNPREL_ROUNDING", "stochastic")
    with pytest.raises(S.SNPRelError, match="SNPREL_ROUNDING"):
        S.Context(0)
