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
