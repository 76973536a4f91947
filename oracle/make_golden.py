"""ORACLE / TEST INFRASTRUCTURE ONLY.

Decode the reference's fixture and golden vectors (read from /root/reference,
build container only) into tests/golden/*.npz.  Re-run with

    python oracle/make_golden.py [/root/reference]

The committed .npz files are what the tests use; /root/reference is never read
at test / smoke / bench time.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.fixture import load_hapmap_gds, load_rdata  # noqa: E402


def pack2(geno):
    """uint8 codes [nsnp, nsamp] -> row-padded 2-bit, 4 per byte LSB first."""
    nsnp, nsamp = geno.shape
    nb = (nsamp + 3) // 4
    g = np.zeros((nsnp, nb * 4), dtype=np.uint8)
    g[:, :nsamp] = np.minimum(geno, 3)
    g[:, nsamp:] = 3
    g = g.reshape(nsnp, nb, 4)
    return (g[:, :, 0] | (g[:, :, 1] << 2) | (g[:, :, 2] << 4) | (g[:, :, 3] << 6)).astype(np.uint8)


def main(ref="/root/reference"):
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    os.makedirs(out, exist_ok=True)
    h = load_hapmap_gds(os.path.join(ref, "inst/extdata/hapmap_geno.gds"))
    np.savez_compressed(
        os.path.join(out, "hapmap_geno.npz"),
        geno2b=pack2(h["geno"]), nsamp=np.int64(h["geno"].shape[1]),
        sample_id=np.array(h["sample_id"]), snp_id=h["snp_id"],
        chromosome=h["chromosome"], position=h["position"])
    v = os.path.join(ref, "inst/unitTests/valid")
    ibs = load_rdata(os.path.join(v, "Validate.IBS.RData"))["ibs"]
    pca = load_rdata(os.path.join(v, "Validate.PCA.RData"))[".rv"]
    king = load_rdata(os.path.join(v, "Validate.KING.RData"))[".king"]
    beta = load_rdata(os.path.join(v, "Validate.Beta.RData"))[".beta"]
    eigmix = load_rdata(os.path.join(v, "Validate.EIGMIX.RData"))[".eigmix"]
    mom = load_rdata(os.path.join(v, "Validate.MoM.RData"))["ibd"]
    np.savez_compressed(
        os.path.join(out, "reference_goldens.npz"),
        ibs=ibs["ibs"], ibs_snp_id=ibs["snp.id"],
        pca_genmat=pca["genmat"], pca_samploading=pca["samploading"],
        pca_snploading=pca["snploading"], pca_corr=pca["corr"],
        king_snp_id=king[0]["snp.id"],
        king_robust_ibs0=king[0]["IBS0"], king_robust_kinship=king[0]["kinship"],
        king_homo_k0=king[1]["k0"], king_homo_k1=king[1]["k1"],
        beta=beta["beta"], beta_snp_id=beta["snp.id"],
        eigmix_ibd=eigmix,
        mom_k0=mom["k0"], mom_k1=mom["k1"], mom_afreq=mom["afreq"])
    print("wrote", out)


if __name__ == "__main__":
    main(*sys.argv[1:])
