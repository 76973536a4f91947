// R-side binding of libsnprel_b200.so: drop-in bodies for the reference's .Call entry points
// of the relatedness path.  The build image has neither R nor gdsfmt; oracle/Makefile compiles this
// file against the reference's own workspace layer (dGenGWAS.h / dGenGWAS.cpp, unmodified) and a
// stand-in for the R / gdsfmt headers (oracle/ref_shim) into oracle/_ref/libsnprelate_b200_rshim.so,
// which tests/test_gpu_rshim.py drives on the GPU.  It replaces the bodies of
//   gnrGRM              src/genPCA.cpp:1614-1717      gnrIBSAve          src/genIBS.cpp:441-497
//   gnrPCA (both algs)  src/genPCA.cpp:1355-1452      gnrIBSNum          src/genIBS.cpp:500-550
//   gnrEigMix           src/genEIGMIX.cpp:656-735     gnrIBD_KING_Robust src/genKING.cpp:576-679
//   gnrGRM_avg_val      src/genPCA.cpp:1608           gnrIBD_KING_Homo   src/genKING.cpp:493-570
//   gnrIBD_PLINK        src/genIBS.cpp:558-639        gnrIBD_Beta        src/genBeta.cpp:361-460
//   gnrPCACorr          src/genPCA.cpp:1456-1485      gnrPCASNPLoading   src/genPCA.cpp:1489-1540
//   gnrPCASampLoading   src/genPCA.cpp:1542-1563      gnrEigMixSNPLoading / gnrEigMixSampLoading
//                                                     src/genEIGMIX.cpp:739-803
// inside SNPRelate.so, keeping the workspace layer (gnrSetGenoSpace / gnrSelSNP_Base,
// src/SNPRelate.cpp:76-214) and the R code untouched: every routine streams the SELECTED
// genotypes through CdBaseWorkSpace::snpRead (src/dGenGWAS.h:94) into the device workspace and
// returns R objects of exactly the reference's shape.
#if defined(SNPREL_BUILD_R_SHIM)

#include <cstdlib>
#include <cstring>
#include <vector>

#include "dGenGWAS.h"          // the reference's workspace (GWAS::MCWorkingGeno, SEXP helpers)
#include "snprel_b200.h"

using namespace GWAS;

namespace {

// One device context per R process, created at the first estimator call and re-used: it is the
// counterpart of the reference's process-global GWAS::MCWorkingGeno (src/dGenGWAS.cpp:2000) and
// saves a CUDA stream / event / buffer set-up per .Call.  Every routine re-loads the SELECTED
// genotypes, because gnrSetGenoSpace / gnrSelSNP_Base may have changed the selection in between.
struct Ctx {
    snprel_ctx *h = nullptr;
    Ctx() {
        static snprel_ctx *shared = nullptr;
        if (!shared && snprel_create(&shared, 0) != 0) {
            shared = nullptr;
            throw ErrCoreArray("%s", snprel_last_error(nullptr));
        }
        h = shared;
    }
    void ck(int rc) {
        if (rc != 0) throw ErrCoreArray("%s", snprel_last_error(h));
    }
};

// CGenoReadBySNP equivalent: selected SNPs x selected samples -> device, block by block
void load_workspace(Ctx &c) {
    CdBaseWorkSpace &sp = MCWorkingGeno.Space();
    const int n = sp.SampleNum(), m = sp.SNPNum();
    c.ck(snprel_geno_begin(c.h, n, m));
    const int block = std::max(1, (64 << 20) / std::max(n, 1));
    std::vector<C_UInt8> buf((size_t)block * n);
    for (int st = 0; st < m; st += block) {
        int cnt = std::min(block, m - st);
        sp.snpRead(st, cnt, &buf[0], RDim_Sample_X_SNP);        // u8 [cnt][n], >2 = missing
        c.ck(snprel_geno_push_u8(c.h, &buf[0], cnt));
    }
}

// Several GPUs from one R process: SNPREL_DEVICES="0,1,2,3" (or "all") makes the N x N estimators
// shard the selected SNPs over those devices (snprel_multi_*, csrc/multi.cu: per-device accumulation,
// peer-memory reduction over NVLink onto the first device), and the routine then finishes on that
// device's context exactly as in the single-device case.  Routines without a sharded form (PLINK
// MoM, KING-homo, loadings, randomized PCA) keep using one device.
snprel_multi *multi_handle() {
    static snprel_multi *shared = nullptr;
    static bool tried = false;
    if (tried) return shared;
    tried = true;
    const char *env = getenv("SNPREL_DEVICES");
    if (!env || !*env) return nullptr;
    std::vector<int> dev;
    if (!strcmp(env, "all")) {
        const int n = snprel_device_count();
        for (int i = 0; i < n; i++) dev.push_back(i);
    } else {
        for (const char *p = env; *p;) {
            char *end = nullptr;
            long v = strtol(p, &end, 10);
            if (end == p) throw ErrCoreArray("SNPREL_DEVICES: expected a comma separated list of device indices or \"all\"");
            dev.push_back((int)v);
            p = (*end == ',') ? end + 1 : end;
        }
    }
    if (dev.size() < 2) return nullptr;
    if (snprel_multi_create(dev.data(), (int)dev.size(), &shared) != 0) {
        shared = nullptr;
        throw ErrCoreArray("%s", snprel_multi_last_error(nullptr));
    }
    return shared;
}

// load the selected genotypes for estimator `est`; with several devices also accumulate + reduce, after
// which c.h is the context that holds the global accumulators
void load_for(Ctx &c, int est, bool bayesian = false) {
    snprel_multi *m = multi_handle();
    if (!m) {
        load_workspace(c);
        return;
    }
    CdBaseWorkSpace &sp = MCWorkingGeno.Space();
    const int n = sp.SampleNum(), msnp = sp.SNPNum();
    auto mck = [&](int rc) { if (rc != 0) throw ErrCoreArray("%s", snprel_multi_last_error(m)); };
    mck(snprel_multi_geno_begin(m, n, msnp));
    const int block = std::max(1, (64 << 20) / std::max(n, 1));
    std::vector<C_UInt8> buf((size_t)block * n);
    for (int st = 0; st < msnp; st += block) {
        int cnt = std::min(block, msnp - st);
        sp.snpRead(st, cnt, &buf[0], RDim_Sample_X_SNP);
        mck(snprel_multi_geno_push_u8(m, &buf[0], cnt));
    }
    mck(snprel_multi_accumulate(m, est, bayesian ? 1 : 0, 0));
    c.h = snprel_multi_ctx(m, 0);
}

// allele frequencies of ALL selected SNPs in multi-device mode (each context holds its own range)
void gather_afreq(double *af) {
    snprel_multi *m = multi_handle();
    int64_t off = 0;
    for (int i = 0; i < snprel_multi_device_count(m); i++) {
        snprel_ctx *h = snprel_multi_ctx(m, i);
        int64_t ns = 0, ms = 0;
        if (snprel_geno_dim(h, &ns, &ms) != 0 || ms <= 0) continue;
        if (snprel_snp_ratefreq(h, af + off, NULL, NULL) != 0) throw ErrCoreArray("%s", snprel_last_error(h));
        for (int64_t l = 0; l < ms; l++)
            if (!(af[off + l] == af[off + l])) af[off + l] = 0.0;      // no valid genotype: avg_geno = 0 (src/genEIGMIX.cpp:116-118)
        off += ms;
    }
}

SEXP sym_result(size_t n, bool packed) {
    return packed ? NEW_NUMERIC(n * (n + 1) / 2) : Rf_allocMatrix(REALSXP, n, n);
}

double g_avg_val = 0;

}  // namespace

extern "C" {

COREARRAY_DLL_EXPORT SEXP gnrGRM_avg_val() { return Rf_ScalarReal(g_avg_val); }

COREARRAY_DLL_EXPORT SEXP gnrGRM(SEXP NumThread, SEXP Method, SEXP GDS, SEXP useMatrix, SEXP Verbose) {
    const char *mt = CHAR(STRING_ELT(Method, 0));
    COREARRAY_TRY
        // R/IBD.R:570-594: with out.fn the matrix goes to the "grm" node of a SNPRELATE_OUTPUT file
        PdGDSObj gdsn = Rf_isNull(GDS) ? NULL : GDS_Node_Path(GDS_R_SEXP2FileRoot(GDS), "grm", TRUE);   // src/genPCA.cpp:1623-1625
        int method = !strcmp(mt, "Eigenstrat") ? SNPREL_GRM_EIGENSTRAT : !strcmp(mt, "GCTA") ? SNPREL_GRM_GCTA
                   : !strcmp(mt, "Corr") ? SNPREL_GRM_CORR : !strcmp(mt, "EIGMIX") ? SNPREL_GRM_EIGMIX
                   : !strcmp(mt, "IndivBeta") ? SNPREL_GRM_INDIVBETA : -1;
        if (method < 0) throw ErrCoreArray("Invalid 'method'!");
        Ctx c;
        load_for(c, method == SNPREL_GRM_INDIVBETA ? SNPREL_EST_BETA : method);
        const size_t n = MCWorkingGeno.Space().SampleNum();
        if (gdsn) {
            // grm_save_to_gds (src/genPCA.cpp:1571-1584): n appends of one full row each, in row order --
            // the node is extendable along its last dimension and usually compressed, so the stream is
            // strictly sequential.  Like the reference (CdMatTri) the upper triangle is held on the host.
            const bool pk = method != SNPREL_GRM_CORR;
            std::vector<double> tri(pk ? n * (n + 1) / 2 : n * n), row(n);
            c.ck(snprel_grm(c.h, method, tri.data(), pk, &g_avg_val));
            for (size_t i = 0; i < n; i++) {
                if (pk) {
                    for (size_t j = 0; j < i; j++) row[j] = tri[j * (2 * n - j - 1) / 2 + i];      // (j, i), j < i
                    memcpy(&row[i], &tri[i * (2 * n - i - 1) / 2 + i], (n - i) * sizeof(double));
                } else {
                    memcpy(&row[0], &tri[i * n], n * sizeof(double));
                }
                GDS_Array_AppendData(gdsn, n, &row[0], svFloat64);
            }
            return R_NilValue;
        }
        const bool packed = (Rf_asLogical(useMatrix) == TRUE) && method != SNPREL_GRM_CORR;
        rv_ans = PROTECT(sym_result(n, packed));
        c.ck(snprel_grm(c.h, method, REAL(rv_ans), packed, &g_avg_val));
        UNPROTECT(1);
    COREARRAY_CATCH
}

COREARRAY_DLL_EXPORT SEXP gnrPCA(SEXP EigenCnt, SEXP Algorithm, SEXP NumThread, SEXP ParamList, SEXP Verbose) {
    COREARRAY_TRY
        const char *alg = CHAR(STRING_ELT(Algorithm, 0));
        if (strcmp(alg, "exact") != 0 && strcmp(alg, "randomized") != 0) throw "Invalid 'algorithm'.";
        Ctx c;
        const int n = MCWorkingGeno.Space().SampleNum();
        if (strcmp(alg, "randomized") == 0) {
            load_workspace(c);
            // src/genPCA.cpp:1436-1442, :781-793: list(sigma[n], V^T [hsize x n], 2 TraceXTX)
            const int aux_dim = Rf_asInteger(RGetListElement(ParamList, "aux.dim"));
            const int iter_num = Rf_asInteger(RGetListElement(ParamList, "iter.num"));
            const size_t hsize = (size_t)aux_dim * (size_t)(iter_num + 1);
            std::vector<double> vt(hsize * (size_t)n);
            PROTECT(rv_ans = NEW_LIST(3));
            SEXP d = PROTECT(NEW_NUMERIC(n)); SET_ELEMENT(rv_ans, 0, d); UNPROTECT(1);
            SEXP h = PROTECT(Rf_allocMatrix(REALSXP, (int)hsize, n)); SET_ELEMENT(rv_ans, 1, h); UNPROTECT(1);
            double tr2 = 0;
            c.ck(snprel_pca_randomized(c.h, REAL(RGetListElement(ParamList, "aux.mat")), aux_dim, iter_num, REAL(d),
                                       vt.data(), &tr2));
            double *ph = REAL(h);      // R matrix hsize x n, column major
            for (size_t r = 0; r < hsize; r++)
                for (size_t i = 0; i < (size_t)n; i++) ph[r + i * hsize] = vt[r * (size_t)n + i];
            SET_ELEMENT(rv_ans, 2, Rf_ScalarReal(tr2));
            UNPROTECT(1);
            return rv_ans;
        }
        int nEig = Rf_asInteger(EigenCnt);
        if (nEig < 0) throw ErrCoreArray("Invalid 'eigen.cnt'.");
        if (nEig > n) nEig = n;
        const bool bayes = Rf_asLogical(RGetListElement(ParamList, "bayesian")) == TRUE;
        const bool need = Rf_asLogical(RGetListElement(ParamList, "need.genmat")) == TRUE;
        const bool only = Rf_asLogical(RGetListElement(ParamList, "genmat.only")) == TRUE;
        load_for(c, SNPREL_GRM_EIGENSTRAT, bayes);
        PROTECT(rv_ans = NEW_LIST(5));
        SEXP genmat = R_NilValue, eval = R_NilValue, evec = R_NilValue;
        if (need) { genmat = PROTECT(Rf_allocMatrix(REALSXP, n, n)); SET_ELEMENT(rv_ans, 1, genmat); UNPROTECT(1); }
        if (!only) {
            eval = PROTECT(NEW_NUMERIC(n)); SET_ELEMENT(rv_ans, 2, eval); UNPROTECT(1);
            evec = PROTECT(Rf_allocMatrix(REALSXP, n, nEig)); SET_ELEMENT(rv_ans, 3, evec); UNPROTECT(1);
        }
        double tx = 0, tv = 0;
        c.ck(snprel_pca(c.h, nEig, bayes, need ? REAL(genmat) : NULL, &tx, &tv,
                        only ? NULL : REAL(eval), only ? NULL : REAL(evec)));
        SET_ELEMENT(rv_ans, 0, Rf_ScalarReal(tx));
        SET_ELEMENT(rv_ans, 4, Rf_ScalarReal(tv));
        UNPROTECT(1);
    COREARRAY_CATCH
}

COREARRAY_DLL_EXPORT SEXP gnrEigMix(SEXP EigenCnt, SEXP NumThread, SEXP ParamList, SEXP Verbose) {
    int diag_adj = Rf_asLogical(RGetListElement(ParamList, "diagadj"));
    if (diag_adj == NA_LOGICAL) Rf_error("'diagadj' must be TRUE or FALSE.");
    int need_ibd = Rf_asLogical(RGetListElement(ParamList, "ibdmat"));
    if (need_ibd == NA_LOGICAL) Rf_error("'ibdmat' must be TRUE or FALSE.");
    COREARRAY_TRY
        Ctx c;
        load_for(c, SNPREL_GRM_EIGMIX);
        const int n = MCWorkingGeno.Space().SampleNum();
        int nEig = Rf_asInteger(EigenCnt);
        if (nEig < 0 || nEig > n) nEig = n;
        PROTECT(rv_ans = NEW_LIST(4));
        SEXP af = PROTECT(NEW_NUMERIC(MCWorkingGeno.Space().SNPNum())); SET_ELEMENT(rv_ans, 2, af); UNPROTECT(1);
        SEXP ibd = R_NilValue, eval = R_NilValue, evec = R_NilValue;
        if (need_ibd) { ibd = PROTECT(Rf_allocMatrix(REALSXP, n, n)); SET_ELEMENT(rv_ans, 3, ibd); UNPROTECT(1); }
        if (nEig > 0) {
            eval = PROTECT(NEW_NUMERIC(n)); SET_ELEMENT(rv_ans, 0, eval); UNPROTECT(1);
            evec = PROTECT(Rf_allocMatrix(REALSXP, n, nEig)); SET_ELEMENT(rv_ans, 1, evec); UNPROTECT(1);
        }
        c.ck(snprel_eigmix(c.h, nEig, diag_adj == TRUE, need_ibd ? REAL(ibd) : NULL, multi_handle() ? NULL : REAL(af),
                           nEig > 0 ? REAL(eval) : NULL, nEig > 0 ? REAL(evec) : NULL));
        if (multi_handle()) gather_afreq(REAL(af));
        UNPROTECT(1);
    COREARRAY_CATCH
}

COREARRAY_DLL_EXPORT SEXP gnrIBSAve(SEXP NumThread, SEXP useMatrix, SEXP Verbose) {
    COREARRAY_TRY
        Ctx c;
        load_for(c, SNPREL_EST_IBS);
        const size_t n = MCWorkingGeno.Space().SampleNum();
        const bool packed = Rf_asLogical(useMatrix) == TRUE;
        rv_ans = PROTECT(sym_result(n, packed));
        c.ck(snprel_ibs_ave(c.h, REAL(rv_ans), packed));
        UNPROTECT(1);
    COREARRAY_CATCH
}

COREARRAY_DLL_EXPORT SEXP gnrIBD_PLINK(SEXP NumThread, SEXP AlleleFreq, SEXP UseSpecificAFreq,
                                       SEXP KinshipConstrict, SEXP useMatrix, SEXP Verbose) {
    COREARRAY_TRY
        Ctx c;
        load_workspace(c);
        const size_t n = MCWorkingGeno.Space().SampleNum();
        const bool packed = Rf_asLogical(useMatrix) == TRUE;
        SEXP k0 = PROTECT(sym_result(n, packed));
        SEXP k1 = PROTECT(sym_result(n, packed));
        SEXP afreq = PROTECT(NEW_NUMERIC(MCWorkingGeno.Space().SNPNum()));
        c.ck(snprel_ibd_mom(c.h, Rf_asLogical(UseSpecificAFreq) == TRUE ? REAL(AlleleFreq) : NULL,
                            Rf_asLogical(KinshipConstrict) == TRUE, packed, REAL(k0), REAL(k1), REAL(afreq)));
        PROTECT(rv_ans = NEW_LIST(3));
        SET_ELEMENT(rv_ans, 0, k0);
        SET_ELEMENT(rv_ans, 1, k1);
        SET_ELEMENT(rv_ans, 2, afreq);
        UNPROTECT(4);
    COREARRAY_CATCH
}

COREARRAY_DLL_EXPORT SEXP gnrIBSNum(SEXP NumThread, SEXP Verbose) {
    COREARRAY_TRY
        Ctx c;
        load_for(c, SNPREL_EST_IBS);
        const int n = MCWorkingGeno.Space().SampleNum();
        PROTECT(rv_ans = NEW_LIST(3));
        SEXP m[3];
        for (int k = 0; k < 3; k++) { m[k] = PROTECT(Rf_allocMatrix(INTSXP, n, n)); SET_ELEMENT(rv_ans, k, m[k]); UNPROTECT(1); }
        c.ck(snprel_ibs_num(c.h, INTEGER(m[0]), INTEGER(m[1]), INTEGER(m[2])));
        UNPROTECT(1);
    COREARRAY_CATCH
}

COREARRAY_DLL_EXPORT SEXP gnrIBD_KING_Robust(SEXP FamilyID, SEXP NumThread, SEXP useMatrix, SEXP Verbose) {
    COREARRAY_TRY
        Ctx c;
        load_for(c, SNPREL_EST_KING_ROBUST);
        const size_t n = MCWorkingGeno.Space().SampleNum();
        const bool packed = Rf_asLogical(useMatrix) == TRUE;
        PROTECT(rv_ans = NEW_LIST(2));
        SEXP a = PROTECT(sym_result(n, packed)); SET_ELEMENT(rv_ans, 0, a); UNPROTECT(1);
        SEXP b = PROTECT(sym_result(n, packed)); SET_ELEMENT(rv_ans, 1, b); UNPROTECT(1);
        // NA_INTEGER == INT32_MIN == SNPREL_NA_INT: the family vector passes through unchanged
        c.ck(snprel_king_robust(c.h, INTEGER(FamilyID), REAL(a), REAL(b), packed));
        UNPROTECT(1);
    COREARRAY_CATCH
}

COREARRAY_DLL_EXPORT SEXP gnrIBD_KING_Homo(SEXP NumThread, SEXP useMatrix, SEXP Verbose) {
    COREARRAY_TRY
        Ctx c;
        load_workspace(c);
        const size_t n = MCWorkingGeno.Space().SampleNum();
        const bool packed = Rf_asLogical(useMatrix) == TRUE;
        PROTECT(rv_ans = NEW_LIST(2));
        SEXP a = PROTECT(sym_result(n, packed)); SET_ELEMENT(rv_ans, 0, a); UNPROTECT(1);
        SEXP b = PROTECT(sym_result(n, packed)); SET_ELEMENT(rv_ans, 1, b); UNPROTECT(1);
        c.ck(snprel_king_homo(c.h, REAL(a), REAL(b), packed));
        UNPROTECT(1);
    COREARRAY_CATCH
}

COREARRAY_DLL_EXPORT SEXP gnrIBD_Beta(SEXP Inbreeding, SEXP NumThread, SEXP useMatrix, SEXP Verbose) {
    int inbreeding = Rf_asLogical(Inbreeding);
    if (inbreeding == NA_LOGICAL) Rf_error("'inbreeding' must be TRUE or FALSE.");
    COREARRAY_TRY
        Ctx c;
        load_for(c, SNPREL_EST_BETA);
        const size_t n = MCWorkingGeno.Space().SampleNum();
        const bool packed = Rf_asLogical(useMatrix) == TRUE;
        rv_ans = PROTECT(sym_result(n, packed));
        c.ck(snprel_indiv_beta(c.h, inbreeding == TRUE, REAL(rv_ans), packed, &g_avg_val));
        UNPROTECT(1);
    COREARRAY_CATCH
}

// ---- loadings / projection / SNP-PC correlation (csrc/project.cu) ----

COREARRAY_DLL_EXPORT SEXP gnrPCACorr(SEXP LenEig, SEXP EigenVect, SEXP NumThread, SEXP GDSNode, SEXP Verbose) {
    const int nEig = Rf_asInteger(LenEig);
    COREARRAY_TRY
        if (!Rf_isNull(GDSNode)) throw ErrCoreArray("GDS output is written by the host wrapper (INTEGRATION.md).");
        Ctx c;
        load_workspace(c);
        rv_ans = PROTECT(Rf_allocMatrix(REALSXP, nEig, MCWorkingGeno.Space().SNPNum()));
        c.ck(snprel_pca_corr(c.h, nEig, REAL(EigenVect), REAL(rv_ans)));
        UNPROTECT(1);
    COREARRAY_CATCH
}

COREARRAY_DLL_EXPORT SEXP gnrPCASNPLoading(SEXP EigenVal, SEXP EigenVect, SEXP TraceXTX, SEXP NumThread,
                                           SEXP Bayesian, SEXP Verbose) {
    const int LenEig = INTEGER(GET_DIM(EigenVect))[1];
    COREARRAY_TRY
        Ctx c;
        load_workspace(c);
        const size_t m = MCWorkingGeno.Space().SNPNum();
        PROTECT(rv_ans = NEW_LIST(3));
        SEXP loading = PROTECT(Rf_allocMatrix(REALSXP, LenEig, m)); SET_ELEMENT(rv_ans, 0, loading); UNPROTECT(1);
        SEXP afreq = PROTECT(NEW_NUMERIC(m)); SET_ELEMENT(rv_ans, 1, afreq); UNPROTECT(1);
        SEXP scale = PROTECT(NEW_NUMERIC(m)); SET_ELEMENT(rv_ans, 2, scale); UNPROTECT(1);
        c.ck(snprel_pca_snp_loading(c.h, LenEig, REAL(EigenVal), REAL(EigenVect), Rf_asReal(TraceXTX),
                                    Rf_asLogical(Bayesian) == TRUE, REAL(loading), REAL(afreq), REAL(scale)));
        UNPROTECT(1);
    COREARRAY_CATCH
}

COREARRAY_DLL_EXPORT SEXP gnrPCASampLoading(SEXP EigenCnt, SEXP SNPLoadings, SEXP AvgFreq, SEXP Scale,
                                            SEXP NumThread, SEXP Verbose) {
    COREARRAY_TRY
        Ctx c;
        load_workspace(c);
        rv_ans = PROTECT(Rf_allocMatrix(REALSXP, MCWorkingGeno.Space().SampleNum(), Rf_asInteger(EigenCnt)));
        c.ck(snprel_pca_samp_loading(c.h, Rf_asInteger(EigenCnt), REAL(SNPLoadings), REAL(AvgFreq), REAL(Scale),
                                     REAL(rv_ans)));
        UNPROTECT(1);
    COREARRAY_CATCH
}

COREARRAY_DLL_EXPORT SEXP gnrEigMixSNPLoading(SEXP EigenVal, SEXP EigenVect, SEXP AFreq, SEXP NumThread,
                                              SEXP Verbose) {
    const int LenEig = INTEGER(GET_DIM(EigenVect))[1];
    COREARRAY_TRY
        Ctx c;
        load_workspace(c);
        rv_ans = PROTECT(Rf_allocMatrix(REALSXP, LenEig, MCWorkingGeno.Space().SNPNum()));
        c.ck(snprel_eigmix_snp_loading(c.h, LenEig, REAL(EigenVal), REAL(EigenVect), REAL(AFreq), REAL(rv_ans)));
        UNPROTECT(1);
    COREARRAY_CATCH
}

COREARRAY_DLL_EXPORT SEXP gnrEigMixSampLoading(SEXP SNPLoadings, SEXP AFreq, SEXP NumThread, SEXP Verbose) {
    const int EigenCnt = INTEGER(GET_DIM(SNPLoadings))[0];
    COREARRAY_TRY
        Ctx c;
        load_workspace(c);
        rv_ans = PROTECT(Rf_allocMatrix(REALSXP, MCWorkingGeno.Space().SampleNum(), EigenCnt));
        c.ck(snprel_eigmix_samp_loading(c.h, EigenCnt, REAL(SNPLoadings), REAL(AFreq), REAL(rv_ans)));
        UNPROTECT(1);
    COREARRAY_CATCH
}

}  // extern "C"
#endif  // SNPREL_BUILD_R_SHIM
