/*
 * snprel_b200.h -- C ABI of libsnprel_b200.so
 *
 * B200-native (sm_100a) replacement for ONE path of zhengxwen/SNPRelate: the
 * pairwise N x N relatedness-matrix accumulation over 2-bit SNP genotype blocks
 * behind snpgdsGRM / snpgdsPCA / snpgdsEIGMIX / snpgdsIBS / snpgdsIBSNum /
 * snpgdsIBDKING / snpgdsIndivBeta.  Every entry point below names the reference
 * interface (file:line under the reference's source tree) it stands in for; the
 * R-side binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 on success and a
 *     non-zero code on failure, never throws; snprel_last_error() returns the
 *     message (the reference stores it via GDS_SetError and exposes it through
 *     gnrErrMsg, src/SNPRelate.cpp:1099-1104).
 *   - a context owns one CUDA device; it is the analogue of the reference's
 *     process-global GWAS::MCWorkingGeno (src/dGenGWAS.cpp:2000): first the
 *     genotype workspace is set, then estimators run on it with no genotype
 *     argument.  One context per host thread; not re-entrant (like the
 *     reference, src/dGenGWAS.cpp:2198-2200).
 *   - host genotype blocks are SNP-major, sample-fastest uint8 [cnt][n_samp]
 *     with 0/1/2 = #A alleles and anything > 2 = missing: exactly what
 *     CdBaseWorkSpace::snpRead(start, cnt, buf, RDim_Sample_X_SNP) produces
 *     (src/dGenGWAS.h:94, src/dGenGWAS.cpp:677-733).
 *   - n x n outputs are full symmetric column-major doubles (Rf_allocMatrix
 *     layout, src/genPCA.cpp:1593-1594) or, with packed != 0, the row-packed
 *     upper triangle of n(n+1)/2 doubles (CdMatTri order, src/dGenGWAS.h:556-561;
 *     what useMatrix=TRUE returns, src/genPCA.cpp:1596-1598).
 *   - there is no CPU fallback: every call fails loudly without a device.
 */
#ifndef SNPREL_B200_H
#define SNPREL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct snprel_ctx snprel_ctx;

/* ---- context ---------------------------------------------------------- */

/* Create a context on CUDA device `device`. */
int snprel_create(snprel_ctx **out, int device);
void snprel_destroy(snprel_ctx *ctx);
/* Last error message of this context (or of the failed create when ctx==NULL).
 * Replaces gnrErrMsg (src/SNPRelate.cpp:1099-1104). */
const char *snprel_last_error(snprel_ctx *ctx);
/* Library version string. */
const char *snprel_version(void);
/* Number of CUDA devices visible to the process (0 when there is none). */
int snprel_device_count(void);

/* ---- genotype workspace  (gnrSetGenoSpace, src/SNPRelate.cpp:76-114;
 *      CGenoReadBySNP block iterator, src/dGenGWAS.cpp:1218-1397) ---------- */

/* Start a workspace of n_samp samples and room for up to snp_capacity SNPs.
 * Device layout: 2-bit genotypes, SNP-major rows padded to 256 samples. */
int snprel_geno_begin(snprel_ctx *ctx, int64_t n_samp, int64_t snp_capacity);
/* Append `cnt` SNPs from a HOST uint8 block [cnt][n_samp] (what one
 * CGenoReadBySNP::Read() yields).  Packed to 2 bits on the device. */
int snprel_geno_push_u8(snprel_ctx *ctx, const uint8_t *geno, int64_t cnt);
/* Append `cnt` SNPs from HOST 2-bit rows (4 genotypes per byte, LSB first,
 * code 3 = missing; GDS dBit2 encoding) with `row_bytes` bytes per SNP row. */
int snprel_geno_push_2b(snprel_ctx *ctx, const uint8_t *packed, int64_t cnt,
                        int64_t row_bytes);
/* The same push without waiting for the copy (the reference double-buffers its block reader against
 * the compute threads, src/dGenGWAS.cpp:1298-1324): the rows go out in chunks on a copy stream and the
 * call returns at once.  `packed` must stay valid (pinned memory recommended) until the next call on
 * this context returns.  snprel_pca / snprel_grm (Eigenstrat, GCTA, Corr, EIGMIX) / snprel_eigmix that
 * follow consume the chunks as they arrive -- statistics, digit tables and tensor passes of a chunk
 * run while later chunks are still crossing PCIe; the fixed-point format is chosen from the first
 * chunk's statistics extrapolated with safety margins and verified against the true statistics at the
 * end (on failure the ordinary path recomputes everything on the resident data; snprel_stream_stats
 * counts both outcomes).  Every other entry point first waits for the copies (snprel_geno_wait). */
int snprel_geno_push_2b_async(snprel_ctx *ctx, const uint8_t *packed, int64_t cnt, int64_t row_bytes);
int snprel_geno_wait(snprel_ctx *ctx);
int snprel_stream_stats(snprel_ctx *ctx, int64_t *streamed, int64_t *fallbacks);
/* Per-item clock stamps of the last table-Gram launch made with snprel_debug_flags(ctx, 1) (tools/k1_trace.py):
 * out[items][8] = entry, set-up done, first MMA, last commit issued, MMAs complete, epilogue done, exit, stages. */
int snprel_k1_trace(snprel_ctx *ctx, int64_t *out, int64_t capacity_items, int64_t *items);
/* device time from the first asynchronous copy chunk being queued to the last one having arrived */
int snprel_stream_last_copy_ms(snprel_ctx *ctx, double *ms);
/* Append `cnt` SNPs straight from the payload of an uncompressed GDS dBit2 genotype node
 * (sample.order layout): one continuous LSB-first 2-bit stream, sample fastest, with NO per-row
 * padding, so rows are not byte aligned when n_samp % 4 != 0 (what CdSNPWorkSpace::snpRead
 * decodes through gdsfmt, src/dGenGWAS.cpp:677-733).  `stream` points to the first byte of the
 * payload, `first_genotype` = index (in genotypes) of the first genotype of the first SNP to push,
 * i.e. first_snp * n_samp.  EXPERIMENTAL: written at the end of round 1, not yet run on a GPU. */
int snprel_geno_push_bitstream(snprel_ctx *ctx, const uint8_t *stream,
                               int64_t first_genotype, int64_t cnt);
/* SNP-sharded loading of a workspace that finally holds ALL SNPs (N x N output tiled across GPUs when
 * N^2 exceeds one GPU's HBM): reserve the whole SNP range with snprel_geno_begin, _seek to
 * this rank's first SNP, push its block (only that block crosses PCIe), receive the other ranks'
 * blocks straight into the device rows (an NCCL all-gather / broadcast over NVLink on the pointer
 * _device_rows returns, pitch row_bytes), then _commit the total row count. */
int snprel_geno_seek(snprel_ctx *ctx, int64_t snp_index);
int snprel_geno_device_rows(snprel_ctx *ctx, void **dev_ptr, int64_t *row_bytes, int64_t *capacity);
int snprel_geno_commit(snprel_ctx *ctx, int64_t n_snp);
/* Fill the workspace with `n_snp` synthetic SNPs on the device (counter-based
 * generator; SNP l uses global index snp_start+l so SNP shards of one data
 * set can be generated independently; a shard, i.e. snp_start beyond the write
 * position, also sets the SNP origin, see snprel_set_snp_origin).  Benchmark /
 * test input only. */
int snprel_geno_synth(snprel_ctx *ctx, int64_t n_snp, uint64_t seed,
                      double maf_lo, double maf_hi, double miss_rate,
                      int64_t snp_start);
/* gnrGetGenoDim (src/SNPRelate.cpp): current selected dimensions. */
int snprel_geno_dim(snprel_ctx *ctx, int64_t *n_samp, int64_t *n_snp);
/* Copy the workspace back as HOST uint8 [n_snp][n_samp] (gnrCopyGenoMem,
 * src/SNPRelate.cpp:322). */
int snprel_geno_copy_u8(snprel_ctx *ctx, uint8_t *out);
/* Copy the workspace back as HOST 2-bit rows (the encoding snprel_geno_push_2b
 * takes), `row_bytes` >= ceil(n_samp/4) bytes per SNP row. */
int snprel_geno_copy_2b(snprel_ctx *ctx, uint8_t *out, int64_t row_bytes);

/* gnrSNPRateFreq (src/SNPRelate.cpp:243) / Get_AF_MR_perSNP
 * (src/dGenGWAS.cpp:472-552): per-SNP allele frequency, minor allele frequency
 * and missing rate (any pointer may be NULL). NaN where all samples missing. */
int snprel_snp_ratefreq(snprel_ctx *ctx, double *af, double *maf, double *mr);
/* gnrSelSNP_Base (src/SNPRelate.cpp:184, src/dGenGWAS.cpp:361-397): drop SNPs
 * failing the filters and compact the workspace.  maf < 0 / missrate > 1
 * disable that filter (R passes -1 / 2 for NaN, R/Internal.R:438-439).
 * out_sel (nullable) receives the 0/1 keep flag per SNP BEFORE compaction;
 * n_removed (nullable) the number of SNPs dropped. */
int snprel_select_snp_base(snprel_ctx *ctx, int remove_mono, double maf,
                           double missrate, uint8_t *out_sel,
                           int64_t *n_removed);
/* gnrSelSNP_Base_Ex -> CdBaseWorkSpace::Select_SNP_Base_Ex
 * (src/dGenGWAS.cpp:399-470): the same filter with the MAF taken from
 * caller-supplied allele frequencies afreq[n_snp]; non-finite = dropped. */
int snprel_select_snp_base_ex(snprel_ctx *ctx, const double *afreq,
                              int remove_mono, double maf, double missrate,
                              uint8_t *out_sel, int64_t *n_removed);

/* ---- packed-bit estimators (integer, bit exact) ----------------------- */

/* gnrIBSNum (src/genIBS.cpp:500-550): three n x n int32 matrices. */
int snprel_ibs_num(snprel_ctx *ctx, int32_t *ibs0, int32_t *ibs1, int32_t *ibs2);
/* gnrIBSAve (src/genIBS.cpp:441-497). */
int snprel_ibs_ave(snprel_ctx *ctx, double *out, int packed);
/* gnrIBD_PLINK (src/genIBS.cpp:558-639): PLINK method-of-moments k0, k1 from the
 * IBS counters.  afreq_in (nullable, double[n_snp]): user allele frequencies
 * (NaN or outside [0,1] = SNP ignored, no finite-sample correction factor, as
 * R/IBD.R:51-52 + src/genIBD.cpp:253-338); NULL = frequencies and correction
 * from the genotype counts.  afreq_out (nullable) receives the frequencies used. */
int snprel_ibd_mom(snprel_ctx *ctx, const double *afreq_in, int kinship_constraint,
                   int packed, double *k0, double *k1, double *afreq_out);
/* The same in two steps, for SNP-sharded ranks: sums6 = the five un-normalised
 * expectation sums {a00,a01,a02,a11,a12} + nValid over this context's SNPs (add
 * them across ranks), then the pair epilogue on the (reduced) IBS counters. */
int snprel_ibd_mom_sums(snprel_ctx *ctx, const double *afreq_in, double *sums6,
                        double *afreq_out);
int snprel_ibd_mom_from_sums(snprel_ctx *ctx, const double *sums6,
                             int kinship_constraint, int packed, double *k0,
                             double *k1);
/* gnrIBD_KING_Robust (src/genKING.cpp:576-679).  family_id: int[n_samp], same
 * id (and not SNPREL_NA_INT) = within-family estimator; NULL = all unrelated. */
int snprel_king_robust(snprel_ctx *ctx, const int32_t *family_id, double *ibs0,
                       double *kinship, int packed);
/* Raw KING-robust counters (TS_KINGRobust, src/genKING.cpp:274-281) as five
 * n x n int32 matrices: IBS0, nLoci, SumSq, N1_Aa, N2_Aa (row = sample 1). */
int snprel_king_robust_counts(snprel_ctx *ctx, int32_t *out5);
/* gnrIBD_KING_Homo (src/genKING.cpp:493-570): k0, k1. */
int snprel_king_homo(snprel_ctx *ctx, double *k0, double *k1, int packed);
/* gnrIBD_Beta (src/genBeta.cpp:361-460).  avg_out (nullable) receives the
 * value the reference publishes through gnrGRM_avg_val (src/genPCA.cpp:1608). */
int snprel_indiv_beta(snprel_ctx *ctx, int inbreeding, double *out, int packed,
                      double *avg_out);
/* Raw IndivBeta counters (TS_Beta, src/genBeta.cpp:50-54): ibscnt, num. */
int snprel_indiv_beta_counts(snprel_ctx *ctx, int32_t *out2);

/* Engine of the pair counters above: 0 (default) = packed-bit XOR/AND/popcount kernels
 * (csrc/bitcount.cu, the reference's formulation); 1 = the same uint32 counters as exact
 * {-1,0,1}-table Grams on the tensor pipe (csrc/count_tc.cu:tensor_count_accumulate), bit-identical
 * and several times faster because the bit kernels are bound by the integer ALU. */
int snprel_set_count_engine(snprel_ctx *ctx, int engine);

/* ---- covariance-type estimators (tcgen05 table-Gram) ------------------ */

#define SNPREL_GRM_EIGENSTRAT 0
#define SNPREL_GRM_GCTA       1
#define SNPREL_GRM_CORR       2
#define SNPREL_GRM_EIGMIX     3
#define SNPREL_GRM_INDIVBETA  4

#define SNPREL_NA_INT INT32_MIN

/* gnrGRM (src/genPCA.cpp:1614-1717): method is one of SNPREL_GRM_*.
 * "Corr" ignores `packed` like the reference (src/genPCA.cpp:1658-1685).
 * avg_out (nullable): gnrGRM_avg_val for IndivBeta. */
int snprel_grm(snprel_ctx *ctx, int method, double *out, int packed,
               double *avg_out);
/* gnrPCA, algorithm "exact" (src/genPCA.cpp:1355-1452): covariance
 * (genmat, nullable), TraceXTX, TraceVal, and the top eigen_cnt eigenpairs of
 * the normalised matrix (eigval[n_samp] with NaN beyond eigen_cnt, eigvec
 * n_samp x eigen_cnt column-major; both nullable => genmat.only). */
int snprel_pca(snprel_ctx *ctx, int eigen_cnt, int bayesian, double *genmat,
               double *trace_xtx, double *trace_val, double *eigval,
               double *eigvec);
/* gnrEigMix (src/genEIGMIX.cpp:656-735): ibd (nullable), afreq[n_snp]
 * (nullable), eigenpairs as in snprel_pca (eigen_cnt == 0 => none). */
int snprel_eigmix(snprel_ctx *ctx, int eigen_cnt, int diagadj, double *ibd,
                  double *afreq, double *eigval, double *eigvec);

/* Rounding of the fixed-point row table T of the covariance estimators' main passes.
 * 0: round to nearest; the format is chosen from the worst-case error bound
 *    2^-(frac_bits+1) * err_weight (every rounding error at its maximum with the same sign).
 * 1: unbiased randomised rounding, floor(T 2^frac_bits + u) with one uniform draw u per (SNP,
 *    genotype) table entry from a counter-based generator keyed by the SNP's global index
 *    (snprel_set_snp_origin).  The error of an entry is then a sum of independent zero-mean terms
 *    of width |B_l| 2^-frac_bits, and by Hoeffding's inequality it stays below
 *    2^-frac_bits * sqrt(1/2 sum_l B_l^2 ln(2 #pairs / 1e-12)) for ALL entries simultaneously with
 *    probability >= 1 - 1e-12 over the draws, for any data set; max_j sum_l B_l[g_jl]^2 is measured
 *    (snprel_plan.err_weight2, else <= 127 err_weight).  The bound grows with sqrt(#SNPs) where the
 *    worst case grows with #SNPs: from about 1e5 SNPs on it saves one base-256 digit, i.e. one of
 *    eight tensor passes at config-2 size (10 000 samples x 1 000 000 SNPs).
 * 2 (default): whichever of the two needs FEWER tensor passes for the requested tolerance; ties
 *    go to 0, and so do small problems (n_samp^2 * n_snp < 2^36, where a pass costs microseconds and
 *    round-to-nearest leaves the larger margin).  snprel_plan.rounding reports what the last
 *    accumulate used.
 * Results are a pure function of (genotypes, SNP origin, mode): run-to-run identical, and identical
 * for any SNP sharding whose origins are the shards' global offsets.  All ranks of a multi-GPU run
 * must use the same mode.  The environment variable SNPREL_ROUNDING = nearest | random | auto gives the
 * initial mode of every context created afterwards (R sessions, the Python host). */
int snprel_set_rounding(snprel_ctx *ctx, int mode);
/* Global index of this context's first SNP row (default 0; snprel_geno_begin resets it).  Only keys
 * the rounding draws of mode 1 / 2: SNP shards of one data set must cover disjoint index ranges
 * (snprel_multi_* sets the shard offsets itself, snprel_geno_synth does for generated shards; a host that
 * pushes its own SNP ranges into several contexts calls this after snprel_geno_begin). */
int snprel_set_snp_origin(snprel_ctx *ctx, int64_t origin);


/* The eigen step of snprel_pca / snprel_eigmix (CalcEigen, LAPACK dspevx on -C,
 * src/genPCA.cpp:1262-1346) runs in csrc/eigen.cu: a Chebyshev-filtered subspace iteration
 * built from cuBLAS / cuSOLVER calls when eigen_cnt << n_samp (n_samp >= 2048, 8 eigen_cnt <=
 * n_samp), otherwise (or if its residuals stall above 1e-10) a full cusolverDnXsyevd.  What the
 * last solve did: solver 0 = dense, 1 = filtered subspace iteration; its filter rounds and the
 * number of n x n by n x b block products; phase_ms[3] (nullable): milliseconds spent in the
 * filter, the orthonormalisation and the Rayleigh-Ritz steps.  snprel_debug_flags(ctx, 4) forces
 * the dense solver. */
int snprel_last_eigen_info(snprel_ctx *ctx, int *solver, int *rounds, int *block_gemms,
                           double *phase_ms);

/* ---- loadings, projection of new samples, SNP-PC correlation ----------
 * The tall-skinny float64 products either side of the eigen-decomposition
 * (csrc/project.cu).  Matrix layouts are the reference's R layouts: an
 * "n x k column-major" eigenvector matrix is eigvect[i + kk * n]; a
 * "k x n_snp column-major" loading / correlation matrix is m[kk + l * k]. */

/* gnrPCASNPLoading (src/genPCA.cpp:1489-1540, CPCA_SNPLoad :938-1040): SNP loadings of the
 * workspace's samples against the top k eigenpairs of snprel_pca.  eigval[k], eigvect n x k,
 * trace_xtx and bayesian as returned by / passed to snprel_pca.  loading: k x n_snp;
 * avgfreq[n_snp] (mean genotype) and scale[n_snp] (nullable) are what
 * snprel_pca_samp_loading needs later. */
int snprel_pca_snp_loading(snprel_ctx *ctx, int k, const double *eigval,
                           const double *eigvect, double trace_xtx, int bayesian,
                           double *loading, double *avgfreq, double *scale);
/* gnrPCASampLoading (src/genPCA.cpp:1542-1563, CPCA_SampleLoad :1042-1123): project the
 * workspace's samples (typically NEW samples over the loading's SNPs) onto k components.
 * loadings: k x n_snp, already scaled by sqrt(((n0-1)/TraceXTX)/eigenval) as R/PCA.R:274-277
 * does; out: n_samp x k column-major. */
int snprel_pca_samp_loading(snprel_ctx *ctx, int k, const double *loadings,
                            const double *avgfreq, const double *scale, double *out);
/* gnrPCACorr (src/genPCA.cpp:1456-1485, CPCA_SNPCorr :809-936): Pearson correlation between
 * every SNP and each of k eigenvectors (n x k) over the SNP's non-missing samples; NaN when
 * undefined.  out: k x n_snp. */
int snprel_pca_corr(snprel_ctx *ctx, int k, const double *eigvect, double *out);
/* gnrEigMixSNPLoading (src/genEIGMIX.cpp:739-775, CEigMix_SNPLoad :445-530); afreq[n_snp]
 * as returned by snprel_eigmix. */
int snprel_eigmix_snp_loading(snprel_ctx *ctx, int k, const double *eigval,
                              const double *eigvect, const double *afreq,
                              double *loading);
/* gnrEigMixSampLoading (src/genEIGMIX.cpp:777-803, CEigMix_SampleLoad :534-640); loadings
 * k x n_snp scaled by sqrt(1/eigenval) (R/PCA.R:296-297); out n_samp x k. */
int snprel_eigmix_samp_loading(snprel_ctx *ctx, int k, const double *loadings,
                               const double *afreq, double *out);

/* gnrPCA algorithm "randomized" (src/genPCA.cpp:1436-1442, CRandomPCA :469-796).
 * aux_mat: the R vector rnorm(aux.dim * n.samp) (R/PCA.R:56), i.e. G_0 as a column-major
 * n_samp x aux_dim matrix; not modified (the reference overwrites it).  Outputs, as the three list
 * elements of the reference: sigma[n_samp] singular values of T (zero padded), vt = the rows of
 * V_T^T, row-major [aux_dim * (iter_num + 1)][n_samp] (rows beyond min(hsize, n_samp) are zero), and
 * 2 * TraceXTX.  R/PCA.R:80-89 turns them into eigenval / eigenvect / varprop.  Leading singular
 * values / vectors agree with the reference run with num.thread = 1 (subspace angle < 1e-6); the
 * trailing directions of the ill-conditioned Krylov basis are not comparable between LAPACKs. */
int snprel_pca_randomized(snprel_ctx *ctx, const double *aux_mat, int aux_dim, int iter_num,
                          double *sigma, double *vt, double *trace_xtx2);

/* ---- split accumulate / reduce / finish (multi-GPU SNP sharding) ------- */

#define SNPREL_EST_IBS          10
#define SNPREL_EST_KING_ROBUST  11
#define SNPREL_EST_BETA         12
#define SNPREL_EST_KING_HOMO    13
/* estimator ids 0..3 are the SNPREL_GRM_* covariance methods */

/* Per-rank plan statistics of the covariance estimators.  All ranks must agree on
 * the fixed-point format, so the host max-reduces max_abs and sum-reduces the other
 * statistics before snprel_accumulate (see snprelate_b200/dist.py).
 *
 * Format choice (grm.cu): every row-table value v is stored as round(v * 2^frac_bits) and
 * split into `digits` balanced base-256 digits (one int8 tensor pass each) and multiplied
 * with a per-SNP INTEGER column table B_l (main passes) or the missing indicator.  The
 * quantisation error of an output entry is at most
 *   2^-(frac_bits+1) * err_weight + 2^-(frac_bits_w+1) * max_missing
 * (first term with randomised rounding: the Hoeffding bound of snprel_set_rounding),
 * so the library picks the fewest passes for which that bound is <= tol * scale, where
 * scale is (a lower bound of) the estimator's normaliser (trace/(n-1), 2 nLocus,
 * sum 4p(1-p)).  tol defaults to 1e-10. */
typedef struct snprel_plan {
    double max_abs;        /* max |T| over local SNPs, T = U / s the row table of the main passes */
    double max_abs_w;      /* max |R|, R = (mu - t/s) U the row table of the missing-data passes  */
    double sum_bound;      /* sum over local SNPs of the per-SNP magnitude (int64 range) */
    double err_weight;     /* max over samples of sum_l |B_l[g]|, B_l = s g - t the integer column table */
    double scale;          /* local share of the normaliser of the final matrix        */
    double tol;            /* in: relative tolerance target (<= 0: 1e-10)              */
    int64_t total_missing; /* missing genotypes among the selected samples             */
    int64_t max_missing;   /* max over samples of the local missing count              */
    int64_t n_snp;         /* local SNP count                                          */
    int32_t frac_bits;     /* in: < 0 = choose; out (after accumulate): chosen          */
    int32_t frac_bits_w;   /* fixed point of the W table (<= frac_bits)                */
    int32_t frac_bits_d;   /* fixed point of the missing-pair denominator plane        */
    int32_t digits;        /* out: digits (tensor passes) of the U table               */
    int32_t digits_w;      /* out: digits of the W table (0 without missing data)      */
    int32_t digits_d;      /* out: digits of the denominator table                     */
    int32_t bayesian;      /* Eigenstrat only                                          */
    int32_t frac_bits_v;   /* out: fixed point of the per-sample vector sum_l R_l[g_il] */
    int32_t rounding;      /* frac_bits < 0 (library chooses): out, 0 = nearest / 1 = randomised (snprel_set_rounding);
                              caller-fixed frac_bits: in, the rounding that goes with it */
    double diag_bound;     /* measured X >= max_i C_ii of the local SNPs (0: not measured); sum-reduced  */
    double sum_rest;       /* the part of sum_bound that is not the main T x B product; sum-reduced       */
    double err_weight2;    /* >= max over samples of sum_l B_l[g]^2 over the local SNPs (0: not measured, 127 err_weight
                              is used); sum-reduced.  Variance proxy of the randomised-rounding bound */
} snprel_plan;

/* sizeof(snprel_plan) of the library build: a binding that mirrors the struct (ctypes, R) checks its own
 * layout against it when it loads the library. */
int64_t snprel_abi_sizeof_plan(void);
int snprel_plan_local(snprel_ctx *ctx, int estimator, snprel_plan *plan);
/* Host-only (no device, no context): the fixed-point format the library chooses for the given
 * (merged) plan statistics -- fills frac_bits*, digits*, rounding exactly as snprel_accumulate
 * would.  mode as in snprel_set_rounding; n_samp enters the union bound of mode 1.  Returns the
 * number of tensor passes, or -1 (snprel_last_error(NULL)). */
int snprel_plan_format(int estimator, snprel_plan *plan, int mode, int64_t n_samp);
/* Accumulate this rank's SNPs into the device accumulators. */
int snprel_accumulate(snprel_ctx *ctx, int estimator, const snprel_plan *plan);
/* Enumerate the device buffers that must be sum-reduced across ranks.
 * kind: 0 = int64, 1 = uint32, 2 = float64.  Returns the number of buffers;
 * call with idx in [0, count). */
int snprel_reduce_buffer_count(snprel_ctx *ctx);
int snprel_reduce_buffer(snprel_ctx *ctx, int idx, void **dev_ptr,
                         int64_t *count, int *kind);
/* After the reduction: the snprel_grm / snprel_pca / snprel_eigmix /
 * snprel_ibs_* / snprel_king_* / snprel_indiv_beta calls above finish from the
 * reduced accumulators instead of accumulating again. */
int snprel_mark_reduced(snprel_ctx *ctx);
/* The plan (with the chosen frac_bits / digits) of the last covariance accumulate. */
int snprel_last_plan(snprel_ctx *ctx, snprel_plan *plan);

/* ---- row windows: N x N outputs that do not fit in HBM (or are tiled across GPUs) ---- */

/* Restrict the accumulators and the results of the following estimator calls to output rows
 * [row0, row0 + rows) (both multiples of 256; rows == 0 restores the whole matrix).  Inside a
 * window snprel_grm (Eigenstrat / GCTA / EIGMIX), snprel_ibs_ave, snprel_king_robust and
 * snprel_king_homo require packed != 0 and write only the window's slice of the row-packed
 * upper triangle -- entries idx(row0,row0) .. idx(r1-1, n-1), contiguous in CdMatTri order
 * (src/dGenGWAS.h:556-561); snprel_ibs_num writes three packed int32 slices.  The host walks the
 * windows (one GPU: sequentially; several GPUs holding the same genotypes: windows dealt
 * round-robin, no collective) and concatenates the slices. */
int snprel_set_row_window(snprel_ctx *ctx, int64_t row0, int64_t rows);
/* Number of packed entries the current window produces. */
int snprel_window_count(snprel_ctx *ctx, int64_t *count);
/* Free / total device memory in bytes (cudaMemGetInfo): the host sizes row windows with it. */
int snprel_mem_info(snprel_ctx *ctx, int64_t *free_bytes, int64_t *total_bytes);

/* ---- introspection for benchmarks / tests ------------------------------ */

/* Number of kernels this library launched on the context so far. */
int64_t snprel_kernel_launches(snprel_ctx *ctx);
/* Device-side duration (ms, CUDA events on the library's stream) and launch
 * count of the dominant kernel of the last accumulate call. */
int snprel_last_hot_kernel(snprel_ctx *ctx, double *ms, int64_t *launches,
                           double *algorithmic_units);
/* Run the estimator's whole accumulation (per-SNP statistics, tables / bit-plane
 * transpose, pair kernels) `reps` times on the resident 2-bit workspace and return
 * the average device time per repetition in ms, measured with CUDA events on the
 * library's stream (bench.py `value`). */
int snprel_time_accumulate(snprel_ctx *ctx, int estimator, int reps, double *ms);
/* Covariance estimators: run the device epilogue (int64 planes -> final float64 matrix,
 * left on the device) for the accumulators at hand -- after snprel_accumulate, or after the
 * all-reduce + snprel_mark_reduced of a sharded run -- and return its device time in ms (also
 * added to snprel_last_step_ms).  snprel_time_accumulate includes it: the metric runs up to
 * "final N x N complete" (the reference's gnrGRM / gnrPCA end with the same normalisation loop,
 * src/genPCA.cpp:1381-1390, :1233-1236). */
int snprel_time_finish(snprel_ctx *ctx, int estimator, double *ms);
/* Device time (ms, CUDA events) of the last snprel_plan_local + snprel_accumulate pair. */
int snprel_last_step_ms(snprel_ctx *ctx, double *ms);
/* Drop everything derived from the resident 2-bit matrix (per-SNP statistics, bit
 * planes, accumulators) so that the next call recomputes it. */
int snprel_invalidate(snprel_ctx *ctx);
/* Exact integer table-Gram of the tcgen05 kernel for tests:
 * out[i][j] = sum_l tabA[l][g_il] * tabB[g_jl], int8 tables, int64 out
 * (n_samp x n_samp row-major, all entries). */
int snprel_table_gram(snprel_ctx *ctx, const int8_t *tabA /*[n_snp][4]*/,
                      const int8_t *tabB /*[4]*/, int64_t *out);
/* Test / experiment hooks (OR of): 1 = record per-item clock stamps of the table Gram (snprel_k1_trace);
 * 2 = one CTA pair walks the whole SNP range of a tile (no SNP splits); 4 = always use the dense eigen
 * solver; table-Gram A/B knobs measured in profiles/r02_k1_variants*.log: 16 = the round-1 16-byte genotype
 * boxes, 32 / 64 = no / 256-byte L2 promotion, 128 = group-major item order, 256 = no SNP segments,
 * 0xN000 = N SNP segments; pair counters: 0x400 / 0x800 / 0xC00 = 1 / 2 / 3 streams straight to POPC
 * (profiles/r02_pair_variants.log).  None of them changes a result. */
int snprel_debug_flags(snprel_ctx *ctx, uint32_t flags);

/* Asynchronous result delivery for row-window pipelines (C4 / C5 style jobs): with on = 1 snprel_ibs_ave,
 * snprel_king_robust and snprel_grm return as soon as the device-to-host copy of their result has been
 * QUEUED (on a second stream), so the copy of window w overlaps the accumulate of window w + 1; the host
 * buffer (pinned memory) must not be read or reused before snprel_output_wait returns.  The next finishing
 * call on the context waits by itself before it rewrites the device-side result scratch. */
int snprel_set_async_output(snprel_ctx *ctx, int on);
int snprel_output_wait(snprel_ctx *ctx);

/* ---- several GPUs of one box behind one handle (single host process) --------------------
 * What an R session needs to spread gnrGRM / gnrPCA / gnrIBSNum / gnrIBD_KING_Robust ... over the
 * eight B200s of a box (north star: "partition across the 8 GPUs by SNP block, each GPU
 * accumulating a partial N x N matrix with one all-reduce over NVLink at the end"): device r owns
 * the r-th contiguous SNP range (what CGenoReadBySNP hands out block by block,
 * src/dGenGWAS.cpp:1218-1397, is routed by SNP index), every device accumulates its partial planes
 * on its own host thread with one agreed fixed-point format, and a hand-written peer-memory
 * reduction over NVLink (reduce-scatter by peer loads, then the finishing device pulls the
 * reduced slices; upper-triangle columns only) leaves the global accumulators on device `root`.
 * The finishing calls of this header (snprel_grm, snprel_pca, snprel_ibs_num, snprel_king_robust,
 * snprel_indiv_beta, ...) are then made on snprel_multi_ctx(m, root).  Per-SNP outputs (afreq of
 * snprel_eigmix, snprel_snp_ratefreq) of a context cover that device's SNP range only; ranges are
 * consecutive, in device order.  `devices` may repeat a GPU (tests on a one-GPU box). */
typedef struct snprel_multi snprel_multi;
int snprel_multi_create(const int *devices, int n_dev, snprel_multi **out);
void snprel_multi_destroy(snprel_multi *m);
const char *snprel_multi_last_error(snprel_multi *m);
int snprel_multi_device_count(snprel_multi *m);
snprel_ctx *snprel_multi_ctx(snprel_multi *m, int i);
/* gnrSetGenoSpace + the block reader, routed: same meaning as the single-device calls */
int snprel_multi_geno_begin(snprel_multi *m, int64_t n_samp, int64_t snp_capacity);
int snprel_multi_geno_push_u8(snprel_multi *m, const uint8_t *geno, int64_t cnt);
int snprel_multi_geno_push_2b(snprel_multi *m, const uint8_t *packed, int64_t cnt, int64_t row_bytes);
int snprel_multi_geno_synth(snprel_multi *m, int64_t n_snp, uint64_t seed, double maf_lo, double maf_hi,
                            double miss_rate, int64_t snp_start);
int snprel_multi_set_row_window(snprel_multi *m, int64_t row0, int64_t rows);
int snprel_multi_set_count_engine(snprel_multi *m, int engine);
int snprel_multi_set_rounding(snprel_multi *m, int mode);      /* snprel_set_rounding on every device */
/* plan (all devices) -> merged format -> accumulate (all devices) -> peer reduction -> mark reduced.
 * est: SNPREL_GRM_EIGENSTRAT / GCTA / CORR / EIGMIX or SNPREL_EST_IBS / KING_ROBUST / BETA.
 * root: device index (position in `devices`) that receives the reduced N x N planes; -1: every device
 * (an all-reduce).  The small per-sample buffers always go to every device. */
int snprel_multi_accumulate(snprel_multi *m, int est, int bayesian, int root);
/* duration (ms) and link traffic (bytes) of the last peer reduction */
int snprel_multi_last_reduce(snprel_multi *m, double *ms, int64_t *bytes);

/* The same peer-memory reduction for ONE PROCESS PER GPU (torchrun, bench.py --gpus N): every rank
 * exports a CUDA IPC handle (+ byte offset inside the allocation) for each of its reduce buffers
 * (snprel_reduce_buffer), the host gathers them from all ranks ([world][n_buffers], 64 bytes each) and
 * opens them; phase 1 sums this rank's row slice of every buffer out of the peers' HBM, phase 2 pulls
 * the other ranks' reduced slices (all ranks when root < 0, else the root; the small per-sample
 * buffers always).  The caller places a cross-process barrier before phase 1, between the phases and
 * after phase 2, then calls snprel_mark_reduced (snprelate_b200/dist.py:peer_reduce_buffers). */
int snprel_reduce_ipc_export(snprel_ctx *ctx, int idx, void *handle64, int64_t *offset);
int snprel_peer_reduce_open(snprel_ctx *ctx, int rank, int world, const void *handles, const int64_t *offsets);
int snprel_peer_reduce_phase(snprel_ctx *ctx, int phase, int root, int64_t *link_bytes);
void snprel_peer_reduce_close(snprel_ctx *ctx);

/* Tiled N x N output over several devices (N^2 exceeds one GPU's HBM): _begin_replicated reserves the
 * WHOLE SNP range on every device but routes each pushed block to its owner only (1/n of the data per
 * PCIe link); _geno_gather then completes every device's copy with peer copies over NVLink.
 * _grm_tiled walks the upper triangle in row windows of `window_rows` (multiple of 256; 0 = chosen from
 * the free memory), window w on device w mod n, each device accumulating its windows over all SNPs with
 * no reduction, and hands the packed rows (CdMatTri order, the layout gnrGRM returns with useMatrix) to
 * `sink` strictly in order: sink(user, first_packed_index, values, count) != 0 aborts.  Methods:
 * SNPREL_GRM_EIGENSTRAT / GCTA / EIGMIX. */
int snprel_multi_geno_begin_replicated(snprel_multi *m, int64_t n_samp, int64_t snp_capacity);
int snprel_multi_geno_gather(snprel_multi *m);
typedef int (*snprel_sink_fn)(void *user, int64_t first_packed_index, const double *values, int64_t count);
int snprel_multi_grm_tiled(snprel_multi *m, int method, int64_t window_rows, snprel_sink_fn sink, void *user);

#ifdef __cplusplus
}
#endif
#endif /* SNPREL_B200_H */
