// Internal definitions shared by the translation units of libsnprel_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/snprel_b200.h"

namespace snprel {

// ---------------------------------------------------------------------------
// errors: never throw across the C ABI
// ---------------------------------------------------------------------------
struct Error {
    std::string msg;
};

[[noreturn]] inline void fail(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw Error{buf};
}

#define CUDA_CHECK(expr)                                                        \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess)                                                  \
            ::snprel::fail("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e),  \
                           __FILE__, __LINE__, cudaGetErrorString(_e));         \
    } while (0)

// ---------------------------------------------------------------------------
// device buffer (RAII over cudaMalloc) -- the allocator of this library
// ---------------------------------------------------------------------------
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void alloc(size_t count) {
        if (count <= n && p) return;
        release();
        if (count == 0) return;
        CUDA_CHECK(cudaMalloc(&p, count * sizeof(T)));
        n = count;
    }
    void zero(cudaStream_t s) {
        if (p) CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    }
    size_t bytes() const { return n * sizeof(T); }
};

// ---------------------------------------------------------------------------
// geometry of the 2-bit workspace
// ---------------------------------------------------------------------------
constexpr int SAMP_PAD = 256;   // samples per row are padded to a multiple of this
constexpr int SNP_PAD = 128;    // SNP rows are padded (all-missing) to a multiple of this

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// per-SNP statistics produced by the stats kernel
struct SnpStat {
    int32_t sum;   // sum of genotypes over non-missing samples
    int32_t num;   // number of non-missing samples
    int32_t n1;    // heterozygotes
    int32_t pad;
};

// Row window of the N x N output that the accumulators currently cover.  The default window
// is the whole matrix; for N where N^2 accumulators do not fit in HBM the host walks windows of
// 256-aligned rows and receives each window's slice of the row-packed upper triangle
// (rows r0..r0+rows are contiguous there, CdMatTri order src/dGenGWAS.h:556-561).
struct RowWin {
    long long r0;      // first row (multiple of 256)
    long long rows;    // window height in rows (multiple of 256)
    long long r1;      // min(r0 + rows, n_samp): end of the valid rows
    long long pbase;   // packed index of (r0, r0)
};
__host__ __device__ inline long long tri_idx(long long n, long long i, long long j) {
    return j + i * (2 * n - i - 1) / 2;
}

struct ReduceBuf {
    void *ptr;
    int64_t count;
    int kind;   // 0 int64, 1 uint32, 2 float64
    // ld > 0: the buffer is a stack of [rows][ld] planes of the row window starting at row0, and only
    // the columns >= (row0 + r) rounded down to 256 of row r are live (upper triangle, 256-tiles): an
    // in-library reduction (multi.cu) skips the dead half; a plain sum over `count` stays correct
    int64_t ld = 0, rows = 0, row0 = 0;
};


// project.cu
void pca_snp_loading(snprel_ctx *c, int k, const double *eigval, const double *eigvect, double trace_xtx,
                     int bayesian, double *loading, double *avgfreq, double *scale);
void pca_samp_loading(snprel_ctx *c, int k, const double *loadings, const double *avgfreq, const double *scale,
                      double *out);
void pca_corr(snprel_ctx *c, int k, const double *eigvect, double *out);
void eigmix_snp_loading(snprel_ctx *c, int k, const double *eigval, const double *eigvect, const double *afreq,
                        double *loading);
void eigmix_samp_loading(snprel_ctx *c, int k, const double *loadings, const double *afreq, double *out);
void pca_randomized(snprel_ctx *c, const double *aux_mat, int aux_dim, int iter_num, double *sigma, double *vt,
                    double *trace_xtx2);

}  // namespace snprel

// ---------------------------------------------------------------------------
// the context
// ---------------------------------------------------------------------------
struct snprel_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evs0 = nullptr, evs1 = nullptr;
    std::string err;
    int64_t launches = 0;
    uint32_t debug_flags = 0;
    int num_sms = 148;
    bool gram_attr_done = false;   // cudaFuncSetAttribute(max dynamic smem) done on this device

    // workspace
    int64_t n_samp = 0;        // N
    int64_t n_samp_pad = 0;    // N rounded up to SAMP_PAD
    int64_t row_bytes = 0;     // n_samp_pad / 4
    int64_t snp_cap = 0;       // allocated SNP rows (multiple of SNP_PAD)
    int64_t n_snp = 0;         // valid SNP rows
    snprel::DevBuf<uint8_t> geno2b;       // [snp_cap][row_bytes]
    snprel::DevBuf<snprel::SnpStat> stat; // [snp_cap]
    bool stat_valid = false;
    snprel::DevBuf<uint8_t> stage_u8;     // host->device staging for push_u8

    // bit planes for the packed-bit estimators: [n_word64][n_samp_pad] uint4
    snprel::DevBuf<uint4> planes;
    int64_t plane_words = 0;
    bool planes_valid = false;

    // accumulators
    int64_t win_r0 = 0, win_rows = 0;   // row window (win_rows == 0: whole matrix)
    int accum_est = -1;          // estimator the accumulators belong to
    bool accum_reduced = false;  // true after snprel_mark_reduced
    int64_t accum_win_r0 = 0;    // first row of the window the accumulators were built for
    int64_t accum_win_rows = 0;  // ... and its height
    int accum_bayesian = 0;      // covariance accumulators: the Bayesian-normalisation flag they were built with
    snprel::DevBuf<uint32_t> cnt;         // packed-bit counters [ncnt][npad][npad]
    int cnt_planes = 0;
    snprel::DevBuf<double> cnt_f64;       // KING-homo f64 pair sums [2][npad][npad]
    snprel::DevBuf<long long> acc;        // fixed-point Gram planes [nplane][npad][npad]
    int acc_planes = 0;
    snprel::DevBuf<long long> samp_sum;   // per-sample fixed-point sums [nvec][npad]
    int samp_vecs = 0;
    snprel::DevBuf<double> scalars;       // global f64 scalars (SumDenominator, ...)
    snprel::DevBuf<long long> iscalars;   // global int64 scalars (nLocus, ...)
    snprel_plan plan{};
    std::vector<snprel::ReduceBuf> reduce_list;
    // persistent scratch of the accumulate step (no cudaMalloc / cudaFree in the hot path)
    snprel::DevBuf<uint32_t> scr_tab;     // digit tables [npass][snp_cap]
    snprel::DevBuf<int> scr_flags;        // [0] digit overflow, [1] pipeline error
    snprel::DevBuf<double> scr_plan;      // plan statistics [3]
    snprel::DevBuf<uint8_t> scr_items, scr_passes;   // K1 work items / pass descriptors (slices, see gram_tc_run)
    snprel::DevBuf<long long> scr_trace;  // [items][8] clock stamps of the last K1 launch (debug flag 1)
    int64_t trace_items = 0;
    uint8_t *stage_host = nullptr;        // mapped pinned staging of the same slices
    size_t stage_bytes = 0, stage_used = 0;
    struct ConstTab {
        uint32_t word = 0;
        snprel::DevBuf<uint32_t> buf;
    };
    std::vector<std::unique_ptr<ConstTab>> const_tabs;   // constant per-SNP tables (gram_const_table)
    snprel::DevBuf<int> scr_cnt;          // per-sample heterozygote / missing counts [2][npad]
    snprel::DevBuf<long long> scr_ew;     // per-sample error weight sum_l |B_l[g_il]| [npad]
    snprel::DevBuf<long long> scr_sq;     // per-sample sum_l ceil(B_l[g_il]^2 / 127) [npad] (randomised-rounding bound)
    snprel::DevBuf<long long> scr_dg;     // per-sample diagonal bound, in units of c (diagtab_kernel) [npad]
    snprel::DevBuf<uint32_t> scr_tabf;    // [snp_cap]: ceil(w (g - mu)^2 / c) as bytes by genotype code
    snprel::DevBuf<int> scr_chunk;        // per GRAM_CHUNK SNPs: max over samples of the chunk's error weight
    snprel::DevBuf<int2> scr_coltab;      // per-SNP integer column table (s_l, t_l), B_l[g] = s_l g - t_l
    snprel::DevBuf<uint32_t> scr_tabb;    // [2][snp_cap]: B_l as int8 bytes by genotype code, and |B_l|
    uint64_t coltab_version = 0;
    int coltab_est = -1, coltab_bayesian = 0;
    std::vector<long long> host_ew;
    std::vector<int64_t> chunk_bound;     // host copy of scr_chunk (int32 accumulator headroom of K1)
    snprel::DevBuf<double> scr_part;      // per-block float64 partial sums of the tables kernel
    snprel::DevBuf<double> scr_out;       // device result of the last epilogue (kept across calls and row windows)
    snprel::DevBuf<double> scr_diag;      // diagonal staging of the trace
    std::vector<double> host_diag;
    snprel::DevBuf<uint32_t> scr_ctab;    // constant tables of the tensor count engine
    snprel::DevBuf<long long> scr_cacc;   // int64 planes of the tensor count engine (acc belongs to the covariance path)
    int count_engine = 0;                 // 0: packed-bit pair kernels (default), 1: tensor pipe
    int round_mode = 2;                   // 0: round to nearest + worst-case bound, 1: randomised + Hoeffding, 2 (default): the one with fewer passes
    int64_t snp_origin = 0;               // global index of the first SNP row: keys the rounding draws (snprel_set_snp_origin)
    std::vector<int> host_cnt;

    // window-invariant products of the covariance path, kept across row windows (tiled N x N
    // output: every window re-uses the plan statistics, digit tables and per-sample vectors).
    // geno_version changes whenever the resident 2-bit matrix (or anything derived) is dropped.
    uint64_t geno_version = 1;
    struct PlanCache {
        uint64_t version = 0;
        int est = -1, bayesian = 0;
        snprel_plan stats{};
    } plan_cache;
    struct PrepCache {
        uint64_t version = 0;
        int est = -1, bayesian = 0, f = 0, fw = 0, fd = 0, fv = 0, nU = 0, nW = 0, nD = 0, nD2 = 0, rounding = 0;
        int64_t origin = 0;     // SNP origin the (randomised) tables were drawn with
        bool reduced = false;   // the per-sample vectors / scalars already hold the all-reduced sums
    } prep_cache;

    // eigen step (eigen.cu): persistent cuBLAS / cuSOLVER handles and what the last solve did
    void *eig_handles = nullptr;
    void *ipc_peers = nullptr;
    // host-to-device copies still in flight (snprel_geno_push_2b_async): SNP rows [l0, l1) are valid once
    // `ev` has fired on copy_stream.  The covariance accumulate consumes them chunk by chunk (grm.cu);
    // every other entry point waits for all of them first (API_BEGIN).
    struct PendingCopy { int64_t l0, l1; cudaEvent_t ev; int64_t copied; bool consumed = false; };   // copied: bytes per row the host copy covered (-1: no padding to fix)
    std::vector<PendingCopy> pending;
    cudaStream_t copy_stream = nullptr;
    // asynchronous result delivery (snprel_set_async_output): the device-to-host copy of a finished row window
    // runs on out_stream while the next window accumulates; the result scratch (scr_out) is not rewritten
    // before output_wait
    int async_output = 0;
    cudaStream_t out_stream = nullptr;
    cudaEvent_t out_ready = nullptr, out_done = nullptr;
    bool out_pending = false;
    cudaEvent_t copy_ev0 = nullptr;
    double last_copy_ms = 0;     // first async copy chunk queued -> last one arrived
    int stream_ranges = 0;       // launches the last streamed accumulate was cut into
    int64_t streamed_steps = 0, stream_fallbacks = 0;   // accumulates that consumed in-flight copies / that had to be redone   // multi.cu: peers' reduce buffers mapped through CUDA IPC (one process per GPU)
    int eig_solver = 0;          // 0 dense (Xsyevd), 1 Chebyshev-filtered subspace iteration
    int eig_rounds = 0, eig_gemms = 0;
    double eig_phase_ms[3] = {0, 0, 0};   // filter, orthonormalisation, Rayleigh-Ritz

    // hot-kernel bookkeeping for bench.py
    double hot_ms = 0;
    int64_t hot_launches = 0;
    int64_t hot_items = 0;       // work items (CTA pairs) of the last table-Gram launch
    double hot_units = 0;
    double step_ms = 0;          // whole plan + accumulate (CUDA events)
    double plan_ms = 0;
    double finish_ms = 0;        // device epilogue of the last snprel_time_accumulate / snprel_time_finish
};

namespace snprel {

inline void count_launch(snprel_ctx *c, int64_t n = 1) { c->launches += n; }

// everything derived from the resident 2-bit matrix is stale
inline void drop_derived(snprel_ctx *c) {
    c->stat_valid = false;
    c->planes_valid = false;
    c->accum_est = -1;
    c->accum_reduced = false;
    c->geno_version++;
}

inline RowWin row_window(const snprel_ctx *c) {
    RowWin w;
    w.r0 = c->win_rows > 0 ? c->win_r0 : 0;
    w.rows = c->win_rows > 0 ? c->win_rows : c->n_samp_pad;
    if (w.r0 + w.rows > c->n_samp_pad) w.rows = c->n_samp_pad - w.r0;
    w.r1 = std::min<long long>(w.r0 + w.rows, c->n_samp);
    w.pbase = tri_idx(c->n_samp, w.r0, w.r0);
    return w;
}
inline bool full_window(const snprel_ctx *c) { return c->win_rows <= 0; }
// number of packed upper-triangle entries in the window's rows
inline size_t window_packed_count(const snprel_ctx *c) {
    RowWin w = row_window(c);
    return (size_t)(tri_idx(c->n_samp, w.r1 - 1, c->n_samp - 1) + 1 - w.pbase);
}

// ---- result delivery ----------------------------------------------------------------
inline void output_wait(snprel_ctx *c) {
    if (!c->out_pending) return;
    CUDA_CHECK(cudaEventSynchronize(c->out_done));
    c->out_pending = false;
}
// copy `count` elements of a finished result to the caller; `last`: no more pieces of this result follow
template <class T>
inline void deliver(snprel_ctx *c, T *host, const T *dev, size_t count, bool last = true) {
    if (!c->async_output) {
        CUDA_CHECK(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
        if (last) CUDA_CHECK(cudaStreamSynchronize(c->stream));
        return;
    }
    if (!c->out_stream) {
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&c->out_ready, cudaEventDisableTiming));
        CUDA_CHECK(cudaEventCreateWithFlags(&c->out_done, cudaEventDisableTiming));
    }
    CUDA_CHECK(cudaEventRecord(c->out_ready, c->stream));
    CUDA_CHECK(cudaStreamWaitEvent(c->out_stream, c->out_ready, 0));
    CUDA_CHECK(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, c->out_stream));
    if (last) {
        CUDA_CHECK(cudaEventRecord(c->out_done, c->out_stream));
        c->out_pending = true;
    }
}

#define KERNEL_CHECK(ctx)                                   \
    do {                                                    \
        CUDA_CHECK(cudaGetLastError());                     \
        ::snprel::count_launch(ctx);                        \
    } while (0)

// geno.cu
void geno_begin(snprel_ctx *c, int64_t n_samp, int64_t cap);
void geno_push_u8(snprel_ctx *c, const uint8_t *host, int64_t cnt);
void geno_push_2b(snprel_ctx *c, const uint8_t *host, int64_t cnt, int64_t row_bytes);
void geno_push_bitstream(snprel_ctx *c, const uint8_t *host, int64_t first_genotype, int64_t cnt);
void geno_synth(snprel_ctx *c, int64_t n_snp, uint64_t seed, double maf_lo, double maf_hi,
                double miss_rate, int64_t snp_start);
void geno_copy_u8(snprel_ctx *c, uint8_t *out);
void geno_copy_2b(snprel_ctx *c, uint8_t *out, int64_t row_bytes);
void geno_seek(snprel_ctx *c, int64_t snp_index);
void snp_stats_range(snprel_ctx *c, int64_t l0, int64_t rows);
void geno_push_2b_async(snprel_ctx *c, const uint8_t *host, int64_t cnt, int64_t row_bytes);
void geno_wait(snprel_ctx *c);
void geno_fix_chunk_padding(snprel_ctx *c, const snprel_ctx::PendingCopy &p);
constexpr int64_t STREAM_CHUNK = 131072;   // SNP rows per in-flight copy chunk = one K1 segment (1024 stages)
void geno_commit(snprel_ctx *c, int64_t n_snp);
void geno_pad_tail(snprel_ctx *c);
void ensure_stats(snprel_ctx *c);
void snp_ratefreq(snprel_ctx *c, double *af, double *maf, double *mr);
void select_snp_base(snprel_ctx *c, const double *afreq, int remove_mono, double maf, double missrate,
                     uint8_t *out_sel, int64_t *n_removed);
void ensure_planes(snprel_ctx *c);

// bitcount.cu
void bitcount_accumulate(snprel_ctx *c, int estimator);
void ibs_num_finish(snprel_ctx *c, int32_t *i0, int32_t *i1, int32_t *i2);
void ibs_ave_finish(snprel_ctx *c, double *out, int packed);
void ibd_mom_sums(snprel_ctx *c, const double *afreq_in, double *sums, double *afreq_out);
void ibd_mom_finish(snprel_ctx *c, const double *sums, int constraint, double *k0, double *k1, int packed);
void king_robust_finish(snprel_ctx *c, const int32_t *fam, double *ibs0, double *kin, int packed);
void king_robust_counts_finish(snprel_ctx *c, int32_t *out5);
void beta_counts_finish(snprel_ctx *c, int32_t *out2);
void indiv_beta_finish(snprel_ctx *c, int inbreeding, int grm_flavour, double *out, int packed,
                       double *avg_out);

// gram_tc.cu
constexpr int GRAM_CHUNK = 512;   // SNPs per chunk of the |B| bounds below
struct GramPass {
    const uint32_t *tabA;   // device [snp_cap]: 4 int8 digits per SNP (byte g = genotype code g)
    const uint32_t *tabB;   // device [snp_cap]: 4 int8 column values per SNP (gram_const_table for a constant table)
    int plane;              // output plane
    int shift;              // contribution = acc << shift
    int b_abs_max;          // max |tabB entry|: int32 accumulator headroom
    // optional, per GRAM_CHUNK SNPs: bound on sum_l |tabB_l[g_jl]| over the chunk for ANY sample j
    // (tighter than b_abs_max x GRAM_CHUNK; lets a long SNP range stay in one int32 accumulation)
    const std::vector<int64_t> *b_chunk_bound;
};
const uint32_t *gram_const_table(snprel_ctx *c, uint32_t word);
void gram_tc_run(snprel_ctx *c, const GramPass *passes, int npass, long long *out_planes,
                 bool upper_only, int64_t snp_lo = 0, int64_t snp_hi = -1, bool sync = true);
void gram_tc_check(snprel_ctx *c);

// grm.cu
void grm_plan_local(snprel_ctx *c, int est, snprel_plan *plan);
void grm_accumulate(snprel_ctx *c, int est, const snprel_plan *plan);
void grm_plan_format(int est, snprel_plan *plan, int round_mode, int64_t n_samp);   // host only
void grm_finish(snprel_ctx *c, int method, double *out, int packed);
void grm_finish_device(snprel_ctx *c, int est);
void pca_finish(snprel_ctx *c, int eigen_cnt, int bayesian, double *genmat, double *trace_xtx,
                double *trace_val, double *eigval, double *eigvec);
void eigmix_finish(snprel_ctx *c, int eigen_cnt, int diagadj, double *ibd, double *afreq,
                   double *eigval, double *eigvec);
void king_homo_finish(snprel_ctx *c, double *k0, double *k1, int packed);
void table_gram_debug(snprel_ctx *c, const int8_t *tabA, const int8_t *tabB, int64_t *out);
void tensor_count_accumulate(snprel_ctx *c, int est);

// eigen.cu
void eigen_topk(snprel_ctx *c, const double *m_upper, int64_t n, int k, double *eigval, double *eigvec);
void eigen_release(snprel_ctx *c);
// small dense helpers on the context's cuBLAS / cuSOLVER handles (column-major, float64)
void la_transpose(snprel_ctx *c, int64_t m, int64_t n, const double *A, int64_t lda, double *B, int64_t ldb);   // B [m x n] = A^T, A [n x m]
void la_orthonormalise(snprel_ctx *c, double *A, int64_t m, int n);   // A [m x n], m >= n, lda = m  <-  Q of its Householder QR
// thin SVD of A [m x n], m >= n, lda = m (destroyed): S -> host [n]; U [m x n] and / or VT [n x n] on the device (may be NULL)
void la_svd_tall(snprel_ctx *c, double *A, int64_t m, int n, double *S_host, double *U, double *VT);


// project.cu
void pca_snp_loading(snprel_ctx *c, int k, const double *eigval, const double *eigvect, double trace_xtx,
                     int bayesian, double *loading, double *avgfreq, double *scale);
void pca_samp_loading(snprel_ctx *c, int k, const double *loadings, const double *avgfreq, const double *scale,
                      double *out);
void pca_corr(snprel_ctx *c, int k, const double *eigvect, double *out);
void eigmix_snp_loading(snprel_ctx *c, int k, const double *eigval, const double *eigvect, const double *afreq,
                        double *loading);
void eigmix_samp_loading(snprel_ctx *c, int k, const double *loadings, const double *afreq, double *out);
void pca_randomized(snprel_ctx *c, const double *aux_mat, int aux_dim, int iter_num, double *sigma, double *vt,
                    double *trace_xtx2);

}  // namespace snprel
