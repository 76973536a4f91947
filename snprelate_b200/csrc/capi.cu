// extern "C" surface of libsnprel_b200.so (see include/snprel_b200.h).
// Every entry point catches snprel::Error and returns a code; messages are kept
// per context (and in a process-wide slot for failures before a context exists).
#include <cstdlib>
#include "common.cuh"

using namespace snprel;

static std::string g_create_error;

#define API_BEGIN(ctx)                                       \
    if (!(ctx)) {                                            \
        g_create_error = "NULL snprel_ctx";                  \
        return 1;                                            \
    }                                                        \
    try {                                                    \
        cudaError_t _se = cudaSetDevice((ctx)->device);      \
        if (_se != cudaSuccess) fail("cudaSetDevice(%d): %s", (ctx)->device, cudaGetErrorString(_se)); \
        geno_wait(ctx);

/* entry points that may consume host-to-device copies still in flight chunk by chunk (grm.cu) */
#define API_BEGIN_STREAMING(ctx)                             \
    if (!(ctx)) {                                            \
        g_create_error = "NULL snprel_ctx";                  \
        return 1;                                            \
    }                                                        \
    try {                                                    \
        cudaError_t _se = cudaSetDevice((ctx)->device);      \
        if (_se != cudaSuccess) fail("cudaSetDevice(%d): %s", (ctx)->device, cudaGetErrorString(_se));

#define API_END(ctx)                                         \
    return 0;                                                \
    }                                                        \
    catch (const Error &e) {                                 \
        (ctx)->err = e.msg;                                  \
        return 1;                                            \
    }                                                        \
    catch (const std::exception &e) {                        \
        (ctx)->err = e.what();                               \
        return 2;                                            \
    }

extern "C" {

int64_t snprel_abi_sizeof_plan(void) { return (int64_t)sizeof(snprel_plan); }

const char *snprel_version(void) { return "snprel_b200 0.2 (sm_100a)"; }

int snprel_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

int snprel_create(snprel_ctx **out, int device) {
    if (!out) {
        g_create_error = "snprel_create: NULL output pointer";
        return 1;
    }
    *out = nullptr;
    // SNPREL_ROUNDING = nearest | random | auto: initial mode of snprel_set_rounding (default auto)
    int round_mode = 2;
    if (const char *env = getenv("SNPREL_ROUNDING")) {
        const std::string v(env);
        if (v == "nearest") round_mode = 0;
        else if (v == "random") round_mode = 1;
        else if (!(v == "auto" || v.empty())) {
            g_create_error = "SNPREL_ROUNDING: expected \"nearest\", \"random\" or \"auto\", got \"" + v + "\"";
            return 1;
        }
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        // no CPU fallback: the product path needs a CUDA device
        g_create_error = std::string("snprel_create: no CUDA device available (") +
                         (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") + ")";
        return 3;
    }
    if (device < 0 || device >= ndev) {
        g_create_error = "snprel_create: device index out of range";
        return 1;
    }
    snprel_ctx *c = new snprel_ctx();
    c->device = device;
    c->round_mode = round_mode;
    try {
        CUDA_CHECK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10)
            fail("snprel_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                 prop.major, prop.minor);
        c->num_sms = prop.multiProcessorCount;
        CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreate(&c->ev0));
        CUDA_CHECK(cudaEventCreate(&c->ev1));
        CUDA_CHECK(cudaEventCreate(&c->evs0));
        CUDA_CHECK(cudaEventCreate(&c->evs1));
    } catch (const Error &err) {
        g_create_error = err.msg;
        delete c;
        return 1;
    }
    *out = c;
    return 0;
}

void snprel_destroy(snprel_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->out_stream) {
        cudaStreamSynchronize(c->out_stream);
        cudaEventDestroy(c->out_ready);
        cudaEventDestroy(c->out_done);
        cudaStreamDestroy(c->out_stream);
    }
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        for (auto &p : c->pending) cudaEventDestroy(p.ev);
        if (c->copy_ev0) cudaEventDestroy(c->copy_ev0);
        cudaStreamDestroy(c->copy_stream);
    }
    eigen_release(c);
    snprel_peer_reduce_close(c);
    if (c->stage_host) cudaFreeHost(c->stage_host);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->evs0) cudaEventDestroy(c->evs0);
    if (c->evs1) cudaEventDestroy(c->evs1);
    cudaStream_t s = c->stream;
    delete c;   // frees device buffers
    if (s) cudaStreamDestroy(s);
}

const char *snprel_last_error(snprel_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int snprel_geno_begin(snprel_ctx *c, int64_t n_samp, int64_t cap) {
    API_BEGIN(c) geno_begin(c, n_samp, cap);
    c->snp_origin = 0;
    API_END(c)
}
int snprel_geno_push_u8(snprel_ctx *c, const uint8_t *geno, int64_t cnt) {
    API_BEGIN(c) geno_push_u8(c, geno, cnt);
    API_END(c)
}
int snprel_geno_push_2b(snprel_ctx *c, const uint8_t *packed, int64_t cnt, int64_t row_bytes) {
    API_BEGIN(c) geno_push_2b(c, packed, cnt, row_bytes);
    API_END(c)
}
int snprel_geno_push_bitstream(snprel_ctx *c, const uint8_t *stream, int64_t first_genotype, int64_t cnt) {
    API_BEGIN(c) geno_push_bitstream(c, stream, first_genotype, cnt);
    API_END(c)
}
int snprel_geno_push_2b_async(snprel_ctx *c, const uint8_t *packed, int64_t cnt, int64_t row_bytes) {
    API_BEGIN_STREAMING(c) geno_push_2b_async(c, packed, cnt, row_bytes);
    API_END(c)
}
int snprel_geno_wait(snprel_ctx *c) {
    API_BEGIN(c)
    API_END(c)
}
int snprel_set_async_output(snprel_ctx *c, int on) {
    API_BEGIN(c)
    output_wait(c);
    c->async_output = on ? 1 : 0;
    API_END(c)
}
int snprel_output_wait(snprel_ctx *c) {
    API_BEGIN_STREAMING(c)
    output_wait(c);
    API_END(c)
}
int snprel_stream_stats(snprel_ctx *c, int64_t *streamed, int64_t *fallbacks) {
    API_BEGIN_STREAMING(c)
    if (streamed) *streamed = c->streamed_steps;
    if (fallbacks) *fallbacks = c->stream_fallbacks;
    API_END(c)
}
// clock stamps of the last K1 launch made with snprel_debug_flags(ctx, 1): out[items][8] (entry, set-up done,
// first MMA, last commit issued, MMAs complete, epilogue done, exit, stages); returns the item count in *items
int snprel_k1_trace(snprel_ctx *c, int64_t *out, int64_t capacity_items, int64_t *items) {
    API_BEGIN(c)
    if (items) *items = c->trace_items;
    if (out && c->scr_trace.p && c->trace_items > 0) {
        const int64_t n = std::min<int64_t>(capacity_items, c->trace_items);
        CUDA_CHECK(cudaMemcpy(out, c->scr_trace.p, (size_t)n * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
    }
    API_END(c)
}
int snprel_stream_last_copy_ms(snprel_ctx *c, double *ms) {
    API_BEGIN_STREAMING(c)
    if (ms) *ms = c->last_copy_ms;
    API_END(c)
}
int snprel_geno_seek(snprel_ctx *c, int64_t snp_index) {
    API_BEGIN(c) geno_seek(c, snp_index);
    API_END(c)
}
int snprel_geno_commit(snprel_ctx *c, int64_t n_snp) {
    API_BEGIN(c) geno_commit(c, n_snp);
    API_END(c)
}
int snprel_geno_device_rows(snprel_ctx *c, void **dev_ptr, int64_t *row_bytes, int64_t *capacity) {
    API_BEGIN(c)
    if (c->n_samp <= 0) fail("snprel_geno_device_rows: no genotype workspace");
    if (dev_ptr) *dev_ptr = c->geno2b.p;
    if (row_bytes) *row_bytes = c->row_bytes;
    if (capacity) *capacity = c->snp_cap;
    API_END(c)
}
int snprel_geno_synth(snprel_ctx *c, int64_t n_snp, uint64_t seed, double maf_lo, double maf_hi,
                      double miss_rate, int64_t snp_start) {
    API_BEGIN(c) geno_synth(c, n_snp, seed, maf_lo, maf_hi, miss_rate, snp_start);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    API_END(c)
}
int snprel_geno_dim(snprel_ctx *c, int64_t *n_samp, int64_t *n_snp) {
    API_BEGIN_STREAMING(c)      // (reads two integers: copies in flight are not waited for)
    if (n_samp) *n_samp = c->n_samp;
    if (n_snp) *n_snp = c->n_snp;
    API_END(c)
}
int snprel_geno_copy_u8(snprel_ctx *c, uint8_t *out) {
    API_BEGIN(c) geno_copy_u8(c, out);
    API_END(c)
}
int snprel_geno_copy_2b(snprel_ctx *c, uint8_t *out, int64_t row_bytes) {
    API_BEGIN(c) geno_copy_2b(c, out, row_bytes);
    API_END(c)
}
int snprel_snp_ratefreq(snprel_ctx *c, double *af, double *maf, double *mr) {
    API_BEGIN(c) snp_ratefreq(c, af, maf, mr);
    API_END(c)
}
int snprel_select_snp_base(snprel_ctx *c, int remove_mono, double maf, double missrate,
                           uint8_t *out_sel, int64_t *n_removed) {
    API_BEGIN(c) select_snp_base(c, nullptr, remove_mono, maf, missrate, out_sel, n_removed);
    API_END(c)
}
int snprel_select_snp_base_ex(snprel_ctx *c, const double *afreq, int remove_mono, double maf, double missrate,
                              uint8_t *out_sel, int64_t *n_removed) {
    API_BEGIN(c)
    if (!afreq) fail("snprel_select_snp_base_ex: NULL allele frequencies");
    select_snp_base(c, afreq, remove_mono, maf, missrate, out_sel, n_removed);
    API_END(c)
}

// ---- packed-bit estimators -------------------------------------------------
int snprel_ibs_num(snprel_ctx *c, int32_t *i0, int32_t *i1, int32_t *i2) {
    API_BEGIN(c) ibs_num_finish(c, i0, i1, i2);
    API_END(c)
}
int snprel_ibs_ave(snprel_ctx *c, double *out, int packed) {
    API_BEGIN(c) ibs_ave_finish(c, out, packed);
    API_END(c)
}
int snprel_ibd_mom_sums(snprel_ctx *c, const double *afreq_in, double *sums6, double *afreq_out) {
    API_BEGIN(c) ibd_mom_sums(c, afreq_in, sums6, afreq_out);
    API_END(c)
}
int snprel_ibd_mom_from_sums(snprel_ctx *c, const double *sums6, int constraint, int packed, double *k0,
                             double *k1) {
    API_BEGIN(c) ibd_mom_finish(c, sums6, constraint, k0, k1, packed);
    API_END(c)
}
int snprel_ibd_mom(snprel_ctx *c, const double *afreq_in, int constraint, int packed, double *k0, double *k1,
                   double *afreq_out) {
    API_BEGIN(c)
    double sums[6];
    ibd_mom_sums(c, afreq_in, sums, afreq_out);
    ibd_mom_finish(c, sums, constraint, k0, k1, packed);
    API_END(c)
}
int snprel_king_robust(snprel_ctx *c, const int32_t *fam, double *ibs0, double *kin, int packed) {
    API_BEGIN(c) king_robust_finish(c, fam, ibs0, kin, packed);
    API_END(c)
}
int snprel_king_robust_counts(snprel_ctx *c, int32_t *out5) {
    API_BEGIN(c) king_robust_counts_finish(c, out5);
    API_END(c)
}
int snprel_king_homo(snprel_ctx *c, double *k0, double *k1, int packed) {
    API_BEGIN(c) king_homo_finish(c, k0, k1, packed);
    API_END(c)
}
int snprel_indiv_beta(snprel_ctx *c, int inbreeding, double *out, int packed, double *avg_out) {
    API_BEGIN(c)
    if (inbreeding != 0 && inbreeding != 1) fail("'inbreeding' must be TRUE or FALSE.");   // src/genBeta.cpp:365-366
    indiv_beta_finish(c, inbreeding, 0, out, packed, avg_out);
    API_END(c)
}
int snprel_indiv_beta_counts(snprel_ctx *c, int32_t *out2) {
    API_BEGIN(c) beta_counts_finish(c, out2);
    API_END(c)
}

// ---- covariance-type estimators --------------------------------------------
int snprel_grm(snprel_ctx *c, int method, double *out, int packed, double *avg_out) {
    API_BEGIN_STREAMING(c)
    if (method == SNPREL_GRM_INDIVBETA) geno_wait(c);
    switch (method) {
        case SNPREL_GRM_EIGENSTRAT:
        case SNPREL_GRM_GCTA:
        case SNPREL_GRM_CORR:
        case SNPREL_GRM_EIGMIX: grm_finish(c, method, out, packed); break;
        case SNPREL_GRM_INDIVBETA: indiv_beta_finish(c, 1, 1, out, packed, avg_out); break;
        default: fail("Invalid 'method'!");   // src/genPCA.cpp:1709
    }
    API_END(c)
}
int snprel_pca(snprel_ctx *c, int eigen_cnt, int bayesian, double *genmat, double *trace_xtx,
               double *trace_val, double *eigval, double *eigvec) {
    API_BEGIN_STREAMING(c) pca_finish(c, eigen_cnt, bayesian, genmat, trace_xtx, trace_val, eigval, eigvec);
    API_END(c)
}
int snprel_eigmix(snprel_ctx *c, int eigen_cnt, int diagadj, double *ibd, double *afreq, double *eigval,
                  double *eigvec) {
    API_BEGIN_STREAMING(c)
    if (diagadj != 0 && diagadj != 1) fail("'diagadj' must be TRUE or FALSE.");   // src/genEIGMIX.cpp:661-662
    eigmix_finish(c, eigen_cnt, diagadj, ibd, afreq, eigval, eigvec);
    API_END(c)
}

// ---- loadings / projection / SNP-PC correlation -------------------------------
int snprel_pca_snp_loading(snprel_ctx *c, int k, const double *eigval, const double *eigvect, double trace_xtx,
                           int bayesian, double *loading, double *avgfreq, double *scale) {
    API_BEGIN(c) pca_snp_loading(c, k, eigval, eigvect, trace_xtx, bayesian, loading, avgfreq, scale);
    API_END(c)
}
int snprel_pca_samp_loading(snprel_ctx *c, int k, const double *loadings, const double *avgfreq, const double *scale,
                            double *out) {
    API_BEGIN(c) pca_samp_loading(c, k, loadings, avgfreq, scale, out);
    API_END(c)
}
int snprel_pca_corr(snprel_ctx *c, int k, const double *eigvect, double *out) {
    API_BEGIN(c) pca_corr(c, k, eigvect, out);
    API_END(c)
}
int snprel_pca_randomized(snprel_ctx *c, const double *aux_mat, int aux_dim, int iter_num, double *sigma, double *vt,
                          double *trace_xtx2) {
    API_BEGIN(c) pca_randomized(c, aux_mat, aux_dim, iter_num, sigma, vt, trace_xtx2);
    API_END(c)
}
int snprel_eigmix_snp_loading(snprel_ctx *c, int k, const double *eigval, const double *eigvect, const double *afreq,
                              double *loading) {
    API_BEGIN(c) eigmix_snp_loading(c, k, eigval, eigvect, afreq, loading);
    API_END(c)
}
int snprel_eigmix_samp_loading(snprel_ctx *c, int k, const double *loadings, const double *afreq, double *out) {
    API_BEGIN(c) eigmix_samp_loading(c, k, loadings, afreq, out);
    API_END(c)
}

// ---- split accumulate / reduce / finish -------------------------------------
static bool is_cov(int est) { return est >= SNPREL_GRM_EIGENSTRAT && est <= SNPREL_GRM_EIGMIX; }

static void timed_plan(snprel_ctx *c, int est, snprel_plan *plan) {
    CUDA_CHECK(cudaEventRecord(c->evs0, c->stream));
    if (is_cov(est)) {
        grm_plan_local(c, est == SNPREL_GRM_CORR ? SNPREL_GRM_GCTA : est, plan);
    } else {
        plan->max_abs = 0;
        plan->max_abs_w = 0;
        plan->sum_bound = 0;
        plan->err_weight = 0;
        plan->scale = 0;
        plan->total_missing = 0;
        plan->max_missing = 0;
        plan->n_snp = c->n_snp;
    }
    CUDA_CHECK(cudaEventRecord(c->evs1, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, c->evs0, c->evs1));
    c->plan_ms = ms;
}
static void timed_accumulate(snprel_ctx *c, int est, const snprel_plan *plan) {
    CUDA_CHECK(cudaEventRecord(c->evs0, c->stream));
    if (is_cov(est)) {
        if (!plan) fail("snprel_accumulate: covariance estimators need a plan");
        grm_accumulate(c, est, plan);
    } else if (est == SNPREL_EST_IBS || est == SNPREL_EST_KING_ROBUST || est == SNPREL_EST_BETA) {
        bitcount_accumulate(c, est);
    } else {
        fail("snprel_accumulate: unsupported estimator %d", est);
    }
    CUDA_CHECK(cudaEventRecord(c->evs1, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, c->evs0, c->evs1));
    c->step_ms = c->plan_ms + ms;
    c->plan_ms = 0;
}

// device epilogue (int64 planes -> float64 result in c->scr_out); adds its time to step_ms
static void timed_finish(snprel_ctx *c, int est) {
    CUDA_CHECK(cudaEventRecord(c->evs0, c->stream));
    grm_finish_device(c, est);
    CUDA_CHECK(cudaEventRecord(c->evs1, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, c->evs0, c->evs1));
    c->finish_ms = ms;
    c->step_ms += ms;
}

int snprel_time_finish(snprel_ctx *c, int est, double *ms) {
    API_BEGIN(c)
    if (!is_cov(est)) fail("snprel_time_finish: covariance estimators only");
    timed_finish(c, est);
    if (ms) *ms = c->finish_ms;
    API_END(c)
}
int snprel_plan_local(snprel_ctx *c, int est, snprel_plan *plan) {
    API_BEGIN(c)
    if (!plan) fail("snprel_plan_local: NULL plan");
    timed_plan(c, est, plan);
    API_END(c)
}
int snprel_accumulate(snprel_ctx *c, int est, const snprel_plan *plan) {
    API_BEGIN(c) timed_accumulate(c, est, plan);
    API_END(c)
}
int snprel_last_plan(snprel_ctx *c, snprel_plan *plan) {
    API_BEGIN(c)
    if (!plan) fail("snprel_last_plan: NULL plan");
    *plan = c->plan;
    API_END(c)
}
int snprel_last_step_ms(snprel_ctx *c, double *ms) {
    API_BEGIN(c)
    if (ms) *ms = c->step_ms;
    API_END(c)
}
int snprel_invalidate(snprel_ctx *c) {
    API_BEGIN(c)
    drop_derived(c);
    API_END(c)
}
int snprel_reduce_buffer_count(snprel_ctx *c) { return c ? (int)c->reduce_list.size() : 0; }
int snprel_reduce_buffer(snprel_ctx *c, int idx, void **dev_ptr, int64_t *count, int *kind) {
    API_BEGIN(c)
    if (idx < 0 || idx >= (int)c->reduce_list.size()) fail("snprel_reduce_buffer: index out of range");
    if (dev_ptr) *dev_ptr = c->reduce_list[idx].ptr;
    if (count) *count = c->reduce_list[idx].count;
    if (kind) *kind = c->reduce_list[idx].kind;
    API_END(c)
}
int snprel_mark_reduced(snprel_ctx *c) {
    API_BEGIN(c)
    if (c->accum_est < 0) fail("snprel_mark_reduced: nothing accumulated");
    c->accum_reduced = true;
    // the per-sample vectors and scalars were summed in place together with the planes
    if (c->prep_cache.version == c->geno_version) c->prep_cache.reduced = true;
    API_END(c)
}

int snprel_set_row_window(snprel_ctx *c, int64_t row0, int64_t rows) {
    API_BEGIN(c)
    if (c->n_samp <= 0) fail("snprel_set_row_window: no genotype workspace");
    if (rows < 0 || row0 < 0 || (row0 % 256) || (rows % 256)) fail("snprel_set_row_window: row0 and rows must be non-negative multiples of 256");
    if (rows > 0 && row0 >= c->n_samp) fail("snprel_set_row_window: window starts past the last sample");
    c->win_r0 = rows > 0 ? row0 : 0;
    c->win_rows = rows;
    c->accum_est = -1;
    c->accum_reduced = false;
    API_END(c)
}
int snprel_mem_info(snprel_ctx *c, int64_t *free_bytes, int64_t *total_bytes) {
    API_BEGIN_STREAMING(c)
    size_t f = 0, t = 0;
    CUDA_CHECK(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    API_END(c)
}
int snprel_window_count(snprel_ctx *c, int64_t *count) {
    API_BEGIN_STREAMING(c)
    if (c->n_samp <= 0) fail("snprel_window_count: no genotype workspace");
    if (count) *count = (int64_t)window_packed_count(c);
    API_END(c)
}

// ---- introspection -----------------------------------------------------------
int64_t snprel_kernel_launches(snprel_ctx *c) { return c ? c->launches : 0; }
int snprel_last_hot_kernel(snprel_ctx *c, double *ms, int64_t *launches, double *units) {
    API_BEGIN(c)
    if (ms) *ms = c->hot_ms;
    if (launches) *launches = c->hot_launches;
    if (units) *units = c->hot_units;
    API_END(c)
}
int snprel_time_accumulate(snprel_ctx *c, int est, int reps, double *ms) {
    API_BEGIN(c)
    if (reps <= 0) fail("snprel_time_accumulate: reps must be positive");
    double total = 0;
    for (int r = 0; r < reps; r++) {
        // everything derived from the resident 2-bit matrix is recomputed inside the timed region
        drop_derived(c);
        snprel_plan plan{};
        plan.frac_bits = -1;
        plan.frac_bits_w = -1;
        plan.frac_bits_d = -1;
        timed_plan(c, est, &plan);
        timed_accumulate(c, est, &plan);
        if (is_cov(est)) timed_finish(c, est);   // ... up to "final N x N complete" 
        total += c->step_ms;
    }
    if (ms) *ms = total / reps;
    API_END(c)
}
int snprel_table_gram(snprel_ctx *c, const int8_t *tabA, const int8_t *tabB, int64_t *out) {
    API_BEGIN(c) table_gram_debug(c, tabA, tabB, out);
    API_END(c)
}
int snprel_set_rounding(snprel_ctx *c, int mode) {
    API_BEGIN(c)
    if (mode < 0 || mode > 2)
        fail("snprel_set_rounding: 0 (round to nearest, worst-case bound), 1 (randomised, Hoeffding bound) or 2 (the one with fewer tensor passes)");
    c->round_mode = mode;
    c->accum_est = -1;
    c->accum_reduced = false;
    API_END(c)
}
int snprel_set_snp_origin(snprel_ctx *c, int64_t origin) {
    API_BEGIN_STREAMING(c)      // (does not wait for copies in flight)
    if (origin < 0) fail("snprel_set_snp_origin: negative origin");
    if (origin != c->snp_origin) {
        c->snp_origin = origin;
        c->accum_est = -1;        // (tables drawn with another origin are rebuilt: the prep cache compares it)
        c->accum_reduced = false;
    }
    API_END(c)
}
int snprel_plan_format(int est, snprel_plan *plan, int mode, int64_t n_samp) {
    try {
        grm_plan_format(est, plan, mode, n_samp);
        return plan->digits + plan->digits_w + plan->digits_d;
    } catch (const Error &e) {
        g_create_error = e.msg;
    } catch (const std::exception &e) {
        g_create_error = e.what();
    }
    return -1;
}
int snprel_last_eigen_info(snprel_ctx *c, int *solver, int *rounds, int *block_gemms, double *phase_ms) {
    API_BEGIN(c)
    if (solver) *solver = c->eig_solver;
    if (rounds) *rounds = c->eig_rounds;
    if (block_gemms) *block_gemms = c->eig_gemms;
    if (phase_ms)
        for (int i = 0; i < 3; i++) phase_ms[i] = c->eig_phase_ms[i];
    API_END(c)
}
int snprel_set_count_engine(snprel_ctx *c, int engine) {
    API_BEGIN(c)
    if (engine != 0 && engine != 1) fail("snprel_set_count_engine: 0 (packed-bit kernels) or 1 (tensor pipe)");
    c->count_engine = engine;
    c->accum_est = -1;
    c->accum_reduced = false;
    API_END(c)
}
int snprel_debug_flags(snprel_ctx *c, uint32_t flags) {
    API_BEGIN(c) c->debug_flags = flags;
    API_END(c)
}

}  // extern "C"
