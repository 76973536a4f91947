// Tall-skinny products of the normalised genotype matrix with a few dense vectors -- the
// steps either side of the eigen-decomposition:
//   SNP loadings        CPCA_SNPLoad      src/genPCA.cpp:938-1040,  gnrPCASNPLoading   :1489-1540
//   sample loadings     CPCA_SampleLoad   src/genPCA.cpp:1042-1123, gnrPCASampLoading  :1542-1563
//   SNP-PC correlation  CPCA_SNPCorr      src/genPCA.cpp:809-936,   gnrPCACorr         :1456-1485
//   EIGMIX flavours     CEigMix_SNPLoad / CEigMix_SampleLoad  src/genEIGMIX.cpp:445-640,
//                       gnrEigMixSNPLoading :739-775, gnrEigMixSampLoading :777-803
//
// Both directions are one float64 GEMM core whose A operand never exists in memory: a stage of
// z = (g - a_l) * b_l values (missing -> exactly 0, as in the reference) is expanded from the
// resident 2-bit matrix into shared memory and multiplied with a 32-column panel of the dense
// operand, 8 x 4 accumulators per thread.
//   snp_project :  out[l][k] = sum_j z_jl V[j][k]      (rows = SNPs, reduction over samples)
//   samp_project:  out[k][i] = sum_l z_il L[l][k]      (rows = samples, reduction over SNPs,
//                  SNP range split over blockIdx.z, partials summed in a fixed order)
// O(N M k) float64 FMAs on the CUDA cores (3.2e11 for 10k x 1M x 32: tens of ms); HBM sees the
// 2-bit matrix once per 32 vectors.  The summation order differs from the reference's serial
// loops, so results agree to float64 rounding (1e-13 relative in the tests), not bitwise.
#include "common.cuh"

#include <cmath>

namespace snprel {

constexpr int PJ_ROWS = 128;   // tile rows = threads per block
constexpr int PJ_STEP = 32;    // reduction steps per shared-memory stage
constexpr int PJ_COLS = 32;    // dense columns per block

// z table of one SNP; mode 0: (g - a) * b (two roundings, as `(*pGeno - avg) * scale`), mode 1: valid indicator
__device__ __forceinline__ void z_table(int mode, double2 ab, double (&tab)[3]) {
#pragma unroll
    for (int g = 0; g < 3; g++)
        tab[g] = mode == 0 ? __dmul_rn(__dsub_rn((double)g, ab.x), ab.y) : 1.0;
}

// acc[r][c] += As[s][ty*8 + r] * Bs[s][tx*4 + c] over the stage
__device__ __forceinline__ void stage_fma(const double (&As)[PJ_STEP][PJ_ROWS], const double (&Bs)[PJ_STEP][PJ_COLS],
                                          int ty, int tx, double (&acc)[8][4]) {
#pragma unroll 4
    for (int s = 0; s < PJ_STEP; s++) {
        double a[8], b[4];
        const double2 *ap = reinterpret_cast<const double2 *>(&As[s][ty * 8]);
        const double2 *bp = reinterpret_cast<const double2 *>(&Bs[s][tx * 4]);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            double2 v = ap[q];
            a[2 * q] = v.x;
            a[2 * q + 1] = v.y;
        }
#pragma unroll
        for (int q = 0; q < 2; q++) {
            double2 v = bp[q];
            b[2 * q] = v.x;
            b[2 * q + 1] = v.y;
        }
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int cc = 0; cc < 4; cc++) acc[r][cc] = fma(a[r], b[cc], acc[r][cc]);
    }
}

// out[l][k] = sum_j z_jl Vt[j][k];  Vt is [npad][kp] (rows >= n_samp and columns >= k are zero)
__global__ void __launch_bounds__(PJ_ROWS)
snp_project_kernel(const uint8_t *__restrict__ geno, int64_t row_bytes, int64_t n_snp, int64_t npad, int mode,
                   const double2 *__restrict__ ab, const double *__restrict__ Vt, int kp, int k_out,
                   double *__restrict__ out, int64_t ld_out) {
    __shared__ __align__(16) double As[PJ_STEP][PJ_ROWS];
    __shared__ __align__(16) double Bs[PJ_STEP][PJ_COLS];
    const int t = threadIdx.x, ty = t >> 3, tx = t & 7;
    const int64_t l0 = (int64_t)blockIdx.x * PJ_ROWS, l = l0 + t;
    const int k0 = blockIdx.y * PJ_COLS;
    const bool live = l < n_snp;
    double tab[3] = {0, 0, 0};
    if (live) z_table(mode, mode == 0 ? ab[l] : make_double2(0, 0), tab);
    const uint2 *grow = reinterpret_cast<const uint2 *>(geno + (live ? l : 0) * row_bytes);
    double acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int cc = 0; cc < 4; cc++) acc[r][cc] = 0.0;

    for (int64_t j0 = 0; j0 < npad; j0 += PJ_STEP) {
        const uint2 w = live ? __ldg(grow + (j0 >> 5)) : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
#pragma unroll
        for (int e = 0; e < PJ_STEP * PJ_COLS / PJ_ROWS; e++) {
            const int idx = e * PJ_ROWS + t, s = idx >> 5, kk = idx & 31;
            Bs[s][kk] = Vt[(j0 + s) * kp + k0 + kk];
        }
#pragma unroll
        for (int s = 0; s < 16; s++) {
            const uint32_t c0 = (w.x >> (2 * s)) & 3u, c1 = (w.y >> (2 * s)) & 3u;
            As[s][t] = c0 == 0 ? tab[0] : c0 == 1 ? tab[1] : c0 == 2 ? tab[2] : 0.0;
            As[16 + s][t] = c1 == 0 ? tab[0] : c1 == 1 ? tab[1] : c1 == 2 ? tab[2] : 0.0;
        }
        __syncthreads();
        stage_fma(As, Bs, ty, tx, acc);
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const int64_t lr = l0 + ty * 8 + r;
        if (lr >= n_snp) continue;
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
            const int kk = k0 + tx * 4 + cc;
            if (kk < k_out) out[lr * ld_out + kk] = acc[r][cc];
        }
    }
}

// part[z][k][i] = sum over the z-th SNP range of z_il Lt[l][k];  Lt is [round_up(n_snp,32)][kp], zero padded
__global__ void __launch_bounds__(PJ_ROWS)
samp_project_kernel(const uint8_t *__restrict__ geno, int64_t row_bytes, int64_t n_snp, int64_t npad,
                    const double2 *__restrict__ ab, const double *__restrict__ Lt, int64_t kp, int64_t snps_per_split,
                    double *__restrict__ part) {
    __shared__ __align__(16) double As[PJ_STEP][PJ_ROWS];
    __shared__ __align__(16) double Bs[PJ_STEP][PJ_COLS];
    __shared__ double Ts[PJ_STEP][4];
    const int t = threadIdx.x, ty = t >> 3, tx = t & 7;
    const int64_t i0 = (int64_t)blockIdx.x * PJ_ROWS;
    const int k0 = blockIdx.y * PJ_COLS;
    const int64_t lb = (int64_t)blockIdx.z * snps_per_split;
    const int64_t le = min(n_snp, lb + snps_per_split);
    const int64_t byte = (i0 + t) >> 2;
    const int sh = 2 * (int)((i0 + t) & 3);
    double acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int cc = 0; cc < 4; cc++) acc[r][cc] = 0.0;

    for (int64_t l0 = lb; l0 < le; l0 += PJ_STEP) {
        if (t < PJ_STEP) {
            double tab[3] = {0, 0, 0};
            if (l0 + t < le) z_table(0, ab[l0 + t], tab);
            Ts[t][0] = tab[0];
            Ts[t][1] = tab[1];
            Ts[t][2] = tab[2];
            Ts[t][3] = 0.0;
        }
#pragma unroll
        for (int e = 0; e < PJ_STEP * PJ_COLS / PJ_ROWS; e++) {
            const int idx = e * PJ_ROWS + t, s = idx >> 5, kk = idx & 31;
            Bs[s][kk] = (l0 + s < le) ? Lt[(l0 + s) * kp + k0 + kk] : 0.0;
        }
        uint32_t code[PJ_STEP];
#pragma unroll
        for (int s = 0; s < PJ_STEP; s++)   // 32 independent byte loads in flight
            code[s] = (l0 + s < le) ? ((uint32_t)__ldg(geno + (l0 + s) * row_bytes + byte) >> sh) & 3u : 3u;
        __syncthreads();
#pragma unroll
        for (int s = 0; s < PJ_STEP; s++) As[s][t] = Ts[s][code[s]];
        __syncthreads();
        stage_fma(As, Bs, ty, tx, acc);
        __syncthreads();
    }
    const int64_t kp_total = (int64_t)gridDim.y * PJ_COLS;
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int cc = 0; cc < 4; cc++)
            part[((int64_t)blockIdx.z * kp_total + k0 + tx * 4 + cc) * npad + i0 + ty * 8 + r] = acc[r][cc];
}

// out[k][i] = sum_z part[z][k][i], z ascending (run-to-run deterministic)
__global__ void samp_reduce_kernel(const double *__restrict__ part, int splits, int64_t kp_total, int64_t npad,
                                   int k_out, int64_t n, double *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int kk = blockIdx.y;
    if (i >= n || kk >= k_out) return;
    double s = 0;
    for (int z = 0; z < splits; z++) s += part[((int64_t)z * kp_total + kk) * npad + i];
    out[(int64_t)kk * n + i] = s;
}

// avg / scale of CPCA_SNPLoad::thread_loading (src/genPCA.cpp:955-972) from the per-SNP counts
__global__ void load_scale_kernel(const SnpStat *__restrict__ st, int64_t n_snp, int bayesian,
                                  double2 *__restrict__ ab) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_snp) return;
    const SnpStat s = st[l];
    double avg = 0, scale = 0;
    if (s.num > 0) {
        avg = (double)s.sum / (double)s.num;
        if (!bayesian) {
            double p = avg * 0.5;
            scale = (0.0 < p && p < 1.0) ? 1.0 / sqrt(__dmul_rn(p, 1.0 - p)) : 0.0;
        } else {
            double p = (double)(s.sum + 1) / (double)(2 * s.num + 2);
            scale = 1.0 / sqrt(__dmul_rn(p, 1.0 - p));
        }
    }
    ab[l] = make_double2(avg, scale);
}

// CPCA_SNPCorr::SNP_PC_Corr (src/genPCA.cpp:820-847) from the three projections and the counts
__global__ void corr_finish_kernel(const double *__restrict__ xy, const double *__restrict__ x,
                                   const double *__restrict__ xx, const SnpStat *__restrict__ st, int64_t n_snp,
                                   int k, double *__restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_snp * k) return;
    const SnpStat s = st[idx / k];
    const double m = (double)s.num, Y = (double)s.sum, YY = (double)(2 * s.sum - s.n1);   // n1 + 4 n2
    double ans = __longlong_as_double(0x7ff8000000000000ll);
    if (s.num > 1) {
        const double X = x[idx], c1 = xx[idx] - X * X / m, c2 = YY - Y * Y / m, val = c1 * c2;
        if (val > 0) ans = (xy[idx] - X * Y / m) / sqrt(val);
    }
    out[idx] = ans;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static void need_ws(snprel_ctx *c, int k, const char *who) {
    if (c->n_samp <= 0) fail("%s: no genotype workspace (call snprel_geno_begin first)", who);
    if (k <= 0) fail("%s: the number of eigenvectors must be positive", who);
}

// dense operand [rows_pad][kp] on the device from a host matrix given as element(row, col) =
// src[row * rs + col * cs] * colscale[col]
static void upload_panel(snprel_ctx *c, DevBuf<double> &dev, const double *src, int64_t rows, int64_t rows_pad,
                         int k, int kp, int64_t rs, int64_t cs, const double *colscale) {
    if (cs == 1 && rs == kp && k == kp && !colscale) {   // already in the device layout: no host staging
        dev.alloc((size_t)rows_pad * kp);
        if (rows_pad > rows)
            CUDA_CHECK(cudaMemsetAsync(dev.p + (size_t)rows * kp, 0, (size_t)(rows_pad - rows) * kp * sizeof(double),
                                       c->stream));
        CUDA_CHECK(cudaMemcpyAsync(dev.p, src, (size_t)rows * kp * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        return;
    }
    std::vector<double> h((size_t)rows_pad * kp, 0.0);
    if (cs == 1) {   // source rows are contiguous (loadings [n_snp][k]): walk both sides linearly
        for (int64_t r = 0; r < rows; r++)
            for (int kk = 0; kk < k; kk++) {
                const double v = src[r * rs + kk];
                h[(size_t)r * kp + kk] = colscale ? v * colscale[kk] : v;
            }
    } else {
        for (int kk = 0; kk < k; kk++) {
            const double sc = colscale ? colscale[kk] : 1.0;
            for (int64_t r = 0; r < rows; r++) {
                const double v = src[r * rs + kk * cs];
                h[(size_t)r * kp + kk] = colscale ? v * sc : v;
            }
        }
    }
    dev.alloc(h.size());
    CUDA_CHECK(cudaMemcpyAsync(dev.p, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

static void run_snp_project(snprel_ctx *c, int mode, const double2 *ab, const double *Vt, int kp, int k,
                            double *out_dev, int64_t ld_out = 0) {
    if (c->n_snp == 0) return;
    dim3 grid((unsigned)((c->n_snp + PJ_ROWS - 1) / PJ_ROWS), (unsigned)(kp / PJ_COLS));
    CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
    snp_project_kernel<<<grid, PJ_ROWS, 0, c->stream>>>(c->geno2b.p, c->row_bytes, c->n_snp, c->n_samp_pad, mode, ab,
                                                        Vt, kp, k, out_dev, ld_out > 0 ? ld_out : (int64_t)k);
    KERNEL_CHECK(c);
    CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->hot_ms = ms;   // snprel_last_hot_kernel: the projection kernel of this call
    c->hot_launches = 1;
    c->hot_units = (double)c->n_samp * (double)c->n_snp * (double)k;
}

static void upload_ab(snprel_ctx *c, DevBuf<double2> &ab, const double *a, double amul, const double *b, double bconst) {
    std::vector<double2> h((size_t)std::max<int64_t>(c->n_snp, 1));
    for (int64_t l = 0; l < c->n_snp; l++) h[l] = make_double2(a[l] * amul, b ? b[l] : bconst);
    ab.alloc(h.size());
    CUDA_CHECK(cudaMemcpyAsync(ab.p, h.data(), h.size() * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

static void run_samp_project(snprel_ctx *c, const double2 *ab, const double *loadings /*host [M][k]*/, int k,
                             double *out /*host [k][n]*/) {
    const int64_t n = c->n_samp, npad = c->n_samp_pad, m = c->n_snp;
    const int kp = (int)round_up(k, PJ_COLS);
    geno_pad_tail(c);
    DevBuf<double> Lt, part, o;
    upload_panel(c, Lt, loadings, m, round_up(std::max<int64_t>(m, 1), PJ_STEP), k, kp, k, 1, nullptr);
    const int64_t row_tiles = npad / PJ_ROWS;
    int64_t splits = std::max<int64_t>(1, ((int64_t)c->num_sms * 4 + row_tiles * (kp / PJ_COLS) - 1) /
                                              (row_tiles * (kp / PJ_COLS)));
    splits = std::min<int64_t>(splits, std::max<int64_t>(1, (m + 1023) / 1024));
    const int64_t sps = round_up((std::max<int64_t>(m, 1) + splits - 1) / splits, PJ_STEP);
    splits = std::max<int64_t>(1, (m + sps - 1) / sps);
    part.alloc((size_t)splits * kp * npad);
    o.alloc((size_t)k * n);
    dim3 grid((unsigned)row_tiles, (unsigned)(kp / PJ_COLS), (unsigned)splits);
    CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
    samp_project_kernel<<<grid, PJ_ROWS, 0, c->stream>>>(c->geno2b.p, c->row_bytes, m, npad, ab, Lt.p, kp, sps, part.p);
    KERNEL_CHECK(c);
    CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
    dim3 rgrid((unsigned)((n + 255) / 256), (unsigned)k);
    samp_reduce_kernel<<<rgrid, 256, 0, c->stream>>>(part.p, (int)splits, kp, npad, k, n, o.p);
    KERNEL_CHECK(c);
    CUDA_CHECK(cudaMemcpyAsync(out, o.p, (size_t)k * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->hot_ms = ms;
    c->hot_launches = 1;
    c->hot_units = (double)n * (double)m * (double)k;
}

// gnrPCASNPLoading (src/genPCA.cpp:1489-1540)
void pca_snp_loading(snprel_ctx *c, int k, const double *eigval, const double *eigvect, double trace_xtx,
                     int bayesian, double *loading, double *avgfreq, double *scale) {
    need_ws(c, k, "snprel_pca_snp_loading");
    if (!eigval || !eigvect || !loading) fail("snprel_pca_snp_loading: NULL argument");
    ensure_stats(c);
    const int64_t n = c->n_samp, m = c->n_snp;
    const int kp = (int)round_up(k, PJ_COLS);
    // eigenvectors scaled by sqrt(Scale / eigenval_i), Scale = (n-1)/TraceXTX (:1500-1510)
    std::vector<double> cs((size_t)k);
    const double Scale = (double)(n - 1) / trace_xtx;
    for (int i = 0; i < k; i++) cs[i] = std::sqrt(Scale / eigval[i]);
    DevBuf<double> Vt, o;
    upload_panel(c, Vt, eigvect, n, c->n_samp_pad, k, kp, 1, n, cs.data());
    DevBuf<double2> ab;
    ab.alloc((size_t)std::max<int64_t>(m, 1));
    if (m > 0) {
        load_scale_kernel<<<(unsigned)((m + 255) / 256), 256, 0, c->stream>>>(c->stat.p, m, bayesian, ab.p);
        KERNEL_CHECK(c);
    }
    o.alloc((size_t)std::max<int64_t>(m, 1) * k);
    run_snp_project(c, 0, ab.p, Vt.p, kp, k, o.p);
    CUDA_CHECK(cudaMemcpyAsync(loading, o.p, (size_t)m * k * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    std::vector<double2> hab((size_t)m);
    CUDA_CHECK(cudaMemcpyAsync(hab.data(), ab.p, (size_t)m * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    for (int64_t l = 0; l < m; l++) {
        if (avgfreq) avgfreq[l] = hab[l].x;
        if (scale) scale[l] = hab[l].y;
    }
}

// gnrPCASampLoading (src/genPCA.cpp:1542-1563)
void pca_samp_loading(snprel_ctx *c, int k, const double *loadings, const double *avgfreq, const double *scale,
                      double *out) {
    need_ws(c, k, "snprel_pca_samp_loading");
    if (!loadings || !avgfreq || !scale || !out) fail("snprel_pca_samp_loading: NULL argument");
    DevBuf<double2> ab;
    upload_ab(c, ab, avgfreq, 1.0, scale, 0.0);
    run_samp_project(c, ab.p, loadings, k, out);
}

// gnrPCACorr (src/genPCA.cpp:1456-1485)
void pca_corr(snprel_ctx *c, int k, const double *eigvect, double *out) {
    need_ws(c, k, "snprel_pca_corr");
    if (!eigvect || !out) fail("snprel_pca_corr: NULL argument");
    ensure_stats(c);
    const int64_t n = c->n_samp, m = c->n_snp;
    const int kp = (int)round_up(k, PJ_COLS);
    if (m == 0) return;
    DevBuf<double> V1, V2, xy, x, xx, o;
    upload_panel(c, V1, eigvect, n, c->n_samp_pad, k, kp, 1, n, nullptr);
    std::vector<double> sq((size_t)n * k);
    for (size_t e = 0; e < sq.size(); e++) sq[e] = eigvect[e] * eigvect[e];
    upload_panel(c, V2, sq.data(), n, c->n_samp_pad, k, kp, 1, n, nullptr);
    DevBuf<double2> ab;   // (g - 0) * 1: the genotype itself
    std::vector<double> zero((size_t)m, 0.0);
    upload_ab(c, ab, zero.data(), 1.0, nullptr, 1.0);
    xy.alloc((size_t)m * k);
    x.alloc((size_t)m * k);
    xx.alloc((size_t)m * k);
    o.alloc((size_t)m * k);
    run_snp_project(c, 0, ab.p, V1.p, kp, k, xy.p);        // sum x v over valid
    run_snp_project(c, 1, nullptr, V1.p, kp, k, x.p);      // sum v over valid
    run_snp_project(c, 1, nullptr, V2.p, kp, k, xx.p);     // sum v^2 over valid
    const int64_t total = m * k;
    corr_finish_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(xy.p, x.p, xx.p, c->stat.p, m, k, o.p);
    KERNEL_CHECK(c);
    CUDA_CHECK(cudaMemcpyAsync(out, o.p, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

// 1 / sqrt(sum_l 4 p_l (1 - p_l)), summed in SNP order (src/genEIGMIX.cpp:508-513, 611-616)
static double eigmix_afreq_scale(const double *afreq, int64_t m) {
    double sum = 0;
    for (int64_t i = 0; i < m; i++) sum += 4 * afreq[i] * (1 - afreq[i]);
    return 1 / std::sqrt(sum);
}

// gnrEigMixSNPLoading (src/genEIGMIX.cpp:739-775)
void eigmix_snp_loading(snprel_ctx *c, int k, const double *eigval, const double *eigvect, const double *afreq,
                        double *loading) {
    need_ws(c, k, "snprel_eigmix_snp_loading");
    if (!eigval || !eigvect || !afreq || !loading) fail("snprel_eigmix_snp_loading: NULL argument");
    const int64_t n = c->n_samp, m = c->n_snp;
    const int kp = (int)round_up(k, PJ_COLS);
    std::vector<double> cs((size_t)k);
    for (int i = 0; i < k; i++) cs[i] = std::sqrt(1 / eigval[i]);   // :751-756
    DevBuf<double> Vt, o;
    upload_panel(c, Vt, eigvect, n, c->n_samp_pad, k, kp, 1, n, cs.data());
    DevBuf<double2> ab;
    upload_ab(c, ab, afreq, 2.0, nullptr, eigmix_afreq_scale(afreq, m));
    o.alloc((size_t)std::max<int64_t>(m, 1) * k);
    run_snp_project(c, 0, ab.p, Vt.p, kp, k, o.p);
    CUDA_CHECK(cudaMemcpyAsync(loading, o.p, (size_t)m * k * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

// gnrEigMixSampLoading (src/genEIGMIX.cpp:777-803)
void eigmix_samp_loading(snprel_ctx *c, int k, const double *loadings, const double *afreq, double *out) {
    need_ws(c, k, "snprel_eigmix_samp_loading");
    if (!loadings || !afreq || !out) fail("snprel_eigmix_samp_loading: NULL argument");
    DevBuf<double2> ab;
    upload_ab(c, ab, afreq, 2.0, nullptr, eigmix_afreq_scale(afreq, c->n_snp));
    run_samp_project(c, ab.p, loadings, k, out);
}

// ---------------------------------------------------------------------------
// Randomized PCA: CRandomPCA (src/genPCA.cpp:469-796), gnrPCA algorithm "randomized" (:1436-1442)
//   Y[l][g] = (g - avg_l) * s_l,  s_l = 1 / sqrt(2 p_l (1 - p_l)) (0 outside (0,1)), missing -> 0   (:497-517)
//   H_it = Y G_it  (snp_project),   G_{it+1} = Y^T H_it / nSNP  (samp_project),  it = 0 .. iter_num  (:710-742)
//   V^T  = orthonormal basis of the row space of MatH [hsize x nSNP]                                  (:752)
//   T    = V^T Y [hsize x nSamp] (samp_project with hsize columns), sigma / right vectors = svd(T)    (:763-779)
// The reference takes V^T from a LAPACK SVD of MatH; T's singular values and right singular vectors
// only depend on the row SPACE of MatH (any other orthonormal basis is Q V^T with Q orthogonal, and
// svd(Q T) has the same sigma and right vectors), so a Householder QR of MatH^T (cuSOLVER geqrf + orgqr)
// serves; the two tall products run on the float64 projection kernels above, straight from the
// resident 2-bit matrix.  Single-thread semantics of the reference (its multi-threaded G update
// re-adds the per-thread partials of earlier blocks, :727-735).
// ---------------------------------------------------------------------------
__global__ void rand_scale_kernel(const SnpStat *__restrict__ st, int64_t n_snp, double2 *__restrict__ ab,
                                  double *__restrict__ trace_part) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double tr = 0;
    if (l < n_snp) {
        const SnpStat s = st[l];
        double avg = 0, scale = 0;
        if (s.num > 0) {
            avg = (double)s.sum / (double)s.num;
            const double p = avg * 0.5;
            scale = (0.0 < p && p < 1.0) ? 1.0 / sqrt(__dmul_rn(__dmul_rn(2.0, p), 1.0 - p)) : 0.0;
        }
        ab[l] = make_double2(avg, scale);
        double tab[3];
        z_table(0, make_double2(avg, scale), tab);
        const int n2 = (s.sum - s.n1) / 2, n0 = s.num - s.n1 - n2;
        tr = (double)n0 * tab[0] * tab[0] + (double)s.n1 * tab[1] * tab[1] + (double)n2 * tab[2] * tab[2];
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = tr;
    __syncthreads();
    for (int o = blockDim.x / 2; o; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) trace_part[blockIdx.x] = sh[0];
}

// G[i][kk] = mul * sum_z part[z][kk][i] (z ascending), written in the dense-operand layout [npad][kp]
__global__ void samp_reduce_panel_kernel(const double *__restrict__ part, int splits, int64_t kp_total, int64_t npad,
                                         int k_out, int64_t n, double mul, double *__restrict__ G, int kp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int kk = blockIdx.y;
    if (i >= n || kk >= k_out) return;
    double s = 0;
    for (int z = 0; z < splits; z++) s += part[((int64_t)z * kp_total + kk) * npad + i];
    G[i * kp + kk] = s * mul;
}

// part / splits for a device-side sample projection of k columns
struct SampPlan { int kp; int64_t splits, sps; };
static SampPlan samp_plan(snprel_ctx *c, int k) {
    const int64_t npad = c->n_samp_pad, m = c->n_snp;
    SampPlan p;
    p.kp = (int)round_up(k, PJ_COLS);
    const int64_t row_tiles = npad / PJ_ROWS, col_tiles = p.kp / PJ_COLS;
    int64_t splits = std::max<int64_t>(1, ((int64_t)c->num_sms * 4 + row_tiles * col_tiles - 1) / (row_tiles * col_tiles));
    splits = std::min<int64_t>(splits, std::max<int64_t>(1, (m + 1023) / 1024));
    p.sps = round_up((std::max<int64_t>(m, 1) + splits - 1) / splits, PJ_STEP);
    p.splits = std::max<int64_t>(1, (m + p.sps - 1) / p.sps);
    return p;
}

void pca_randomized(snprel_ctx *c, const double *aux_mat, int aux_dim, int iter_num, double *sigma, double *vt,
                    double *trace_xtx2) {
    if (c->n_samp <= 0 || c->n_snp <= 0) fail("snprel_pca_randomized: no genotype workspace");
    if (!aux_mat || aux_dim <= 0) fail("snprel_pca_randomized: 'aux.mat' / 'aux.dim' are required");
    if (iter_num < 0) fail("snprel_pca_randomized: invalid 'iter.num'");
    ensure_stats(c);
    geno_pad_tail(c);
    const int64_t n = c->n_samp, npad = c->n_samp_pad, m = c->n_snp;
    const int64_t hsize = (int64_t)aux_dim * (iter_num + 1);
    if (hsize > m) fail("snprel_pca_randomized: aux.dim * (iter.num + 1) = %lld exceeds the number of SNPs (%lld)",
                        (long long)hsize, (long long)m);
    const int kp = (int)round_up(aux_dim, PJ_COLS);
    const int64_t ld = round_up(hsize, PJ_COLS) + PJ_COLS;     // row pitch of MatH: a 32-column panel read never leaves the row
    const int64_t m32 = round_up(m, PJ_STEP);

    // Y lookup tables and TraceXTX
    DevBuf<double2> ab;
    DevBuf<double> tpart;
    const unsigned tb = (unsigned)((m + 255) / 256);
    ab.alloc((size_t)m);
    tpart.alloc(tb);
    rand_scale_kernel<<<tb, 256, 0, c->stream>>>(c->stat.p, m, ab.p, tpart.p);
    KERNEL_CHECK(c);
    {
        std::vector<double> h(tb);
        CUDA_CHECK(cudaMemcpyAsync(h.data(), tpart.p, tb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        double tr = 0;
        for (unsigned b = 0; b < tb; b++) tr += h[b];
        if (trace_xtx2) *trace_xtx2 = 2.0 * tr;     // :792
    }

    // power iteration
    DevBuf<double> G, H, part;
    upload_panel(c, G, aux_mat, n, npad, aux_dim, kp, 1, n, nullptr);      // G_0[i][j] = aux.mat[j * nSamp + i]
    H.alloc((size_t)m32 * ld);
    H.zero(c->stream);
    const SampPlan sp = samp_plan(c, aux_dim);
    part.alloc((size_t)sp.splits * sp.kp * npad);
    for (int it = 0; it <= iter_num; it++) {
        run_snp_project(c, 0, ab.p, G.p, kp, aux_dim, H.p + (int64_t)it * aux_dim, ld);
        if (it == iter_num) break;
        dim3 grid((unsigned)(npad / PJ_ROWS), (unsigned)(sp.kp / PJ_COLS), (unsigned)sp.splits);
        samp_project_kernel<<<grid, PJ_ROWS, 0, c->stream>>>(c->geno2b.p, c->row_bytes, m, npad, ab.p,
                                                             H.p + (int64_t)it * aux_dim, ld, sp.sps, part.p);
        KERNEL_CHECK(c);
        dim3 rgrid((unsigned)((n + 255) / 256), (unsigned)aux_dim);
        samp_reduce_panel_kernel<<<rgrid, 256, 0, c->stream>>>(part.p, (int)sp.splits, sp.kp, npad, aux_dim, n,
                                                              1.0 / (double)m, G.p, kp);
        KERNEL_CHECK(c);
    }
    part.release();

    // orthonormal basis of the row space of MatH: Householder QR of MatH^T [nSNP x hsize]
    {
        DevBuf<double> Q;
        Q.alloc((size_t)m * hsize);
        la_transpose(c, m, hsize, H.p, ld, Q.p, m);            // H is column-major [ld x nSNP]
        la_orthonormalise(c, Q.p, m, (int)hsize);
        la_transpose(c, hsize, m, Q.p, m, H.p, ld);            // rows 0 .. hsize-1 of H <- Q^T
    }

    // T = Q^T Y  [hsize x nSamp]
    const SampPlan st = samp_plan(c, (int)hsize);
    part.alloc((size_t)st.splits * st.kp * npad);
    DevBuf<double> T;
    T.alloc((size_t)hsize * n);
    {
        dim3 grid((unsigned)(npad / PJ_ROWS), (unsigned)(st.kp / PJ_COLS), (unsigned)st.splits);
        CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
        samp_project_kernel<<<grid, PJ_ROWS, 0, c->stream>>>(c->geno2b.p, c->row_bytes, m, npad, ab.p, H.p, ld, st.sps, part.p);
        KERNEL_CHECK(c);
        CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
        dim3 rgrid((unsigned)((n + 255) / 256), (unsigned)hsize);
        samp_reduce_kernel<<<rgrid, 256, 0, c->stream>>>(part.p, (int)st.splits, st.kp, npad, (int)hsize, n, T.p);
        KERNEL_CHECK(c);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        c->hot_ms = ms;
        c->hot_launches = 1;
        c->hot_units = (double)n * (double)m * (double)hsize;
    }
    part.release();
    H.release();

    // sigma and the right singular vectors of T.  T.p holds T row-major [hsize][n] = T^T column-major [n x hsize].
    const int64_t r = std::min<int64_t>(hsize, n);
    std::vector<double> hs((size_t)r);
    if (sigma) std::fill(sigma, sigma + n, 0.0);           // vector<double> sigma(nSamp), :777
    if (vt) std::fill(vt, vt + hsize * n, 0.0);
    if (n >= hsize) {
        DevBuf<double> U;
        U.alloc((size_t)n * hsize);
        la_svd_tall(c, T.p, n, (int)hsize, hs.data(), U.p, nullptr);         // T^T = U S W^T: the columns of U
        if (vt) CUDA_CHECK(cudaMemcpyAsync(vt, U.p, (size_t)n * hsize * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    } else {
        DevBuf<double> Tc, VT;
        Tc.alloc((size_t)hsize * n);
        VT.alloc((size_t)n * n);
        la_transpose(c, hsize, n, T.p, n, Tc.p, hsize);                      // T column-major [hsize x n]
        la_svd_tall(c, Tc.p, hsize, (int)n, hs.data(), nullptr, VT.p);       // T = A S V^T: the rows of V^T
        std::vector<double> hv((size_t)n * n);
        CUDA_CHECK(cudaMemcpyAsync(hv.data(), VT.p, hv.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (vt)
            for (int64_t h = 0; h < n; h++)
                for (int64_t i = 0; i < n; i++) vt[h * n + i] = hv[(size_t)(h + i * n)];
    }
    if (sigma) std::copy(hs.begin(), hs.end(), sigma);
}

}  // namespace snprel
