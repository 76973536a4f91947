// Covariance-type estimators on top of the tcgen05 table Gram (gram_tc.cu):
//   Eigenstrat / PCA   CExactPCA::Run + gnrPCA       src/genPCA.cpp:395-464,1355-1452
//   GCTA, Corr         CGCTA_AlgArith::Run + gnrGRM  src/genPCA.cpp:1148-1237,1614-1717
//   EIGMIX             CEigMix_AlgArith::Run         src/genEIGMIX.cpp:60-156,645-735
//   KING-homo sums     CKINGHomo                     src/genKING.cpp:69-266,493-570
//
// Algebra.  With g the genotype code, v = valid and m = missing indicators, mu_l the mean
// genotype of SNP l over non-missing samples and w_l the per-SNP weight (1/(p(1-p)) for
// Eigenstrat/GCTA, 1 for EIGMIX), the reference accumulates
//     C_ij = sum_l w_l z_il z_jl,     z_il = (g_il - mu_l) v_il.
// Per SNP pick integers s_l >= 1, t_l >= 0 with the INTEGER column table
//     B_l[g] = s_l g - t_l  (|B_l[g]| <= 127, B_l[missing] = 0),      delta_l = mu_l - t_l / s_l,
// so that z_jl = B_l[g_jl] / s_l - delta_l v_jl exactly.  With the real row tables
//     U_l[g] = w_l (g - mu_l) (0 for missing),   T_l = U_l / s_l,   R_l = delta_l U_l
// this gives
//     C_ij = sum_l T_l[g_il] B_l[g_jl]  -  sum_l R_l[g_il]  +  sum_l R_l[g_il] m_jl.
// The first and third sums are table Grams (real row table against a small-integer column
// channel: B_l, resp. the missing indicator); the middle one is a per-sample vector.  s_l follows
// the magnitude of U_l (the row tables of all SNPs get the same dynamic range) and t_l / s_l is
// the best rational approximation of mu_l in a window of s, which makes R two orders of
// magnitude smaller than the plain choice s = 1, t = 0 (R = mu U) and saves one digit pass.
// Row tables are quantised to fixed point (2^-frac_bits) and split into balanced base-256
// digits, one int8 tensor-core pass per digit; the quantisation error of an entry is bounded by
//     2^-(f+1) max_j sum_l |B_l[g_jl]|  +  2^-(fw+1) max_j #missing_j
// (both maxima are MEASURED by sample_stats_kernel) and everything after quantisation is exact
// integer arithmetic.  Denominators (GCTA: #polymorphic SNPs with i or j missing; EIGMIX: sum
// 4p(1-p) over SNPs with i or j missing) use
//     D_ij = r_i + r_j - sum_l d_l m_il m_jl,   r_i = sum_l d_l m_il,
// i.e. one more table Gram (channel m on both sides) and a per-sample vector.
#include <cmath>

#include "common.cuh"

namespace snprel {

constexpr uint32_t TABB_M = 0x01000000u;   // m channel: code -> {0,0,0,1}
constexpr int MAX_DIGITS = 8;

enum { VEC_W = 0, VEC_D = 1, VEC_HET = 2, VEC_D2 = 3, VEC_WLO = 4, NVEC = 5 };
constexpr int LIMB = 20;           // the per-sample vector is accumulated in 20-bit limbs of int32
// VEC_W is in units of 2^-(frac_bits_v - 20), VEC_WLO in 2^-frac_bits_v
// scalars[]: 0 = sum of d_l (EIGMIX SumDenominator / KING-homo sum p(1-p)), 1 = sum d2_l
// iscalars[]: 0 = nLocus (GCTA), 2 = low limb and 3 = upper limbs of sum_l qb_l (constant of the vector)

constexpr int AUX_MAX = 84;        // largest entry of the auxiliary byte tables (diagonal bound, squares): THREE rows add up in a byte lane
constexpr int SQ_UNIT = 193;       // unit of the per-sample sum of squared column-table entries: ceil(127^2 / 193) = 84

struct SnpTables {
    long long qU[4];   // fixed-point T = U / s (row table of the main passes)
    long long qW[4];   // fixed-point R = delta U (row table of the missing-data passes)
    long long qa, qb;  // the per-sample vector in linear form: R[g] = a g - b, fixed point 2^-frac_bits_v
    double diag;       // sum over samples of U[g] (g - mu): this SNP's contribution to trace(C)
    long long qD;      // fixed-point d (GCTA: 0/1 unscaled; EIGMIX / KING-homo: 2^frac_bits scaled)
    long long qD2;     // KING-homo: (p(1-p))^2
    double d, d2;      // float64 d, d2 (for the global scalars)
    double maxU, maxW; // max |T|, max |R|
    int maxB;          // max |B_l[g]|
};

constexpr uint64_t DITHER_SEED = 0xD17E5ull | 1ull;
// uniform [0, 1) draw of table entry (global SNP index l, genotype g): randomised rounding (snprel_set_rounding)
__device__ __forceinline__ double dither_u01(uint64_t seed, long long l, int g) {
    uint64_t x = seed ^ ((uint64_t)l * 4u + (uint64_t)g) * 0xD1342543DE82EF95ull;
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

// mean, weight and denominators of one SNP.  est: SNPREL_GRM_EIGENSTRAT / GCTA / EIGMIX or
// SNPREL_EST_KING_HOMO
struct SnpCoef {
    double mu, w, d, d2;
    long long qD;   // GCTA: polymorphic flag
};
__device__ __forceinline__ SnpCoef snp_coef(const SnpStat st, int est, int bayesian) {
    SnpCoef k;
    k.mu = st.num > 0 ? (double)st.sum / (double)st.num : 0.0;   // DivideGeno, src/genPCA.cpp:98-142
    k.w = 0;
    k.d = 0;
    k.d2 = 0;
    k.qD = 0;
    if (est == SNPREL_GRM_EIGMIX) {
        k.w = 1.0;
        double af = 0.5 * k.mu;
        k.d = 4 * af * (1 - af);                  // src/genEIGMIX.cpp:116-119
    } else if (est == SNPREL_EST_KING_HOMO) {
        double p = st.num > 0 ? 0.5 * (double)st.sum / (double)st.num : 0.0;   // src/genKING.cpp:239-241
        k.d = p * (1 - p);
        k.d2 = k.d * k.d;
    } else {
        if (bayesian) {                           // src/genPCA.cpp:445-452
            double s = ((double)st.sum + 1.0) / (double)(2 * st.num + 2);
            double r = 1.0 / sqrt(s * (1 - s));
            k.w = r * r;
        } else {                                  // rsqrt_prod, src/genPCA.cpp:145-181
            double s = k.mu * 0.5;
            if (0 < s && s < 1) {
                double r = 1.0 / sqrt(s * (1 - s));
                k.w = r * r;
            }
        }
        bool poly = (0 < st.sum) && (st.sum < 2 * st.num);   // src/genPCA.cpp:1206
        k.d = poly ? 1.0 : 0.0;
        k.qD = poly ? 1 : 0;
    }
    return k;
}

// ---- per-SNP integer column tables ---------------------------------------------
// coltab[l] = (s_l, t_l); tabB[l] = bytes (B[0], B[1], B[2], 0); tabBabs[l] = their magnitudes;
// tabBsq[l] = bytes ceil(B[g]^2 / SQ_UNIT) (<= AUX_MAX): summed per sample they bound sum_l B_l[g]^2, the
// variance proxy of the randomised-rounding error bound (u_table_error).
// A pure function of the SNP's own counts, so every rank of a sharded run picks the same table
// for the same SNP without communication.
__global__ void coltab_kernel(const SnpStat *__restrict__ st, int64_t n_snp, int est, int bayesian,
                              int2 *__restrict__ coltab, uint32_t *__restrict__ tabB, uint32_t *__restrict__ tabBabs,
                              uint32_t *__restrict__ tabBsq) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_snp) return;
    const SnpCoef k = snp_coef(st[l], est, bayesian);
    int s = 1, t = 0;
    if (est != SNPREL_EST_KING_HOMO && k.w > 0 && st[l].num > 0) {
        const double mu = k.mu;
        const double umax = k.w * fmax(mu, 2.0 - mu);
        const double uref = est == SNPREL_GRM_EIGMIX ? 2.0 : 40.0;   // |U| at MAF 0.05: s saturates there
        const double s_hi = fmin(floor(127.0 / fmax(2.0 - mu, 1e-9)), 127.0);   // 2 s - t <= 127 with t ~ s mu
        // (s >= 24 keeps the residual R = delta U of the common SNPs, whose |U| is small, below ~0.1)
        const double s_tgt = fmin(s_hi, fmax(24.0, 127.0 * umax / uref));
        double best_e = 1e300;
        t = (int)rint(mu);
        for (int q = 0; q <= 30; q++) {      // the s in [0.7, 1] s_tgt whose t / s approximates mu best
            const double sc = fmax(1.0, floor(s_tgt * (0.7 + 0.01 * q)));
            const double tc = rint(sc * mu);
            if (2 * sc - tc > 127.0 || tc > 127.0) continue;
            const double e = fabs(mu - tc / sc);
            if (e < best_e) {
                best_e = e;
                s = (int)sc;
                t = (int)tc;
            }
        }
    }
    coltab[l] = make_int2(s, t);
    const int b0 = -t, b1 = s - t, b2 = 2 * s - t;
    tabB[l] = (uint32_t)(b0 & 255) | ((uint32_t)(b1 & 255) << 8) | ((uint32_t)(b2 & 255) << 16);
    tabBabs[l] = (uint32_t)abs(b0) | ((uint32_t)abs(b1) << 8) | ((uint32_t)abs(b2) << 16);
    tabBsq[l] = (uint32_t)((b0 * b0 + SQ_UNIT - 1) / SQ_UNIT) | ((uint32_t)((b1 * b1 + SQ_UNIT - 1) / SQ_UNIT) << 8) |
                ((uint32_t)((b2 * b2 + SQ_UNIT - 1) / SQ_UNIT) << 16);
}

// dither != 0: the T table is rounded at random, floor(v 2^f + u) with u ~ U[0, 1) drawn per
// (global SNP index snp_index, genotype) -- unbiased and independent across SNPs, which is what the
// Hoeffding bound of choose_format relies on; 0: round to nearest (the default, worst-case bound)
__device__ __forceinline__ void snp_tables(const SnpStat st, int est, int bayesian, int2 ct, int frac_bits,
                                           int frac_bits_w, int frac_bits_d, int frac_bits_v, SnpTables &t,
                                           long long snp_index = 0, uint64_t dither = 0) {
    const double sc = exp2((double)frac_bits);
    const double scw = exp2((double)frac_bits_w);
    const double scd = exp2((double)frac_bits_d);
    const double scv = exp2((double)frac_bits_v);
    const SnpCoef k = snp_coef(st, est, bayesian);
    const double mu = k.mu, w = k.w;
    const double inv_s = 1.0 / (double)ct.x;
    const double delta = mu - (double)ct.y * inv_s;
    t.maxU = 0;
    t.maxW = 0;
    const int n2 = (st.sum - st.n1) / 2, n0 = st.num - st.n1 - n2;
    const double cnt[3] = {(double)n0, (double)st.n1, (double)n2};
    t.diag = 0;
#pragma unroll
    for (int g = 0; g < 3; g++) {
        const double u = (est == SNPREL_EST_KING_HOMO) ? 0.0 : w * ((double)g - mu);
        const double tt = u * inv_s, rr = delta * u;
        t.qU[g] = dither ? (long long)floor(tt * sc + dither_u01(dither, snp_index, g)) : llrint(tt * sc);
        t.qW[g] = llrint(rr * scw);
        t.maxU = fmax(t.maxU, fabs(tt));
        t.maxW = fmax(t.maxW, fabs(rr));
        t.diag += cnt[g] * u * ((double)g - mu);
    }
    t.qU[3] = 0;
    t.qW[3] = 0;
    const double a = (est == SNPREL_EST_KING_HOMO) ? 0.0 : delta * w;
    t.qa = llrint(a * scv);
    t.qb = llrint(a * mu * scv);
    t.maxB = max(ct.y, max(abs(ct.x - ct.y), 2 * ct.x - ct.y));
    const bool realD = est == SNPREL_GRM_EIGMIX || est == SNPREL_EST_KING_HOMO;
    t.qD = realD ? llrint(k.d * scd) : k.qD;
    t.qD2 = est == SNPREL_EST_KING_HOMO ? llrint(k.d2 * scd) : 0;
    t.d = k.d;
    t.d2 = k.d2;
}

__device__ __forceinline__ uint32_t digit_of(long long &q) {
    long long dgt = ((q + 128) & 255) - 128;   // balanced digit in [-128, 127]
    q = (q - dgt) >> 8;
    return (uint32_t)(dgt & 255);
}

// ---- plan statistics ---------------------------------------------------------
// out[0] max |T|, out[1] int64-range bound, out[2] total missing,
// out[3] local share of the normaliser (trace(C) / nLocus / sum d / sum d2), out[4] max |R|,
// out[5] max over SNPs and genotypes of w (g - mu)^2 (scale of the measured diagonal bound, diagtab_kernel),
// out[6] the part of out[1] that is not the main T x B product (2 max|R| + 2 dmax + w delta^2)
__global__ void plan_kernel(const SnpStat *__restrict__ st, const int2 *__restrict__ coltab, int64_t n_snp,
                            int64_t n_samp, int est, int bayesian, double *__restrict__ out) {
    double mx = 0, mxw = 0, sb = 0, tm = 0, sc = 0, mv = 0, sr = 0;
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < n_snp;
         l += (int64_t)gridDim.x * blockDim.x) {
        SnpTables t;
        snp_tables(st[l], est, bayesian, coltab[l], 0, 0, 0, 0, t);
        const bool realD = (est == SNPREL_GRM_EIGMIX || est == SNPREL_EST_KING_HOMO);
        double dmax = realD ? fmax(t.d, t.d2) : 0.0;
        mx = fmax(mx, t.maxU);
        mxw = fmax(mxw, t.maxW);
        const SnpCoef k = snp_coef(st[l], est, bayesian);
        const double delta = k.mu - (double)coltab[l].y / (double)coltab[l].x;
        const double rest = 2 * t.maxW + 2 * dmax + k.w * delta * delta;
        sb += (double)t.maxB * t.maxU + rest;
        sr += rest;
        if (est != SNPREL_EST_KING_HOMO) mv = fmax(mv, k.w * fmax(k.mu, 2.0 - k.mu) * fmax(k.mu, 2.0 - k.mu));
        tm += (double)(n_samp - st[l].num);
        if (est == SNPREL_GRM_EIGENSTRAT) sc += t.diag;
        else if (est == SNPREL_GRM_EIGMIX) sc += t.d;
        else if (est == SNPREL_EST_KING_HOMO) sc += t.d2;
        else sc += t.d;   // GCTA: number of polymorphic SNPs
    }
    __shared__ double s0[256], s1[256], s2[256], s3[256], s4[256], s5[256], s6[256];
    s5[threadIdx.x] = mv;
    s6[threadIdx.x] = sr;
    s4[threadIdx.x] = mxw;
    s0[threadIdx.x] = mx;
    s1[threadIdx.x] = sb;
    s2[threadIdx.x] = tm;
    s3[threadIdx.x] = sc;
    __syncthreads();
    for (int o = blockDim.x / 2; o; o >>= 1) {
        if (threadIdx.x < o) {
            s0[threadIdx.x] = fmax(s0[threadIdx.x], s0[threadIdx.x + o]);
            s4[threadIdx.x] = fmax(s4[threadIdx.x], s4[threadIdx.x + o]);
            s5[threadIdx.x] = fmax(s5[threadIdx.x], s5[threadIdx.x + o]);
            s6[threadIdx.x] += s6[threadIdx.x + o];
            s1[threadIdx.x] += s1[threadIdx.x + o];
            s2[threadIdx.x] += s2[threadIdx.x + o];
            s3[threadIdx.x] += s3[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicMax(reinterpret_cast<unsigned long long *>(out), (unsigned long long)__double_as_longlong(s0[0]));
        atomicAdd(out + 1, s1[0]);
        atomicAdd(out + 2, s2[0]);
        atomicAdd(out + 3, s3[0]);
        atomicMax(reinterpret_cast<unsigned long long *>(out + 4), (unsigned long long)__double_as_longlong(s4[0]));
        atomicMax(reinterpret_cast<unsigned long long *>(out + 5), (unsigned long long)__double_as_longlong(s5[0]));
        atomicAdd(out + 6, s6[0]);
    }
}

// ---- measured bound of the diagonal --------------------------------------------------------
// tabF[l] = bytes ceil(w_l (g - mu_l)^2 / c) for g = 0, 1, 2 (0 for missing), c = out[5] / AUX_MAX: summed per
// sample by sample_stats_kernel it gives X_i >= C_ii, and by Cauchy-Schwarz every entry of the main
// plane obeys |sum_l T_l[g_il] B_l[g_jl]| <= sqrt(X_i (2 X_j + 2 sum_l w_l delta_l^2)) -- a bound that follows
// the data (about 2 per SNP) where the worst case sum_l max|T_l| max|B_l| assumes a sample that is
// homozygous for the rare allele everywhere (about 16 per SNP at the bench's MAF range).  It only decides how
// many fractional bits the int64 planes can carry.
__global__ void diagtab_kernel(const SnpStat *__restrict__ st, int64_t n_snp, int est, int bayesian,
                               const double *__restrict__ plan_out, uint32_t *__restrict__ tabF) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_snp) return;
    const SnpCoef k = snp_coef(st[l], est, bayesian);
    const double c = plan_out[5] / (double)AUX_MAX;
    uint32_t word = 0;
    if (c > 0 && k.w > 0) {
        for (int g = 0; g < 3; g++) {
            const double v = k.w * ((double)g - k.mu) * ((double)g - k.mu);
            const double q = fmin((double)AUX_MAX, ceil(v / c * (1.0 + 1e-12)));
            word |= (uint32_t)q << (8 * g);
        }
    }
    tabF[l] = word;
}

// ---- per-sample statistics in one pass over the 2-bit matrix -----------------
//   ew[i]      = sum_l |B_l[g_il]|     (error weight of the main passes; int64)
//   sq[i]      = sum_l ceil(B_l[g_il]^2 / SQ_UNIT)   (optional, tabS != nullptr)
//   cnt[0][i]  = #heterozygous, cnt[1][i] = #missing
//   chunk_ew[y] = max over samples of the sum over SNP chunk y (int32 accumulator headroom of K1)
// One thread per 32-bit word column (16 samples), one block row per GRAM_CHUNK SNPs.  |B| comes from a
// byte-permute of the SNP's magnitude table (the selector trick of K1's producers) and is summed
// in 16-bit lanes (512 x 127 < 2^16); the two indicator counts use bit-sliced vertical counters
// (3 rows in 2-bit fields -> 15 in 4-bit fields -> 255 in bytes -> 16-bit lanes).
constexpr int SS_THREADS = 128;
template <bool with_f, bool with_s>
__global__ void __launch_bounds__(SS_THREADS)
sample_stats_kernel(const uint32_t *__restrict__ geno, const uint32_t *__restrict__ tabBabs, int64_t n_snp,
                    int64_t row_words, int64_t npad, long long *__restrict__ ew, int *__restrict__ cnt,
                    int *__restrict__ chunk_ew, const uint32_t *__restrict__ tabF, long long *__restrict__ dg,
                    const uint32_t *__restrict__ tabS, long long *__restrict__ sq) {
    __shared__ uint32_t tab[GRAM_CHUNK + 4], tabf[GRAM_CHUNK + 4], tabs[GRAM_CHUNK + 4];
    __shared__ int smax[SS_THREADS / 32];
    const int64_t l0 = (int64_t)blockIdx.y * GRAM_CHUNK;
    const int nl = (int)min((int64_t)GRAM_CHUNK, n_snp - l0);
    for (int s = threadIdx.x; s < GRAM_CHUNK + 4; s += blockDim.x) {
        tab[s] = s < nl ? tabBabs[l0 + s] : 0u;
        tabf[s] = (with_f && s < nl) ? tabF[l0 + s] : 0u;
        tabs[s] = (with_s && s < nl) ? tabS[l0 + s] : 0u;
    }
    __syncthreads();
    const int64_t wc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = wc < row_words;
    const uint32_t *p = geno + l0 * row_words + (live ? wc : 0);
    uint32_t e16[8] = {0, 0, 0, 0, 0, 0, 0, 0};        // |B| sums: selector q -> e16[2q] (bytes 0,2), e16[2q+1] (bytes 1,3)
    uint32_t f16[8] = {0, 0, 0, 0, 0, 0, 0, 0};        // diagonal-bound sums, same lanes (entries <= AUX_MAX: 512 rows fit 16 bits)
    uint32_t s16[8] = {0, 0, 0, 0, 0, 0, 0, 0};        // squared column-table sums in SQ_UNIT, same lanes
    uint32_t m16[8] = {0, 0, 0, 0, 0, 0, 0, 0}, h16[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t m8[4] = {0, 0, 0, 0}, h8[4] = {0, 0, 0, 0};
    int in8 = 0;
    for (int r0 = 0; r0 < nl; r0 += 15) {
        uint32_t m4a = 0, m4b = 0, h4a = 0, h4b = 0;
#pragma unroll
        for (int g5 = 0; g5 < 5; g5++) {
            uint32_t m2 = 0, h2 = 0;
            uint32_t fb[4] = {0, 0, 0, 0}, sb[4] = {0, 0, 0, 0};   // byte-lane sums of the group's three rows (3 x AUX_MAX <= 255)
            uint32_t w[3];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const int row = r0 + 3 * g5 + r;
                w[r] = (live && row < nl) ? __ldg(p + (int64_t)row * row_words) : 0u;   // beyond the chunk: code 0 against a zero table
            }
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const int row = min(r0 + 3 * g5 + r, GRAM_CHUNK + 3);
                const uint32_t x = w[r], t = tab[row], tf = tabf[row], ts = tabs[row];
                const uint32_t hi = x >> 1;
                m2 += x & hi & 0x55555555u;
                h2 += x & ~hi & 0x55555555u;
                const uint32_t e = x & 0x33333333u, o = hi >> 1 & 0x33333333u;
                const uint32_t sel[4] = {e, o, e >> 16, o >> 16};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t v = __byte_perm(t, 0, sel[q]);
                    e16[2 * q] += __byte_perm(v, 0, 0x4240);
                    e16[2 * q + 1] += __byte_perm(v, 0, 0x4341);
                    if (with_f) fb[q] += __byte_perm(tf, 0, sel[q]);
                    if (with_s) sb[q] += __byte_perm(ts, 0, sel[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {        // spread the three-row byte sums into the 16-bit lanes
                if (with_f) {
                    f16[2 * q] += __byte_perm(fb[q], 0, 0x4240);
                    f16[2 * q + 1] += __byte_perm(fb[q], 0, 0x4341);
                }
                if (with_s) {
                    s16[2 * q] += __byte_perm(sb[q], 0, 0x4240);
                    s16[2 * q + 1] += __byte_perm(sb[q], 0, 0x4341);
                }
            }
            m4a += m2 & 0x33333333u;
            m4b += (m2 >> 2) & 0x33333333u;
            h4a += h2 & 0x33333333u;
            h4b += (h2 >> 2) & 0x33333333u;
        }
        m8[0] += m4a & 0x0F0F0F0Fu;
        m8[1] += (m4a >> 4) & 0x0F0F0F0Fu;
        m8[2] += m4b & 0x0F0F0F0Fu;
        m8[3] += (m4b >> 4) & 0x0F0F0F0Fu;
        h8[0] += h4a & 0x0F0F0F0Fu;
        h8[1] += (h4a >> 4) & 0x0F0F0F0Fu;
        h8[2] += h4b & 0x0F0F0F0Fu;
        h8[3] += (h4b >> 4) & 0x0F0F0F0Fu;
        if (++in8 == 17 || r0 + 15 >= nl) {      // 17 x 15 = 255 rows: bytes are about to overflow
#pragma unroll
            for (int q = 0; q < 4; q++) {
                m16[2 * q] += m8[q] & 0x00FF00FFu;
                m16[2 * q + 1] += (m8[q] >> 8) & 0x00FF00FFu;
                h16[2 * q] += h8[q] & 0x00FF00FFu;
                h16[2 * q + 1] += (h8[q] >> 8) & 0x00FF00FFu;
                m8[q] = 0;
                h8[q] = 0;
            }
            in8 = 0;
        }
    }
    // sample positions.  |B| lanes: selector q = (e, o, e>>16, o>>16) covers samples 2j + (q & 1) + 8 (q >> 1),
    // j = byte index 0..3; e16[2q] holds bytes 0 and 2, e16[2q+1] bytes 1 and 3.
    // indicator lanes: m8[0] samples 0,4,8,12; m8[1] 2,6,10,14; m8[2] 1,5,9,13; m8[3] 3,7,11,15 (byte index b);
    // m16[2q] holds bytes 0 and 2 of m8[q], m16[2q+1] bytes 1 and 3.
    int best = 0;
    if (live) {
        const int64_t s0 = wc * 16;
#pragma unroll
        for (int q = 0; q < 4; q++) {
#pragma unroll
            for (int hb = 0; hb < 2; hb++) {
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    const int byte = hb + 2 * half;                     // byte index inside the 32-bit word
                    const int ev = (int)((e16[2 * q + hb] >> (16 * half)) & 0xFFFFu);
                    const int es = 2 * byte + (q & 1) + 8 * (q >> 1);
                    best = max(best, ev);
                    if (ev) atomicAdd(reinterpret_cast<unsigned long long *>(ew) + s0 + es, (unsigned long long)ev);
                    if (with_f) {
                        const int fv = (int)((f16[2 * q + hb] >> (16 * half)) & 0xFFFFu);
                        if (fv) atomicAdd(reinterpret_cast<unsigned long long *>(dg) + s0 + es, (unsigned long long)fv);
                    }
                    if (with_s) {
                        const int sv = (int)((s16[2 * q + hb] >> (16 * half)) & 0xFFFFu);
                        if (sv) atomicAdd(reinterpret_cast<unsigned long long *>(sq) + s0 + es, (unsigned long long)sv);
                    }
                    const int ms = (q == 0 ? 0 : q == 1 ? 2 : q == 2 ? 1 : 3) + 4 * byte;
                    const int mv = (int)((m16[2 * q + hb] >> (16 * half)) & 0xFFFFu);
                    const int hv = (int)((h16[2 * q + hb] >> (16 * half)) & 0xFFFFu);
                    if (hv) atomicAdd(cnt + s0 + ms, hv);
                    if (mv) atomicAdd(cnt + npad + s0 + ms, mv);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < SS_THREADS / 32; k++) best = max(best, smax[k]);
        if (best) atomicMax(chunk_ew + blockIdx.y, best);
    }
}

// ---- digit tables: tab[pass][snp] ------------------------------------------
// pass order: T digits (nU), R digits (nW), D digits (nD), D2 digits (nD2)
__global__ void tables_kernel(const SnpStat *__restrict__ st, const int2 *__restrict__ coltab, int64_t n_snp, int64_t cap, int est,
                              int bayesian, int frac_bits, int frac_bits_w, int frac_bits_d, int frac_bits_v, int nU, int nW, int nD, int nD2,
                              uint32_t *__restrict__ tab, double *__restrict__ scalars /*[gridDim.x][2] partials*/,
                              long long *__restrict__ iscalars, int *__restrict__ overflow, uint64_t dither,
                              long long snp_base /* global index of st[0] */) {
    int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double d = 0, d2 = 0;
    long long poly = 0, qb = 0;
    if (l < n_snp) {
        SnpTables t;
        snp_tables(st[l], est, bayesian, coltab[l], frac_bits, frac_bits_w, frac_bits_d, frac_bits_v, t, snp_base + (long long)l, dither);
        d = t.d;
        d2 = t.d2;
        qb = t.qb;
        poly = (est == SNPREL_GRM_GCTA || est == SNPREL_GRM_CORR) ? t.qD : 0;
        int pass = 0;
        long long q[4];
        for (int g = 0; g < 4; g++) q[g] = t.qU[g];
        for (int k = 0; k < nU; k++, pass++) {
            uint32_t word = 0;
            for (int g = 0; g < 4; g++) word |= digit_of(q[g]) << (8 * g);
            tab[(int64_t)pass * cap + l] = word;
        }
        for (int g = 0; g < 4; g++)
            if (nU > 0 && q[g] != 0) atomicExch(overflow, 1);
        for (int g = 0; g < 4; g++) q[g] = t.qW[g];
        for (int k = 0; k < nW; k++, pass++) {
            uint32_t word = 0;
            for (int g = 0; g < 4; g++) word |= digit_of(q[g]) << (8 * g);
            tab[(int64_t)pass * cap + l] = word;
        }
        for (int g = 0; g < 4; g++)
            if (nW > 0 && q[g] != 0) atomicExch(overflow, 2);
        long long qd = t.qD;
        for (int k = 0; k < nD; k++, pass++) tab[(int64_t)pass * cap + l] = digit_of(qd) << 24;
        if (nD > 0 && qd != 0) atomicExch(overflow, 3);
        long long qd2 = t.qD2;
        for (int k = 0; k < nD2; k++, pass++) tab[(int64_t)pass * cap + l] = digit_of(qd2) << 24;
        if (nD2 > 0 && qd2 != 0) atomicExch(overflow, 4);
    }
    // global scalars: deterministic per-block tree, then one atomic per block
    __shared__ double s0[256], s1[256];
    __shared__ long long s2[256], s3[256], s4[256];
    s0[threadIdx.x] = d;
    s1[threadIdx.x] = d2;
    s2[threadIdx.x] = poly;
    s3[threadIdx.x] = qb & ((1ll << LIMB) - 1);     // the constant of the per-sample vector, in two limbs
    s4[threadIdx.x] = qb >> LIMB;
    __syncthreads();
    for (int o = blockDim.x / 2; o; o >>= 1) {
        if (threadIdx.x < o) {
            s0[threadIdx.x] += s0[threadIdx.x + o];
            s1[threadIdx.x] += s1[threadIdx.x + o];
            s2[threadIdx.x] += s2[threadIdx.x + o];
            s3[threadIdx.x] += s3[threadIdx.x + o];
            s4[threadIdx.x] += s4[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // float64 block partials are summed in block order by the host (run-to-run deterministic)
        scalars[2 * blockIdx.x + 0] = s0[0];
        scalars[2 * blockIdx.x + 1] = s1[0];
        atomicAdd(reinterpret_cast<unsigned long long *>(iscalars), (unsigned long long)s2[0]);
        atomicAdd(reinterpret_cast<unsigned long long *>(iscalars) + 2, (unsigned long long)s3[0]);
        atomicAdd(reinterpret_cast<unsigned long long *>(iscalars) + 3, (unsigned long long)s4[0]);
    }
}

// ---- per-sample vectors (exact int64) ----------------------------------------
//   V_i = sum_l R_l[g_il] = sum_l (a_l x_il + b_l m_il) - sum_l b_l     (R_l[g] = a_l g - b_l for valid g)
//   VEC_WLO[i] = low 20-bit limb sums (units 2^-fv), VEC_W[i] = upper limbs (units 2^-(fv-20));
//   the constant sum_l b_l goes to iscalars[2..3] (tables_kernel);
//   VEC_D[i] = sum_l qD_l m_il, VEC_D2[i] = sum_l qD2_l m_il  (missing-pair denominators).
// A block owns 32 word columns (512 samples) and SV_ROWS SNP rows; its four warps take every fourth
// row.  a_l x_il runs as one IMAD per 20-bit limb into int32 registers (512 rows x 2 x 2^20 < 2^31);
// the sparse m terms take a divergent slow path with 64-bit shared-memory atomics.
constexpr int SV_ROWS = 2048, SV_THREADS = 128;
template <int NL>
__global__ void __launch_bounds__(SV_THREADS)
sample_vec_kernel(const uint32_t *__restrict__ geno, const SnpStat *__restrict__ st, const int2 *__restrict__ coltab,
                  int64_t n_snp, int64_t n_samp, int64_t row_words, int64_t npad, int est, int bayesian, int frac_bits_d,
                  int frac_bits_v, long long *__restrict__ vec) {
    __shared__ int tA[GRAM_CHUNK][NL];
    __shared__ long long tB[GRAM_CHUNK], tD[GRAM_CHUNK], tD2[GRAM_CHUNK];
    __shared__ unsigned long long sv[4][512];      // VEC_W, VEC_WLO, VEC_D, VEC_D2 of the block's 512 samples
    for (int k = threadIdx.x; k < 4 * 512; k += blockDim.x) (&sv[0][0])[k] = 0ull;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t wc = (int64_t)blockIdx.x * 32 + lane;
    const bool live = wc < row_words;
    // padding samples (code 3 in every row) would send their lanes down the slow path for every SNP
    uint32_t real = 0;
    for (int k = 0; k < 16; k++)
        if (wc * 16 + k < n_samp) real |= 1u << (2 * k);
    int acc[16][NL];
#pragma unroll
    for (int k = 0; k < 16; k++)
#pragma unroll
        for (int j = 0; j < NL; j++) acc[k][j] = 0;
    const int64_t lb = (int64_t)blockIdx.y * SV_ROWS;
    for (int64_t l0 = lb; l0 < min(lb + (int64_t)SV_ROWS, n_snp); l0 += GRAM_CHUNK) {
        const int nl = (int)min((int64_t)GRAM_CHUNK, n_snp - l0);
        __syncthreads();
        for (int s = threadIdx.x; s < nl; s += blockDim.x) {
            SnpTables t;
            snp_tables(st[l0 + s], est, bayesian, coltab[l0 + s], 0, 0, frac_bits_d, frac_bits_v, t);
            long long qa = t.qa;
#pragma unroll
            for (int j = 0; j < NL; j++) {      // low limbs unsigned 20 bits, the top one signed
                tA[s][j] = j + 1 < NL ? (int)(qa & ((1ll << LIMB) - 1)) : (int)qa;
                qa >>= LIMB;
            }
            tB[s] = t.qb;
            tD[s] = t.qD;
            tD2[s] = t.qD2;
        }
        __syncthreads();
        const uint32_t *p = geno + l0 * row_words + (live ? wc : 0);
        for (int r = warp; r < nl; r += 4 * 4) {
            uint32_t wv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) wv[u] = (live && r + 4 * u < nl) ? __ldg(p + (int64_t)(r + 4 * u) * row_words) : 0u;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int row = r + 4 * u;
                if (row >= nl) break;
                const uint32_t x = wv[u];
                const uint32_t m3 = x & (x >> 1) & 0x55555555u;
                const uint32_t xv = x & ~(m3 * 3u);            // missing -> 0
                int a[NL];
#pragma unroll
                for (int j = 0; j < NL; j++) a[j] = tA[row][j];
#pragma unroll
                for (int k = 0; k < 16; k++) {
                    const int g = (int)((xv >> (2 * k)) & 3u);
#pragma unroll
                    for (int j = 0; j < NL; j++) acc[k][j] += g * a[j];
                }
                uint32_t mm = live ? (m3 & real) : 0u;
                while (mm) {                                    // sparse: b_l, d_l, d2_l of the missing genotypes
                    const int k = (__ffs(mm) - 1) >> 1;
                    mm &= mm - 1;
                    const int sidx = lane * 16 + k;
                    const long long qb = tB[row];
                    atomicAdd(&sv[1][sidx], (unsigned long long)(qb & ((1ll << LIMB) - 1)));
                    atomicAdd(&sv[0][sidx], (unsigned long long)(qb >> LIMB));
                    if (tD[row]) atomicAdd(&sv[2][sidx], (unsigned long long)tD[row]);
                    if (tD2[row]) atomicAdd(&sv[3][sidx], (unsigned long long)tD2[row]);
                }
            }
        }
    }
    // limb sums of this warp -> the block's per-sample sums
#pragma unroll
    for (int k = 0; k < 16; k++) {
        long long hi = 0;
#pragma unroll
        for (int j = NL - 1; j >= 1; j--) hi = (hi << LIMB) + (long long)acc[k][j];
        if (acc[k][0]) atomicAdd(&sv[1][lane * 16 + k], (unsigned long long)(long long)acc[k][0]);
        if (hi) atomicAdd(&sv[0][lane * 16 + k], (unsigned long long)hi);
    }
    __syncthreads();
    const int vmap[4] = {VEC_W, VEC_WLO, VEC_D, VEC_D2};
    for (int k = threadIdx.x; k < 4 * 512; k += blockDim.x) {
        const int v = k >> 9, sidx = k & 511;
        const int64_t i = (int64_t)blockIdx.x * 512 + sidx;
        const unsigned long long val = sv[v][sidx];
        if (val && i < npad) atomicAdd(reinterpret_cast<unsigned long long *>(vec) + (int64_t)vmap[v] * npad + i, val);
    }
}

__global__ void het_to_vec_kernel(const int *__restrict__ cnt, long long *__restrict__ vec, int64_t npad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npad) vec[VEC_HET * npad + i] = cnt[i];
}

// ---- helpers ---------------------------------------------------------------
// largest frac_bits such that |round(max_abs * 2^f)| fits K balanced base-256 digits
// (K digits cover [-128, 127] * (256^K - 1) / 255)
static int frac_cap(double max_abs, int K) {
    if (!(max_abs > 0)) return 60;
    double lim = 127.0 * (std::pow(256.0, K) - 1.0) / 255.0 - 1.0;
    return (int)std::floor(std::log2(lim / max_abs));
}
static int digits_for(double max_abs, int frac_bits) {
    for (int K = 1; K <= MAX_DIGITS; K++)
        if (frac_cap(max_abs, K) >= frac_bits) return K;
    fail("fixed-point format needs more than %d digits (max |value| %.3g, %d fractional bits)", MAX_DIGITS,
         max_abs, frac_bits);
}

// per-SNP column tables of the current genotypes / estimator
static void ensure_coltab(snprel_ctx *c, int est, int bayesian) {
    ensure_stats(c);
    if (c->coltab_version == c->geno_version && c->coltab_est == est && c->coltab_bayesian == bayesian) return;
    c->scr_coltab.alloc((size_t)c->snp_cap);
    c->scr_tabb.alloc((size_t)3 * c->snp_cap);
    c->scr_tabb.zero(c->stream);               // padding rows: all-zero tables
    if (c->n_snp > 0) {
        coltab_kernel<<<(unsigned)((c->n_snp + 255) / 256), 256, 0, c->stream>>>(
            c->stat.p, c->n_snp, est, bayesian, c->scr_coltab.p, c->scr_tabb.p, c->scr_tabb.p + c->snp_cap,
            c->scr_tabb.p + 2 * c->snp_cap);
        KERNEL_CHECK(c);
    }
    c->coltab_version = c->geno_version;
    c->coltab_est = est;
    c->coltab_bayesian = bayesian;
}

// scr_ew[npad], scr_cnt[2][npad], scr_chunk[nchunk]: this rank's per-sample error weights, het / missing
// counts and the per-chunk maxima of the error weight
static void sample_stats(snprel_ctx *c, int est = -1, int bayesian = 0) {   // est >= 0: also the diagonal bound (needs scr_plan of plan_kernel)
    const int64_t npad = c->n_samp_pad;
    const bool with_diag = est >= 0 && est != SNPREL_EST_KING_HOMO && c->n_snp > 0;
    c->scr_dg.alloc((size_t)npad);
    c->scr_dg.zero(c->stream);
    if (with_diag) {
        c->scr_tabf.alloc((size_t)c->snp_cap);
        diagtab_kernel<<<(unsigned)((c->n_snp + 255) / 256), 256, 0, c->stream>>>(c->stat.p, c->n_snp, est, bayesian, c->scr_plan.p,
                                                                                  c->scr_tabf.p);
        KERNEL_CHECK(c);
    }
    const int64_t nchunk = (c->snp_cap + GRAM_CHUNK - 1) / GRAM_CHUNK;
    c->scr_cnt.alloc((size_t)2 * npad);
    c->scr_cnt.zero(c->stream);
    c->scr_ew.alloc((size_t)npad);
    c->scr_ew.zero(c->stream);
    c->scr_sq.alloc((size_t)npad);
    c->scr_sq.zero(c->stream);
    c->scr_chunk.alloc((size_t)nchunk);
    c->scr_chunk.zero(c->stream);
    if (c->n_snp > 0) {
        const int64_t row_words = c->row_bytes / 4;
        dim3 grid((unsigned)((row_words + SS_THREADS - 1) / SS_THREADS), (unsigned)((c->n_snp + GRAM_CHUNK - 1) / GRAM_CHUNK));
        const uint32_t *g32 = reinterpret_cast<const uint32_t *>(c->geno2b.p);
        const uint32_t *tabA = c->scr_tabb.p + c->snp_cap, *tabS = c->scr_tabb.p + 2 * c->snp_cap;
        if (with_diag)      // the plan: error weight, counts, diagonal bound and the sum of squares
            sample_stats_kernel<true, true><<<grid, SS_THREADS, 0, c->stream>>>(g32, tabA, c->n_snp, row_words, npad, c->scr_ew.p, c->scr_cnt.p,
                                                                               c->scr_chunk.p, c->scr_tabf.p, c->scr_dg.p, tabS, c->scr_sq.p);
        else if (est >= 0)
            sample_stats_kernel<false, true><<<grid, SS_THREADS, 0, c->stream>>>(g32, tabA, c->n_snp, row_words, npad, c->scr_ew.p, c->scr_cnt.p,
                                                                                c->scr_chunk.p, nullptr, nullptr, tabS, c->scr_sq.p);
        else                // counts only
            sample_stats_kernel<false, false><<<grid, SS_THREADS, 0, c->stream>>>(g32, tabA, c->n_snp, row_words, npad, c->scr_ew.p, c->scr_cnt.p,
                                                                                 c->scr_chunk.p, nullptr, nullptr, nullptr, nullptr);
        KERNEL_CHECK(c);
    }
}

void grm_plan_local(snprel_ctx *c, int est, snprel_plan *plan) {
    if (!plan) fail("snprel_plan_local: NULL plan");
    {   // same genotypes, same estimator: the statistics do not depend on the row window
        const snprel_ctx::PlanCache &pc = c->plan_cache;
        if (pc.version == c->geno_version && pc.est == est && pc.bayesian == plan->bayesian) {
            const snprel_plan &s = pc.stats;
            plan->max_abs = s.max_abs;
            plan->max_abs_w = s.max_abs_w;
            plan->sum_bound = s.sum_bound;
            plan->total_missing = s.total_missing;
            plan->scale = s.scale;
            plan->err_weight = s.err_weight;
            plan->max_missing = s.max_missing;
            plan->n_snp = s.n_snp;
            plan->diag_bound = s.diag_bound;
            plan->sum_rest = s.sum_rest;
            plan->err_weight2 = s.err_weight2;
            return;
        }
    }
    ensure_coltab(c, est, plan->bayesian);
    const int64_t npad = c->n_samp_pad;
    DevBuf<double> &out = c->scr_plan;
    out.alloc(8);
    out.zero(c->stream);
    if (c->n_snp > 0) {
        int blocks = (int)std::min<int64_t>((c->n_snp + 255) / 256, 1024);
        plan_kernel<<<blocks, 256, 0, c->stream>>>(c->stat.p, c->scr_coltab.p, c->n_snp, c->n_samp, est,
                                                   plan->bayesian, out.p);
        KERNEL_CHECK(c);
    }
    sample_stats(c, est, plan->bayesian);
    double h[8];
    c->host_cnt.resize((size_t)npad);
    c->host_ew.resize((size_t)npad);
    std::vector<long long> hdg((size_t)npad);
    CUDA_CHECK(cudaMemcpyAsync(hdg.data(), c->scr_dg.p, (size_t)npad * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    const size_t nchunk = (size_t)((c->n_snp + GRAM_CHUNK - 1) / GRAM_CHUNK);
    std::vector<int> hchunk(nchunk);
    CUDA_CHECK(cudaMemcpyAsync(h, out.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->host_cnt.data(), c->scr_cnt.p + npad, (size_t)npad * sizeof(int),
                               cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->host_ew.data(), c->scr_ew.p, (size_t)npad * sizeof(long long),
                               cudaMemcpyDeviceToHost, c->stream));
    std::vector<long long> hsq((size_t)npad);
    CUDA_CHECK(cudaMemcpyAsync(hsq.data(), c->scr_sq.p, (size_t)npad * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    if (nchunk)
        CUDA_CHECK(cudaMemcpyAsync(hchunk.data(), c->scr_chunk.p, nchunk * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    long long ew = 0, mm = 0, dg = 0, sq = 0;
    for (int64_t i = 0; i < c->n_samp; i++) {
        ew = std::max(ew, c->host_ew[i]);
        mm = std::max<long long>(mm, c->host_cnt[i]);
        dg = std::max(dg, hdg[i]);
        sq = std::max(sq, hsq[i]);
    }
    plan->err_weight2 = (est == SNPREL_EST_KING_HOMO) ? 0.0 : (double)sq * (double)SQ_UNIT;   // >= max_i sum_l B_l[g_il]^2
    // X = max_i sum_l ceil(w (g - mu)^2 / c) c  >=  max_i C_ii   (0: not measured)
    plan->diag_bound = (est == SNPREL_EST_KING_HOMO) ? 0.0 : (double)dg * (h[5] / (double)AUX_MAX);
    plan->sum_rest = h[6];
    c->chunk_bound.assign(hchunk.begin(), hchunk.end());
    plan->max_abs = h[0];
    plan->max_abs_w = h[4];
    plan->sum_bound = h[1];
    plan->total_missing = (int64_t)h[2];
    plan->scale = (est == SNPREL_GRM_EIGENSTRAT) ? h[3] / (double)std::max<int64_t>(c->n_samp - 1, 1)
                  : (est == SNPREL_GRM_GCTA)     ? 2.0 * h[3]
                                                 : h[3];
    plan->err_weight = (est == SNPREL_EST_KING_HOMO) ? 0.0 : (double)ew;
    plan->max_missing = mm;
    plan->n_snp = c->n_snp;
    c->plan_cache.version = c->geno_version;
    c->plan_cache.est = est;
    c->plan_cache.bayesian = plan->bayesian;
    c->plan_cache.stats = *plan;
}

// Choose the fixed-point formats from the (global) plan statistics.
//   numerator plane: quantisation error of an entry
//        <= 2^-(f+1) * err_weight  (T table against the integer column table B)
//         + 2^-(fw+1) * max_missing (R table against the m channel),   wanted <= tol * scale
//   denominator plane (EIGMIX / KING-homo): error <= 2^-(fd+1) * 2 max_missing, wanted <= tol * scale
// Items run two passes at a time; a table with an odd digit count sends its last digit through
// 256 x 512 items (one pass, two column tiles) at the same cost per pass.
static double launch_cost(int n) { return (double)n; }

// round_mode 1 (snprel_set_rounding): the T table is rounded at random, so for a fixed pair (i, j)
// the error sum_l e_l[g_il] B_l[g_jl] is a sum of independent zero-mean terms, the l-th confined to an
// interval of width |B_l[g_jl]| 2^-f, and by Hoeffding it exceeds 2^-f sqrt(1/2 sum_l B^2 ln(2/delta)) with
// probability < delta; delta is 1e-12 divided by the number of pairs (union bound over the whole matrix).
// sum_l B^2 is bounded per column sample: measured (err_weight2, sample_stats_kernel), else <= 127 sum |B|.
static double u_table_error(const snprel_plan &plan, int fa, int round_mode, double n_samp) {
    if (round_mode != 1) return std::ldexp(plan.err_weight, -(fa + 1));
    const double pairs = std::max(1.0, 0.5 * n_samp * (n_samp + 1.0));
    double s2 = 127.0 * plan.err_weight;
    if (plan.err_weight2 > 0) s2 = std::min(s2, plan.err_weight2);
    return std::ldexp(std::sqrt(0.5 * s2 * std::log(2.0 * pairs / 1e-12)), -fa);
}

// one rounding rule (0 nearest / 1 randomised); true: the error bound meets the tolerance
static bool choose_format_one(int est, snprel_plan &plan, int &nU, int &nW, int &nD, int round_mode,
                              double n_samp) {
    bool feasible = true;
    const double tol = (plan.tol > 0 ? plan.tol : 1e-10) * 0.9;   // 10 % left for the per-sample vector and float64 rounding in the epilogue
    const bool homo = est == SNPREL_EST_KING_HOMO;
    const bool any_missing = plan.total_missing > 0;
    // int64 plane headroom.  Sums wrap modulo 2^64 harmlessly on the way (atomics, reductions over GPUs): only
    // the FINAL entries must fit.  Worst case: sum_bound = sum_l max|T_l| max|B_l| + rest.  Measured
    // (diagtab_kernel / sample_stats_kernel, Cauchy-Schwarz + AM-GM): 1.5 X + rest with X >= max_i C_ii.
    double range_bound = plan.sum_bound;
    if (plan.diag_bound > 0) range_bound = std::min(range_bound, 1.5 * plan.diag_bound + plan.sum_rest);
    const int head = 61 - (int)std::ceil(std::log2(std::max(range_bound, 1.0)));
    if (head < 16) fail("fixed-point accumulator cannot hold this data set (sum bound %.3g)", plan.sum_bound);
    double scale = plan.scale;
    if (est == SNPREL_GRM_GCTA) scale -= 4.0 * (double)plan.max_missing;   // 2 (nLocus - D_ij), D_ij <= 2 max_missing
    if (est == SNPREL_GRM_EIGMIX) scale -= 2.0 * (double)plan.max_missing;
    // the subtraction is a worst case (every missing SNP at full weight for both samples);
    // never let it shrink the normaliser below 5 % of its complete-data value
    scale = std::max(scale, 0.05 * plan.scale);
    const double budget = tol * std::max(scale, 1e-300);
    const int fmax = head;
    nU = nW = 0;
    if (!homo) {
        if (plan.frac_bits >= 0) {           // caller-fixed format
            nU = digits_for(plan.max_abs, plan.frac_bits);
            if (any_missing) {
                if (plan.frac_bits_w < 0 || plan.frac_bits_w > plan.frac_bits) plan.frac_bits_w = plan.frac_bits;
                nW = digits_for(plan.max_abs_w, plan.frac_bits_w);
            }
        } else {
            double best_cost = 1e30, best_err = 1e300;
            int bU = MAX_DIGITS, bW = any_missing ? MAX_DIGITS : 0;
            feasible = false;
            for (int a = 1; a <= MAX_DIGITS; a++) {
                int fa = std::min(frac_cap(plan.max_abs, a), fmax);
                if (fa < 8) continue;
                double ea = u_table_error(plan, fa, round_mode, n_samp);
                for (int b = any_missing ? 1 : 0; b <= (any_missing ? MAX_DIGITS : 0); b++) {
                    int fb = b ? std::min(frac_cap(plan.max_abs_w, b), fa) : fa;
                    if (b && fb < 8) continue;
                    double err = ea + (b ? std::ldexp((double)plan.max_missing, -(fb + 1)) : 0.0);
                    bool ok = err <= budget;
                    double cost = launch_cost(a) + launch_cost(b);
                    if ((ok && (!feasible || cost < best_cost - 1e-9 || (std::fabs(cost - best_cost) < 1e-9 && err < best_err))) ||
                        (!ok && !feasible && err < best_err)) {
                        best_cost = cost;
                        best_err = err;
                        bU = a;
                        bW = b;
                        feasible = feasible || ok;
                    }
                }
            }
            nU = bU;
            nW = bW;
            plan.frac_bits = std::min(frac_cap(plan.max_abs, nU), fmax);
            plan.frac_bits_w = nW ? std::min(frac_cap(plan.max_abs_w, nW), plan.frac_bits) : plan.frac_bits;
        }
    } else {
        plan.frac_bits = 0;
        plan.frac_bits_w = 0;
    }
    plan.digits = nU;
    plan.digits_w = nW;
    // the per-sample vector: 1.5 n_snp 2^-fv (rounding of a_l x and b_l, x <= 2) <= 5 % of the budget
    {
        const double need = 1.5 * (double)std::max<int64_t>(plan.n_snp, 1) / (0.05 * budget / 0.9);
        int fv = (int)std::ceil(std::log2(std::max(need, 2.0)));
        const int lim = 57 - (int)std::ceil(std::log2(std::max(2.0 * plan.max_abs_w, 1e-300)));   // |qa|, |qb| < 2^58: three limbs
        plan.frac_bits_v = std::max(8, std::min(std::min(fv, lim), 60));
    }
    nD = 0;
    plan.frac_bits_d = std::max(plan.frac_bits_d, 0);
    if (any_missing) {
        if (est == SNPREL_GRM_GCTA) {
            nD = 1;               // d in {0,1}: exact integers
            plan.frac_bits_d = 0;
        } else if (est == SNPREL_GRM_EIGMIX || homo) {
            double dmax = homo ? 0.25 : 1.0;
            int fd_req = 24;
            if (scale > 0)
                fd_req = (int)std::ceil(std::log2(std::max(2.0 * (double)plan.max_missing, 1.0) / budget)) - 1;
            fd_req = std::max(16, std::min(fd_req, head));
            nD = digits_for(dmax, fd_req);
            plan.frac_bits_d = std::min(std::min(frac_cap(dmax, nD), head), 50);
        }
    }
    plan.digits_d = nD;
    plan.rounding = (round_mode == 1 && !homo) ? 1 : 0;
    return feasible;
}

constexpr double AUTO_ROUNDING_MIN_WORK = 68719476736.0;   // 2^36 pair-SNPs per pass
// round_mode as in snprel_set_rounding.  A caller-fixed format (frac_bits >= 0) brings its rounding
// along in plan.rounding; otherwise mode 2 takes randomised rounding only where its (probabilistic)
// bound is met with FEWER tensor passes than the worst-case bound needs.
static void choose_format(int est, snprel_plan &plan, int &nU, int &nW, int &nD, int round_mode, double n_samp) {
    if (plan.frac_bits >= 0) {
        choose_format_one(est, plan, nU, nW, nD, plan.rounding ? 1 : 0, n_samp);
        return;
    }
    // mode 2 on a small problem (n_samp^2 n_snp < 2^36: a tensor pass costs microseconds): nothing to gain, and
    // round-to-nearest leaves the larger margin to the tolerance
    const bool small = n_samp * n_samp * (double)std::max<int64_t>(plan.n_snp, 0) < AUTO_ROUNDING_MIN_WORK;
    if (round_mode != 2 || est == SNPREL_EST_KING_HOMO || small) {
        choose_format_one(est, plan, nU, nW, nD, round_mode == 1 ? 1 : 0, n_samp);
        return;
    }
    snprel_plan pr = plan;
    int rU = 0, rW = 0, rD = 0;
    const bool ok_near = choose_format_one(est, plan, nU, nW, nD, 0, n_samp);
    const bool ok_rand = choose_format_one(est, pr, rU, rW, rD, 1, n_samp);
    if (ok_rand && (!ok_near || rU + rW + rD < nU + nW + nD)) {
        plan = pr;
        nU = rU;
        nW = rW;
        nD = rD;
    }
}

void grm_plan_format(int est, snprel_plan *plan, int round_mode, int64_t n_samp) {
    if (!plan) fail("snprel_plan_format: NULL plan");
    if (est == SNPREL_GRM_CORR) est = SNPREL_GRM_GCTA;
    if (!((est >= SNPREL_GRM_EIGENSTRAT && est <= SNPREL_GRM_EIGMIX) || est == SNPREL_EST_KING_HOMO))
        fail("snprel_plan_format: not a covariance estimator (%d)", est);
    if (round_mode < 0 || round_mode > 2) fail("snprel_plan_format: rounding mode 0, 1 or 2");
    int nU = 0, nW = 0, nD = 0;
    choose_format(est, *plan, nU, nW, nD, round_mode, (double)n_samp);
}

void grm_accumulate(snprel_ctx *c, int est, const snprel_plan *plan_in) {
    if (est == SNPREL_GRM_CORR) est = SNPREL_GRM_GCTA;
    snprel_plan plan = *plan_in;
    ensure_coltab(c, est, plan.bayesian);
    const bool homo = est == SNPREL_EST_KING_HOMO;
    int nU = 0, nW = 0, nD = 0;
    choose_format(est, plan, nU, nW, nD, c->round_mode, (double)c->n_samp);
    const int f = plan.frac_bits, fw = plan.frac_bits_w, fd = plan.frac_bits_d, fv = plan.frac_bits_v;
    int nD2 = homo ? nD : 0;
    const int npass = nU + nW + nD + nD2;
    const int64_t cap = c->snp_cap, npad = c->n_samp_pad;

    // Digit tables, per-sample vectors and global scalars depend on the genotypes and the
    // fixed-point format only, not on the row window: a tiled N x N run builds them once.
    snprel_ctx::PrepCache &pc = c->prep_cache;
    const bool prep_hit = pc.version == c->geno_version && pc.est == est && pc.bayesian == plan.bayesian &&
                          pc.rounding == plan.rounding && (!plan.rounding || pc.origin == c->snp_origin) && pc.fv == fv &&
                          pc.f == f && pc.fw == fw && pc.fd == fd && pc.nU == nU && pc.nW == nW && pc.nD == nD &&
                          pc.nD2 == nD2 && c->scr_tab.p && c->samp_sum.p && c->scr_cnt.p;
    DevBuf<uint32_t> &tab = c->scr_tab;
    if (!prep_hit) {
        // scr_cnt was summed over the ranks together with an earlier format's vectors while the
        // cached plan statistics kept it from being rebuilt: restore this rank's own counts
        if ((pc.reduced && pc.version == c->geno_version) || !c->scr_cnt.p || c->plan_cache.version != c->geno_version)
            sample_stats(c);      // (counts only: the diagonal bound lives in the plan)
        pc.version = 0;
        tab.alloc((size_t)std::max(npass, 1) * cap);
        tab.zero(c->stream);
        c->scalars.alloc(4);
        c->scalars.zero(c->stream);
        c->iscalars.alloc(4);
        c->iscalars.zero(c->stream);
        DevBuf<int> &ovf = c->scr_flags;
        ovf.alloc(2);
        ovf.zero(c->stream);
        const unsigned tblocks = (unsigned)((c->n_snp + 255) / 256);
        c->scr_part.alloc((size_t)std::max(tblocks, 1u) * 2);
        if (c->n_snp > 0) {
            // (the draws are keyed by the GLOBAL SNP index, origin + local row: one draw per SNP of the data
            //  set however it is sharded, independent across SNPs)
            tables_kernel<<<tblocks, 256, 0, c->stream>>>(c->stat.p, c->scr_coltab.p, c->n_snp, cap, est, plan.bayesian, f, fw, fd, fv, nU,
                                                          nW, nD, nD2, tab.p, c->scr_part.p, c->iscalars.p, ovf.p,
                                                          (plan.rounding && !homo) ? DITHER_SEED : 0ull, (long long)c->snp_origin);
            KERNEL_CHECK(c);
        }
        int hovf = 0;
        std::vector<double> part((size_t)tblocks * 2);
        CUDA_CHECK(cudaMemcpyAsync(&hovf, ovf.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        if (tblocks)
            CUDA_CHECK(cudaMemcpyAsync(part.data(), c->scr_part.p, part.size() * sizeof(double),
                                       cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        double hsc[4] = {0, 0, 0, 0};
        for (unsigned b = 0; b < tblocks; b++) {
            hsc[0] += part[2 * b];
            hsc[1] += part[2 * b + 1];
        }
        CUDA_CHECK(cudaMemcpyAsync(c->scalars.p, hsc, sizeof(hsc), cudaMemcpyHostToDevice, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));   // hsc lives on this stack frame
        if (hovf) fail("internal: fixed-point digit overflow (table %d, frac_bits %d/%d/%d)", hovf, f, fw, fd);

        // per-sample vectors
        c->samp_sum.alloc((size_t)NVEC * npad);
        c->samp_sum.zero(c->stream);
        c->samp_vecs = NVEC;
        if (c->n_snp > 0) {
            const int64_t row_words = c->row_bytes / 4;
            dim3 grid((unsigned)((row_words + 31) / 32), (unsigned)((c->n_snp + SV_ROWS - 1) / SV_ROWS));
            // |qa| < 2^(fv + 1) max|R|: two 20-bit limbs up to 2^39, three up to 2^59
            const double qbits = (double)fv + 1.0 + std::log2(std::max(plan.max_abs_w, 1e-300));
            const uint32_t *g32 = reinterpret_cast<const uint32_t *>(c->geno2b.p);
            if (qbits < 38.0)
                sample_vec_kernel<2><<<grid, SV_THREADS, 0, c->stream>>>(g32, c->stat.p, c->scr_coltab.p, c->n_snp, c->n_samp, row_words,
                                                                        npad, est, plan.bayesian, fd, fv, c->samp_sum.p);
            else
                sample_vec_kernel<3><<<grid, SV_THREADS, 0, c->stream>>>(g32, c->stat.p, c->scr_coltab.p, c->n_snp, c->n_samp, row_words,
                                                                        npad, est, plan.bayesian, fd, fv, c->samp_sum.p);
            KERNEL_CHECK(c);
            het_to_vec_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, c->stream>>>(c->scr_cnt.p, c->samp_sum.p, npad);
            KERNEL_CHECK(c);
        }
        pc.version = c->geno_version;
        pc.est = est;
        pc.bayesian = plan.bayesian;
        pc.f = f;
        pc.fw = fw;
        pc.fd = fd;
        pc.fv = fv;
        pc.nU = nU;
        pc.nW = nW;
        pc.nD = nD;
        pc.nD2 = nD2;
        pc.rounding = plan.rounding;
        pc.origin = c->snp_origin;
        pc.reduced = false;
    }

    // Gram planes: 0 = numerator, 1 = missing-pair denominator (or KING-homo d), 2 = KING-homo d2
    const int nplanes = homo ? 2 : ((nD > 0) ? 2 : 1);
    c->acc.alloc((size_t)nplanes * row_window(c).rows * npad);
    c->acc.zero(c->stream);
    c->acc_planes = nplanes;

    // int32 headroom of the main passes: |digit| <= 128 against the measured per-chunk column weights
    const uint32_t *tabB = c->scr_tabb.p;
    const uint32_t *tabM = gram_const_table(c, TABB_M);
    const std::vector<int64_t> *cb = c->chunk_bound.size() == (size_t)((c->n_snp + GRAM_CHUNK - 1) / GRAM_CHUNK)
                                         ? &c->chunk_bound : nullptr;
    std::vector<GramPass> passes;
    int pass = 0;
    for (int k = 0; k < nU; k++, pass++) passes.push_back({tab.p + (int64_t)pass * cap, tabB, 0, 8 * k, 127, cb});
    for (int k = 0; k < nW; k++, pass++) passes.push_back({tab.p + (int64_t)pass * cap, tabM, 0, 8 * k + (f - fw), 1, nullptr});
    for (int k = 0; k < nD; k++, pass++)
        passes.push_back({tab.p + (int64_t)pass * cap, tabM, homo ? 0 : 1, 8 * k, 1, nullptr});
    for (int k = 0; k < nD2; k++, pass++) passes.push_back({tab.p + (int64_t)pass * cap, tabM, 1, 8 * k, 1, nullptr});

    c->hot_launches = 0;
    CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
    gram_tc_run(c, passes.data(), (int)passes.size(), c->acc.p, true);
    CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->hot_ms = ms;
    c->hot_units = 0.5 * (double)c->n_samp * (double)c->n_samp * (double)c->n_snp;

    c->plan = plan;
    c->accum_win_r0 = row_window(c).r0;
    c->accum_win_rows = row_window(c).rows;
    c->accum_bayesian = plan.bayesian ? 1 : 0;
    c->accum_est = est;
    c->accum_reduced = false;
    c->reduce_list.clear();
    c->reduce_list.push_back({c->acc.p, (int64_t)nplanes * row_window(c).rows * npad, 0, npad, row_window(c).rows, row_window(c).r0});
    if (!pc.reduced) {   // (kept from an earlier window: already summed over the ranks)
        c->reduce_list.push_back({c->samp_sum.p, (int64_t)c->samp_sum.n, 0});
        c->reduce_list.push_back({c->scalars.p, (int64_t)c->scalars.n, 2});
        c->reduce_list.push_back({c->iscalars.p, (int64_t)c->iscalars.n, 0});
        c->reduce_list.push_back({c->scr_cnt.p, (int64_t)c->scr_cnt.n, 1});   // per-sample het / missing counts
    }
}

// ---------------------------------------------------------------------------
// Streamed accumulate: the covariance path over host-to-device copies that are still in flight
// (snprel_geno_push_2b_async).  The reference double-buffers its block reader against the compute
// threads (CGenoReadBySNP, src/dGenGWAS.cpp:1298-1324); here the 2-bit matrix arrives in chunks of
// STREAM_CHUNK SNP rows on a copy stream and every chunk is consumed as soon as its event has fired:
// per-SNP statistics, column tables, plan statistics, per-sample statistics, digit tables, per-sample
// vectors and the tensor passes of that SNP range -- all of them sums over SNPs that accumulate
// across chunks.  The one thing that needs the whole matrix is the fixed-point FORMAT (it is chosen
// from global statistics), so it is chosen speculatively from the first chunk's statistics
// extrapolated to the full SNP count with safety margins, and verified at the end against the true
// statistics (digit overflow flag, int64 headroom, error bound <= tol): on failure everything is
// recomputed by the ordinary path on the now resident data.  Whole matrix (no row window) only.
// ---------------------------------------------------------------------------
static void prep_stats_range(snprel_ctx *c, int est, int bayesian, int64_t l0, int64_t l1, int64_t l1_pad) {
    // per-SNP statistics (padding rows of the last chunk included: they read as all-missing)
    snp_stats_range(c, l0, l1_pad - l0);
    const int64_t n = l1 - l0;
    if (n <= 0) return;
    coltab_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->stat.p + l0, n, est, bayesian, c->scr_coltab.p + l0,
                                                                   c->scr_tabb.p + l0, c->scr_tabb.p + c->snp_cap + l0,
                                                                   c->scr_tabb.p + 2 * c->snp_cap + l0);
    KERNEL_CHECK(c);
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, 1024);
    plan_kernel<<<blocks, 256, 0, c->stream>>>(c->stat.p + l0, c->scr_coltab.p + l0, n, c->n_samp, est, bayesian, c->scr_plan.p);
    KERNEL_CHECK(c);
    const int64_t row_words = c->row_bytes / 4;
    dim3 grid((unsigned)((row_words + SS_THREADS - 1) / SS_THREADS), (unsigned)((n + GRAM_CHUNK - 1) / GRAM_CHUNK));
    sample_stats_kernel<false, true><<<grid, SS_THREADS, 0, c->stream>>>(reinterpret_cast<const uint32_t *>(c->geno2b.p) + l0 * row_words,
                                                           c->scr_tabb.p + c->snp_cap + l0, n, row_words, c->n_samp_pad,
                                                           c->scr_ew.p, c->scr_cnt.p, c->scr_chunk.p + l0 / GRAM_CHUNK, nullptr, nullptr,
                                                           c->scr_tabb.p + 2 * c->snp_cap + l0, c->scr_sq.p);
    KERNEL_CHECK(c);
}

// plan statistics accumulated so far -> host
static void read_plan_stats(snprel_ctx *c, int est, snprel_plan &plan, int64_t n_snp_seen) {
    const int64_t npad = c->n_samp_pad;
    double h[8];
    c->host_cnt.resize((size_t)npad);
    c->host_ew.resize((size_t)npad);
    CUDA_CHECK(cudaMemcpyAsync(h, c->scr_plan.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->host_cnt.data(), c->scr_cnt.p + npad, (size_t)npad * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->host_ew.data(), c->scr_ew.p, (size_t)npad * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    std::vector<long long> hsq((size_t)npad);
    CUDA_CHECK(cudaMemcpyAsync(hsq.data(), c->scr_sq.p, (size_t)npad * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    long long ew = 0, mm = 0, sq = 0;
    for (int64_t i = 0; i < c->n_samp; i++) {
        ew = std::max(ew, c->host_ew[i]);
        mm = std::max<long long>(mm, c->host_cnt[i]);
        sq = std::max(sq, hsq[i]);
    }
    plan.err_weight2 = (double)sq * (double)SQ_UNIT;
    plan.max_abs = h[0];
    plan.max_abs_w = h[4];
    plan.sum_bound = h[1];
    plan.total_missing = (int64_t)h[2];
    plan.scale = (est == SNPREL_GRM_EIGENSTRAT) ? h[3] / (double)std::max<int64_t>(c->n_samp - 1, 1)
                 : (est == SNPREL_GRM_GCTA)     ? 2.0 * h[3]
                                                : h[3];
    plan.err_weight = (double)ew;
    plan.max_missing = mm;
    plan.n_snp = n_snp_seen;
    plan.diag_bound = 0;          // (not measured chunk by chunk: the worst-case headroom is used)
    plan.sum_rest = h[6];
}

// true: accumulated (c->acc etc. hold the result, caches set); false: the caller must take the ordinary path
static bool grm_accumulate_streamed(snprel_ctx *c, int est, int bayesian) {
    if (est == SNPREL_GRM_CORR) est = SNPREL_GRM_GCTA;
    if (c->pending.empty() || !full_window(c) || (c->debug_flags & 2u) || c->n_snp <= 0) return false;
    std::vector<snprel_ctx::PendingCopy> chunks = c->pending;
    for (size_t k = 0; k < chunks.size(); k++) {   // contiguous, chunk aligned, up to the last row
        const int64_t expect = k ? chunks[k - 1].l1 : chunks[0].l0;
        if (chunks[k].l0 != expect || (chunks[k].l0 % GRAM_CHUNK)) return false;
    }
    if (chunks.back().l1 != c->n_snp) return false;
    const int64_t m = c->n_snp, cap = c->snp_cap, npad = c->n_samp_pad, first_l0 = chunks[0].l0;
    if (chunks.back().l1 - first_l0 < 2 * STREAM_CHUNK) return false;      // nothing to overlap
    CUDA_CHECK(cudaEventRecord(c->evs0, c->stream));
    geno_pad_tail(c);

    // accumulators of the statistics
    c->scr_coltab.alloc((size_t)cap);
    c->scr_tabb.alloc((size_t)3 * cap);
    c->scr_tabb.zero(c->stream);
    c->scr_sq.alloc((size_t)npad);
    c->scr_sq.zero(c->stream);
    c->scr_plan.alloc(8);
    c->scr_plan.zero(c->stream);
    c->scr_cnt.alloc((size_t)2 * npad);
    c->scr_cnt.zero(c->stream);
    c->scr_ew.alloc((size_t)npad);
    c->scr_ew.zero(c->stream);
    const int64_t nchunk = (cap + GRAM_CHUNK - 1) / GRAM_CHUNK;
    c->scr_chunk.alloc((size_t)nchunk);
    c->scr_chunk.zero(c->stream);

    // ---- first range: rows that were already resident plus the first chunk in flight -> speculative format
    const int64_t m_pad = round_up(m, SNP_PAD);
    auto pad_end = [&](int64_t l1) { return l1 == m ? m_pad : l1; };
    CUDA_CHECK(cudaStreamWaitEvent(c->stream, chunks[0].ev, 0));
    geno_fix_chunk_padding(c, chunks[0]);
    c->pending[0].consumed = true;
    const int64_t a1 = chunks[0].l1;
    prep_stats_range(c, est, bayesian, 0, a1, pad_end(a1));
    snprel_plan plan{};
    plan.frac_bits = plan.frac_bits_w = plan.frac_bits_d = -1;
    plan.bayesian = bayesian;
    read_plan_stats(c, est, plan, a1);
    {
        const double fac = (double)m / (double)a1;
        // the row tables are range-equalised by construction (coltab_kernel: s follows |U|), so a later SNP
        // exceeds the first 131k SNPs' maximum only marginally; a wrong guess is caught by the overflow flag
        plan.max_abs *= 1.25;
        plan.max_abs_w *= 1.25;
        plan.sum_bound *= fac * 1.03;
        plan.err_weight *= fac * 1.08;
        plan.err_weight2 *= fac * 1.08;
        plan.max_missing = (int64_t)std::ceil((double)plan.max_missing * fac * 1.15 + 16.0);
        plan.total_missing = (int64_t)((double)plan.total_missing * fac) + (m > a1 ? 1 : 0);
        plan.scale *= fac * 0.95;                  // a LOWER bound of the normaliser
        plan.sum_rest *= fac * 1.03;
        plan.n_snp = m;
    }
    int nU = 0, nW = 0, nD = 0;
    choose_format(est, plan, nU, nW, nD, c->round_mode, (double)c->n_samp);
    if (plan.total_missing > 0 && nW == 0) return false;
    const int f = plan.frac_bits, fw = plan.frac_bits_w, fd = plan.frac_bits_d, fv = plan.frac_bits_v;
    const int npass = nU + nW + nD;

    DevBuf<uint32_t> &tab = c->scr_tab;
    tab.alloc((size_t)std::max(npass, 1) * cap);
    tab.zero(c->stream);
    c->scalars.alloc(4);
    c->scalars.zero(c->stream);
    c->iscalars.alloc(4);
    c->iscalars.zero(c->stream);
    c->scr_flags.alloc(2);
    c->scr_flags.zero(c->stream);
    const int64_t tblocks_max = (m + 255) / 256 + (int64_t)chunks.size() + 1;     // (every range rounds its block count up)
    c->scr_part.alloc((size_t)tblocks_max * 2);
    CUDA_CHECK(cudaMemsetAsync(c->scr_part.p, 0, (size_t)tblocks_max * 2 * sizeof(double), c->stream));
    c->samp_sum.alloc((size_t)NVEC * npad);
    c->samp_sum.zero(c->stream);
    c->samp_vecs = NVEC;
    const int nplanes = (nD > 0) ? 2 : 1;
    c->acc.alloc((size_t)nplanes * npad * npad);
    c->acc.zero(c->stream);
    c->acc_planes = nplanes;

    const uint32_t *tabB = c->scr_tabb.p;
    const uint32_t *tabM = gram_const_table(c, TABB_M);
    std::vector<GramPass> passes;
    {
        int pass = 0;
        // |B| <= 127 and at most STREAM_CHUNK SNPs per item: 127 x 131072 x 128 < 2^31, no measured bound needed
        for (int k = 0; k < nU; k++, pass++) passes.push_back({tab.p + (int64_t)pass * cap, tabB, 0, 8 * k, 127, nullptr});
        for (int k = 0; k < nW; k++, pass++) passes.push_back({tab.p + (int64_t)pass * cap, tabM, 0, 8 * k + (f - fw), 1, nullptr});
        for (int k = 0; k < nD; k++, pass++) passes.push_back({tab.p + (int64_t)pass * cap, tabM, 1, 8 * k, 1, nullptr});
    }
    const double qbits = (double)fv + 1.0 + std::log2(std::max(plan.max_abs_w, 1e-300));
    const int64_t row_words = c->row_bytes / 4;
    const uint32_t *g32 = reinterpret_cast<const uint32_t *>(c->geno2b.p);
    int64_t tblock0 = 0;
    c->hot_launches = 0;
    CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
    auto consume = [&](int64_t l0, int64_t l1) {
        const int64_t n = l1 - l0;
        const unsigned tb = (unsigned)((n + 255) / 256);
        tables_kernel<<<tb, 256, 0, c->stream>>>(c->stat.p + l0, c->scr_coltab.p + l0, n, cap, est, bayesian, f, fw, fd, fv, nU, nW, nD, 0,
                                                 tab.p + l0, c->scr_part.p + 2 * tblock0, c->iscalars.p, c->scr_flags.p,
                                                 plan.rounding ? DITHER_SEED : 0ull, (long long)(c->snp_origin + l0));
        KERNEL_CHECK(c);
        tblock0 += tb;
        dim3 grid((unsigned)((row_words + 31) / 32), (unsigned)((n + SV_ROWS - 1) / SV_ROWS));
        if (qbits < 38.0)
            sample_vec_kernel<2><<<grid, SV_THREADS, 0, c->stream>>>(g32 + l0 * row_words, c->stat.p + l0, c->scr_coltab.p + l0, n, c->n_samp,
                                                                    row_words, npad, est, bayesian, fd, fv, c->samp_sum.p);
        else
            sample_vec_kernel<3><<<grid, SV_THREADS, 0, c->stream>>>(g32 + l0 * row_words, c->stat.p + l0, c->scr_coltab.p + l0, n, c->n_samp,
                                                                    row_words, npad, est, bayesian, fd, fv, c->samp_sum.p);
        KERNEL_CHECK(c);
        gram_tc_run(c, passes.data(), (int)passes.size(), c->acc.p, true, l0, l1, false);
    };
    // Pacing: the host stays exactly one range ahead of the device.  `gate` fires when the previous
    // range's tensor passes have been reached in the stream; the next range then takes EVERY chunk that
    // has arrived by that time (the copies run five times faster than the tensor passes, so after the
    // first two single-chunk ranges everything left is usually one launch).
    cudaEvent_t gate;
    CUDA_CHECK(cudaEventCreateWithFlags(&gate, cudaEventDisableTiming));
    CUDA_CHECK(cudaEventRecord(gate, c->stream));
    consume(0, a1);
    int nranges = 1;
    for (size_t k = 1; k < chunks.size();) {
        cudaEventSynchronize(gate);
        cudaEventSynchronize(chunks[k].ev);
        size_t k2 = k + 1;
        while (k2 < chunks.size() && cudaEventQuery(chunks[k2].ev) == cudaSuccess) k2++;
        for (size_t j = k; j < k2; j++) {
            CUDA_CHECK(cudaStreamWaitEvent(c->stream, chunks[j].ev, 0));
            geno_fix_chunk_padding(c, chunks[j]);
            c->pending[j].consumed = true;
        }
        const int64_t l0 = chunks[k].l0, l1 = chunks[k2 - 1].l1;
        prep_stats_range(c, est, bayesian, l0, l1, pad_end(l1));
        CUDA_CHECK(cudaEventRecord(gate, c->stream));
        consume(l0, l1);
        nranges++;
        k = k2;
    }
    cudaEventDestroy(gate);
    c->stream_ranges = nranges;
    CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
    het_to_vec_kernel<<<(unsigned)((npad + 255) / 256), 256, 0, c->stream>>>(c->scr_cnt.p, c->samp_sum.p, npad);
    KERNEL_CHECK(c);
    gram_tc_check(c);           // waits for everything queued
    geno_wait(c);               // the copies are done (their events were waited on): drop them

    // ---- verification against the true statistics
    snprel_plan truth{};
    truth.frac_bits = truth.frac_bits_w = truth.frac_bits_d = -1;
    truth.bayesian = bayesian;
    read_plan_stats(c, est, truth, m);
    int hovf = 0;
    const unsigned tblocks = (unsigned)tblock0;
    std::vector<double> part((size_t)tblocks * 2);
    CUDA_CHECK(cudaMemcpyAsync(&hovf, c->scr_flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(part.data(), c->scr_part.p, part.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    std::vector<int> hchunk((size_t)((m + GRAM_CHUNK - 1) / GRAM_CHUNK));
    CUDA_CHECK(cudaMemcpyAsync(hchunk.data(), c->scr_chunk.p, hchunk.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    bool ok = hovf == 0;
    {
        // the format that was used must satisfy the ordinary criterion under the true statistics
        const double tol = (truth.tol > 0 ? truth.tol : 1e-10) * 0.9;
        double scale = truth.scale;
        if (est == SNPREL_GRM_GCTA) scale -= 4.0 * (double)truth.max_missing;
        if (est == SNPREL_GRM_EIGMIX) scale -= 2.0 * (double)truth.max_missing;
        scale = std::max(scale, 0.05 * truth.scale);
        const double budget = tol * std::max(scale, 1e-300);
        const int head = 61 - (int)std::ceil(std::log2(std::max(truth.sum_bound, 1.0)));
        double err = u_table_error(truth, f, plan.rounding, (double)c->n_samp);
        if (truth.total_missing > 0) err += nW > 0 ? std::ldexp((double)truth.max_missing, -(fw + 1)) : 1e300;
        ok = ok && err <= budget && f <= head;
        if (truth.total_missing > 0 && est == SNPREL_GRM_GCTA) ok = ok && nD == 1;
        if (truth.total_missing > 0 && est == SNPREL_GRM_EIGMIX)
            ok = ok && nD > 0 && std::ldexp(2.0 * (double)truth.max_missing, -(fd + 1)) <= budget;
        // the per-sample vector's format
        ok = ok && 1.5 * (double)m * std::ldexp(1.0, -fv) <= 0.05 * budget / 0.9 * 1.0001;
    }
    if (!ok) {
        c->stream_fallbacks++;
        c->plan_cache.version = 0;
        c->prep_cache.version = 0;
        c->stat_valid = false;
        c->coltab_version = 0;
        return false;
    }
    double hsc[4] = {0, 0, 0, 0};
    for (unsigned b = 0; b < tblocks; b++) {
        hsc[0] += part[2 * b];
        hsc[1] += part[2 * b + 1];
    }
    CUDA_CHECK(cudaMemcpyAsync(c->scalars.p, hsc, sizeof(hsc), cudaMemcpyHostToDevice, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));

    // ---- the state an ordinary plan + accumulate leaves behind
    plan.max_abs = truth.max_abs;
    plan.max_abs_w = truth.max_abs_w;
    plan.sum_bound = truth.sum_bound;
    plan.err_weight = truth.err_weight;
    plan.err_weight2 = truth.err_weight2;
    plan.scale = truth.scale;
    plan.total_missing = truth.total_missing;
    plan.max_missing = truth.max_missing;
    plan.n_snp = m;
    plan.digits = nU;
    plan.digits_w = nW;
    plan.digits_d = nD;
    c->plan = plan;
    c->chunk_bound.assign(hchunk.begin(), hchunk.end());
    c->stat_valid = true;
    c->coltab_version = c->geno_version;
    c->coltab_est = est;
    c->coltab_bayesian = bayesian;
    c->plan_cache.version = c->geno_version;
    c->plan_cache.est = est;
    c->plan_cache.bayesian = bayesian;
    c->plan_cache.stats = truth;
    snprel_ctx::PrepCache &pc = c->prep_cache;
    pc.version = c->geno_version;
    pc.est = est;
    pc.bayesian = bayesian;
    pc.f = f; pc.fw = fw; pc.fd = fd; pc.fv = fv;
    pc.nU = nU; pc.nW = nW; pc.nD = nD; pc.nD2 = 0;
    pc.rounding = plan.rounding;
    pc.origin = c->snp_origin;
    pc.reduced = false;
    float ms = 0, total = 0;
    CUDA_CHECK(cudaEventRecord(c->evs1, c->stream));
    CUDA_CHECK(cudaEventSynchronize(c->evs1));
    CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    CUDA_CHECK(cudaEventElapsedTime(&total, c->evs0, c->evs1));
    c->hot_ms = ms;            // first tensor pass queued -> last one done (includes waiting for copies)
    c->step_ms = total;
    c->hot_units = 0.5 * (double)c->n_samp * (double)c->n_samp * (double)m;
    c->accum_win_r0 = 0;
    c->accum_win_rows = row_window(c).rows;
    c->accum_bayesian = bayesian ? 1 : 0;
    c->accum_est = est;
    c->accum_reduced = false;
    c->streamed_steps++;
    c->reduce_list.clear();
    c->reduce_list.push_back({c->acc.p, (int64_t)nplanes * npad * npad, 0, npad, npad, 0});
    c->reduce_list.push_back({c->samp_sum.p, (int64_t)c->samp_sum.n, 0});
    c->reduce_list.push_back({c->scalars.p, (int64_t)c->scalars.n, 2});
    c->reduce_list.push_back({c->iscalars.p, (int64_t)c->iscalars.n, 0});
    c->reduce_list.push_back({c->scr_cnt.p, (int64_t)c->scr_cnt.n, 1});
    return true;
}

// ---------------------------------------------------------------------------
// epilogues
// ---------------------------------------------------------------------------
__device__ __forceinline__ void store_sym2(double *out, int packed, int64_t n, int64_t i, int64_t j,
                                           double v) {
    if (packed) {
        out[j + i * (2 * n - i - 1) / 2] = v;
    } else {
        out[i * n + j] = v;
        out[j * n + i] = v;
    }
}

// ---- fused epilogue: int64 planes -> float64 result in its final layout, one pass -------------
// numerator C_ij = acc0[i][j] 2^-f - V_i (the per-sample vector in its own fixed point), then the
// estimator's normalisation.  mode 0: Eigenstrat (scale = (n-1)/trace); 1: GCTA; 3: EIGMIX (mul = 1 ibd / 2 GRM)
struct FinalArgs {
    const long long *acc;     // [planes][win.rows][npad]
    const long long *vec;     // [NVEC][npad]
    const int *nmiss;         // [npad] missing genotypes per sample
    long long n_snp_total;
    double inv_scale;                 // 2^-f
    double inv_v_hi, inv_v_lo;        // per-sample vector: 2^-(fv - 20) for VEC_W, 2^-fv for VEC_WLO
    long long sb_hi, sb_lo;           // its constant sum_l b_l in the same two units
    int mode, has_den, diagadj;
    double scale, sum_den, inv_fix, mul;
    long long nlocus;
    int64_t n, npad;
    RowWin win;
};

__device__ __forceinline__ double grm_numerator(const FinalArgs &a, int64_t i, int64_t j) {
    // a sample without a single valid genotype contributes exactly 0, as in the reference
    // (its centred genotypes are all 0, src/genPCA.h:103)
    if (a.nmiss[i] == a.n_snp_total || a.nmiss[j] == a.n_snp_total) return 0.0;
    // C_ij = sum T B + sum R m  -  V_i,   V_i = (VEC_W - sb_hi) 2^-(fv-20) + (VEC_WLO - sb_lo) 2^-fv
    const long long hi = a.vec[VEC_W * a.npad + i] - a.sb_hi, lo = a.vec[VEC_WLO * a.npad + i] - a.sb_lo;
    return (double)a.acc[(i - a.win.r0) * a.npad + j] * a.inv_scale - ((double)hi * a.inv_v_hi + (double)lo * a.inv_v_lo);
}

__device__ __forceinline__ double grm_entry(const FinalArgs &a, int64_t i, int64_t j) {   // i <= j
    double c = grm_numerator(a, i, j);
    if (a.mode == 0) return c * a.scale;
    long long dq = 0;
    if (a.has_den)
        dq = a.vec[VEC_D * a.npad + i] + a.vec[VEC_D * a.npad + j] -
             a.acc[a.win.rows * a.npad + (i - a.win.r0) * a.npad + j];
    if (a.mode == 1) return c / (double)(2 * (a.nlocus - dq));      // src/genPCA.cpp:1233-1236
    if (a.diagadj && i == j) c -= (double)a.vec[VEC_HET * a.npad + i];   // src/genEIGMIX.cpp:147-151
    return c / (a.sum_den - (double)dq * a.inv_fix) * a.mul;            // :153-155, :645-652
}

// packed slice of the window's rows (row-packed upper triangle, CdMatTri order)
__global__ void grm_final_packed_kernel(const FinalArgs a, double *__restrict__ out) {
    const int64_t i = a.win.r0 + blockIdx.x, j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= a.n || j < i) return;
    out[tri_idx(a.n, i, j) - a.win.pbase] = grm_entry(a, i, j);
}

// full symmetric n x n: 32 x 32 tiles of the upper triangle, the mirror image goes through shared
// memory so that both the direct and the transposed store are coalesced
__global__ void __launch_bounds__(256) grm_final_full_kernel(const FinalArgs a, double *__restrict__ out) {
    __shared__ double t[32][33];
    const int64_t bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int64_t i = bi * 32 + ty + 8 * r, j = bj * 32 + tx;
        double v = 0.0;
        if (i < a.n && j < a.n && j >= i) {
            v = grm_entry(a, i, j);
            out[i * a.n + j] = v;
        }
        t[ty + 8 * r][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int64_t j = bj * 32 + ty + 8 * r, i = bi * 32 + tx;   // out[j][i] = entry (i, j)
        if (i < a.n && j < a.n && j > i) out[j * a.n + i] = t[tx][ty + 8 * r];
    }
}

// numerators of the diagonal (trace of the covariance, CdMatTri::Trace src/genPCA.cpp:1387)
__global__ void diag_numerator_kernel(const FinalArgs a, double *__restrict__ d) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.n) d[i] = grm_numerator(a, i, i);
}

__global__ void diag_kernel(const double *__restrict__ m, double *__restrict__ d, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = m[i * n + i];
}

// gnrGRM "Corr": p_ij / sqrt(p_ii p_jj), diagonal forced to 1 (src/genPCA.cpp:1670-1685)
__global__ void corr_kernel(double *__restrict__ g, const double *__restrict__ diag, int64_t n) {
    int64_t i = blockIdx.x, j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= n || j < i) return;
    double v = (i == j) ? 1.0 : g[i * n + j] / (sqrt(diag[i]) * sqrt(diag[j]));
    g[i * n + j] = v;
    g[j * n + i] = v;
}

__global__ void symmetrize_neg_kernel(const double *__restrict__ src, double *__restrict__ dst,
                                      double sign, int64_t n) {
    int64_t i = blockIdx.x, j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= n || j < i) return;
    double v = sign * src[i * n + j];
    dst[i * n + j] = v;
    dst[j * n + i] = v;
}

static dim3 tri_grid(int64_t n) { return dim3((unsigned)n, (unsigned)((n + 127) / 128)); }

static size_t out_count(int64_t n, int packed) {
    return packed ? (size_t)n * (n + 1) / 2 : (size_t)n * n;
}

template <class T>
static void d2h(snprel_ctx *c, T *host, const T *dev, size_t count) {
    CUDA_CHECK(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

static void need_grm_accum(snprel_ctx *c, int est, int bayesian) {
    if (est == SNPREL_GRM_CORR) est = SNPREL_GRM_GCTA;
    const RowWin w = row_window(c);
    if (c->accum_est == est && c->accum_reduced && c->accum_win_r0 == w.r0 && c->accum_win_rows == w.rows) {
        // accumulators already summed over the ranks: they must have been built for this request
        if (c->accum_bayesian != (bayesian ? 1 : 0))
            fail("the reduced accumulators were built with bayesian = %s; accumulate again with the flag you finish with",
                 c->accum_bayesian ? "TRUE" : "FALSE");
        return;
    }
    // copies still in flight: consume them chunk by chunk (or wait and take the ordinary path)
    if (!c->pending.empty() && grm_accumulate_streamed(c, est, bayesian)) return;
    geno_wait(c);
    snprel_plan plan{};
    plan.frac_bits = -1;
    plan.frac_bits_w = -1;
    plan.frac_bits_d = -1;
    plan.bayesian = bayesian;
    grm_plan_local(c, est, &plan);
    grm_accumulate(c, est, &plan);
}

struct Globals {
    double sum_den, sum_den2;
    long long nlocus, sb_lo, sb_hi;
};
static Globals read_globals(snprel_ctx *c) {
    double hs[4];
    long long hi[4];
    d2h(c, hs, c->scalars.p, 4);
    d2h(c, hi, c->iscalars.p, 4);
    return Globals{hs[0], hs[1], hi[0], hi[2], hi[3]};
}

static dim3 win_grid(snprel_ctx *c) {
    RowWin w = row_window(c);
    return dim3((unsigned)(w.r1 - w.r0), (unsigned)((c->n_samp + 127) / 128));
}
static size_t win_out_count(snprel_ctx *c, int packed) {
    if (!full_window(c)) return window_packed_count(c);
    return out_count(c->n_samp, packed);
}

// sum of n device doubles in index order on the host (the reference's CdMatTri::Trace order);
// the staging buffers are persistent
static double host_sum(snprel_ctx *c, const double *dev, int64_t n) {
    c->host_diag.resize((size_t)n);
    d2h(c, c->host_diag.data(), dev, (size_t)n);
    double t = 0;
    for (int64_t i = 0; i < n; i++) t += c->host_diag[i];
    return t;
}

static double trace_of(snprel_ctx *c, const double *m, int64_t n) {
    c->scr_diag.alloc((size_t)n);
    diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(m, c->scr_diag.p, n);
    KERNEL_CHECK(c);
    return host_sum(c, c->scr_diag.p, n);
}

// device-side result of the covariance family: full symmetric / packed matrix in c->scr_out
// (persistent: a cudaMalloc + cudaFree of the 800 MB result per call cost more than the whole
// epilogue -- and up to 0.5 s on a cold driver)
static void grm_device(snprel_ctx *c, int method, int packed, int diagadj, double mul, double *trace_xtx) {
    const int64_t n = c->n_samp;
    if (!full_window(c) && !packed) fail("a row window returns the packed upper triangle only (useMatrix)");
    output_wait(c);
    DevBuf<double> &o = c->scr_out;
    o.alloc(win_out_count(c, packed));
    FinalArgs a{};
    a.acc = c->acc.p;
    a.vec = c->samp_sum.p;
    a.nmiss = c->scr_cnt.p + c->n_samp_pad;
    a.n_snp_total = (long long)c->plan.n_snp;
    const Globals g = read_globals(c);
    a.inv_scale = std::ldexp(1.0, -c->plan.frac_bits);
    a.inv_v_hi = std::ldexp(1.0, -(c->plan.frac_bits_v - LIMB));
    a.inv_v_lo = std::ldexp(1.0, -c->plan.frac_bits_v);
    a.sb_hi = g.sb_hi;
    a.sb_lo = g.sb_lo;
    a.has_den = c->acc_planes > 1;
    a.diagadj = diagadj;
    a.mul = mul;
    a.n = n;
    a.npad = c->n_samp_pad;
    a.win = row_window(c);
    if (method == SNPREL_GRM_EIGENSTRAT) {
        // whole matrix: trace of the computed numerator (CdMatTri::Trace, src/genPCA.cpp:1387);
        // row window: the diagonal lives in other windows, so use the same trace evaluated from
        // the per-SNP genotype counts by the plan kernel (equal up to float64 rounding)
        double tr;
        if (full_window(c)) {
            c->scr_diag.alloc((size_t)n);
            diag_numerator_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(a, c->scr_diag.p);
            KERNEL_CHECK(c);
            tr = host_sum(c, c->scr_diag.p, n);
        } else {
            tr = c->plan.scale * (double)std::max<int64_t>(n - 1, 1);
        }
        if (trace_xtx) *trace_xtx = tr;
        a.mode = 0;
        a.scale = (double)(n - 1) / tr;
    } else if (method == SNPREL_GRM_GCTA || method == SNPREL_GRM_CORR) {
        a.mode = 1;
        a.nlocus = g.nlocus;
    } else {
        a.mode = 3;
        a.sum_den = g.sum_den;
        a.inv_fix = std::ldexp(1.0, -c->plan.frac_bits_d);
    }
    if (packed) {
        grm_final_packed_kernel<<<win_grid(c), 128, 0, c->stream>>>(a, o.p);
    } else {
        const unsigned nb = (unsigned)((n + 31) / 32);
        grm_final_full_kernel<<<dim3(nb, nb), 256, 0, c->stream>>>(a, o.p);
    }
    KERNEL_CHECK(c);
}

// device epilogue of the accumulators at hand, result left in c->scr_out (bench timing:
// SURVEY section 8d counts the step up to "final N x N complete")
void grm_finish_device(snprel_ctx *c, int est) {
    if (est == SNPREL_GRM_CORR) est = SNPREL_GRM_GCTA;
    if (c->accum_est != est) fail("snprel_time_finish: no accumulators for estimator %d", est);
    grm_device(c, est, full_window(c) ? 0 : 1, 0, est == SNPREL_GRM_EIGMIX ? 2.0 : 1.0, nullptr);
}

void grm_finish(snprel_ctx *c, int method, double *out, int packed) {
    if (!out) fail("snprel_grm: NULL output");
    int64_t n = c->n_samp;
    if (!full_window(c) && method == SNPREL_GRM_CORR) fail("method \"Corr\" needs the whole matrix, not a row window");
    need_grm_accum(c, method, 0);
    DevBuf<double> &o = c->scr_out;
    if (method == SNPREL_GRM_CORR) {
        grm_device(c, SNPREL_GRM_GCTA, 0, 0, 1, nullptr);
        c->scr_diag.alloc((size_t)n);
        diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(o.p, c->scr_diag.p, n);
        KERNEL_CHECK(c);
        corr_kernel<<<tri_grid(n), 128, 0, c->stream>>>(o.p, c->scr_diag.p, n);
        KERNEL_CHECK(c);
        d2h(c, out, o.p, (size_t)n * n);
        return;
    }
    grm_device(c, method, packed, 0, method == SNPREL_GRM_EIGMIX ? 2.0 : 1.0, nullptr);
    deliver(c, out, o.p, win_out_count(c, packed));
}

void pca_finish(snprel_ctx *c, int eigen_cnt, int bayesian, double *genmat, double *trace_xtx,
                double *trace_val, double *eigval, double *eigvec) {
    if (eigen_cnt < 0) fail("Invalid 'eigen.cnt'.");   // src/genPCA.cpp:1423-1424
    if (!full_window(c)) fail("snprel_pca needs the whole matrix, not a row window");
    int64_t n = c->n_samp;
    need_grm_accum(c, SNPREL_GRM_EIGENSTRAT, bayesian);
    DevBuf<double> &o = c->scr_out;
    double tx = 0;
    grm_device(c, SNPREL_GRM_EIGENSTRAT, 0, 0, 1, &tx);
    if (trace_xtx) *trace_xtx = tx;
    // the big copy first: the 80 KB diagonal for TraceVal rides behind it on the same stream
    if (genmat) CUDA_CHECK(cudaMemcpyAsync(genmat, o.p, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (trace_val) *trace_val = trace_of(c, o.p, n);
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (eigval || eigvec) eigen_topk(c, o.p, n, eigen_cnt, eigval, eigvec);
}

void eigmix_finish(snprel_ctx *c, int eigen_cnt, int diagadj, double *ibd, double *afreq, double *eigval,
                   double *eigvec) {
    int64_t n = c->n_samp;
    if (!full_window(c)) fail("snprel_eigmix needs the whole matrix, not a row window (use snprel_grm EIGMIX)");
    need_grm_accum(c, SNPREL_GRM_EIGMIX, 0);
    if (eigen_cnt < 0 || eigen_cnt > n) eigen_cnt = (int)n;   // src/genEIGMIX.cpp:675-676
    DevBuf<double> &o = c->scr_out;
    grm_device(c, SNPREL_GRM_EIGMIX, 0, diagadj, 1.0, nullptr);
    if (ibd) d2h(c, ibd, o.p, (size_t)n * n);
    if (afreq) {   // af = 0.5 * avg_geno (src/genEIGMIX.cpp:116-118)
        std::vector<SnpStat> h((size_t)c->n_snp);
        d2h(c, h.data(), c->stat.p, (size_t)c->n_snp);
        for (int64_t l = 0; l < c->n_snp; l++)
            afreq[l] = 0.5 * (h[l].num > 0 ? (double)h[l].sum / (double)h[l].num : 0.0);
    }
    if ((eigval || eigvec) && eigen_cnt > 0) eigen_topk(c, o.p, n, eigen_cnt, eigval, eigvec);
}

// ---- KING-homo: IBS0/SumSq from the packed-bit kernel, the two float sums from
// the table Gram (jointly-valid sums = total - r_i - r_j + both-missing Gram) ----
__global__ void king_homo_kernel(const uint32_t *__restrict__ cnt, const long long *__restrict__ acc,
                                 const long long *__restrict__ vec, double *__restrict__ k0o,
                                 double *__restrict__ k1o, int packed, double s1, double s2,
                                 double inv_fix, int has_den, int64_t n, int64_t npad, RowWin win) {
    int64_t i = win.r0 + blockIdx.x, j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= n || j < i) return;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double *k0p = packed ? k0o - win.pbase : k0o, *k1p = packed ? k1o - win.pbase : k1o;
    if (i == j) {   // src/genKING.cpp:524
        store_sym2(k0p, packed, n, i, j, 0.0);
        store_sym2(k1p, packed, n, i, j, 0.0);
        return;
    }
    int64_t plane = win.rows * npad, k = (i - win.r0) * npad + j;
    double ibs0 = (double)cnt[k], sumsq = (double)(cnt[2 * plane + k] + 4u * cnt[k]);
    double a1 = s1, a2 = s2;
    if (has_den) {
        long long q1 = vec[VEC_D * npad + i] + vec[VEC_D * npad + j] - acc[k];
        long long q2 = vec[VEC_D2 * npad + i] + vec[VEC_D2 * npad + j] - acc[plane + k];
        a1 -= (double)q1 * inv_fix;
        a2 -= (double)q2 * inv_fix;
    }
    double theta = 0.5 - sumsq / (8 * a1);          // src/genKING.cpp:527-529
    double k0 = ibs0 / (2 * a2);
    double k1 = 2 - 2 * k0 - 4 * theta;
    store_sym2(k0p, packed, n, i, j, isfinite(k0) ? k0 : nan);
    store_sym2(k1p, packed, n, i, j, isfinite(k1) ? k1 : nan);
}

void king_homo_finish(snprel_ctx *c, double *k0, double *k1, int packed) {
    if (!k0 || !k1) fail("snprel_king_homo: NULL output");
    int64_t n = c->n_samp;
    // float sums first (they own acc/samp_sum), then the integer counters (they own cnt)
    snprel_plan plan{};
    plan.frac_bits = -1;
    plan.frac_bits_w = -1;
    plan.frac_bits_d = -1;
    grm_plan_local(c, SNPREL_EST_KING_HOMO, &plan);
    grm_accumulate(c, SNPREL_EST_KING_HOMO, &plan);
    Globals g = read_globals(c);
    const int has_den = plan.total_missing > 0;
    const int fb = c->plan.frac_bits_d;
    bitcount_accumulate(c, SNPREL_EST_KING_ROBUST);
    if (!full_window(c) && !packed) fail("snprel_king_homo: a row window returns the packed upper triangle only");
    DevBuf<double> o;
    size_t oc = win_out_count(c, packed);
    o.alloc(2 * oc);
    king_homo_kernel<<<win_grid(c), 128, 0, c->stream>>>(c->cnt.p, c->acc.p, c->samp_sum.p, o.p, o.p + oc,
                                                         packed, g.sum_den, g.sum_den2,
                                                         std::ldexp(1.0, -fb), has_den, n, c->n_samp_pad,
                                                         row_window(c));
    KERNEL_CHECK(c);
    d2h(c, k0, o.p, oc);
    d2h(c, k1, o.p + oc, oc);
    c->accum_est = -1;
}

// ---- test hook: exact table Gram with caller tables -------------------------
void table_gram_debug(snprel_ctx *c, const int8_t *tabA, const int8_t *tabB, int64_t *out) {
    if (!tabA || !tabB || !out) fail("snprel_table_gram: NULL argument");
    if (c->n_samp <= 0) fail("snprel_table_gram: no genotype workspace");
    const int64_t cap = c->snp_cap, npad = c->n_samp_pad, n = c->n_samp;
    DevBuf<uint32_t> tab;
    tab.alloc((size_t)cap);
    tab.zero(c->stream);
    if (c->n_snp > 0)
        CUDA_CHECK(cudaMemcpyAsync(tab.p, tabA, (size_t)c->n_snp * 4, cudaMemcpyHostToDevice, c->stream));
    uint32_t tb;
    memcpy(&tb, tabB, 4);
    DevBuf<long long> acc;
    acc.alloc((size_t)npad * npad);
    acc.zero(c->stream);
    int bmax = 0;
    for (int g = 0; g < 4; g++) bmax = std::max(bmax, std::abs((int)tabB[g]));
    GramPass p{tab.p, gram_const_table(c, tb), 0, 0, std::max(bmax, 1), nullptr};
    c->hot_launches = 0;
    gram_tc_run(c, &p, 1, acc.p, false);
    CUDA_CHECK(cudaMemcpy2DAsync(out, (size_t)n * 8, acc.p, (size_t)npad * 8, (size_t)n * 8, (size_t)n,
                                 cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

}  // namespace snprel
