// Covariance-type estimators on top of the tcgen05 table Gram (gram_tc.cu):
//   Eigenstrat / PCA   CExactPCA::Run + gnrPCA       src/genPCA.cpp:395-464,1355-1452
//   GCTA, Corr         CGCTA_AlgArith::Run + gnrGRM  src/genPCA.cpp:1148-1237,1614-1717
//   EIGMIX             CEigMix_AlgArith::Run         src/genEIGMIX.cpp:60-156,645-735
//   KING-homo sums     CKINGHomo                     src/genKING.cpp:69-266,493-570
//
// Algebra.  With x = genotype (0 when missing), m = missing indicator, mu_l the mean
// genotype of SNP l over non-missing samples and w_l the per-SNP weight
// (1/(p(1-p)) for Eigenstrat/GCTA, 1 for EIGMIX), the reference accumulates
//     C_ij = sum_l w_l (x_il - mu_l (1-m_il)) (x_jl - mu_l (1-m_jl)).
// Define the per-SNP 4-entry tables over the genotype code g (0,1,2,3=missing)
//     U_l[g] = w_l (g - mu_l)  (U_l[3] = 0),      W_l[g] = mu_l U_l[g].
// Then  C_ij = sum_l U_l[g_il] x_jl  -  sum_l W_l[g_il]  +  sum_l W_l[g_il] m_jl.
// The first and third sums are table Grams against the small-integer channels x and
// m; the middle one is a per-sample vector.  Tables are quantised to fixed point
// (2^-frac_bits) and split into balanced base-256 digits, one int8 tensor-core pass
// per digit; the quantisation error of an entry is bounded by
// 2^-(frac_bits+1) * (2 #SNP + #missing) and everything after quantisation is exact
// integer arithmetic.  Denominators (GCTA: #polymorphic SNPs with i or j missing;
// EIGMIX: sum 4p(1-p) over SNPs with i or j missing) use
//     D_ij = r_i + r_j - sum_l d_l m_il m_jl,   r_i = sum_l d_l m_il,
// i.e. one more table Gram (channel m on both sides) and a per-sample vector.
#include <cmath>

#include "common.cuh"

namespace snprel {

constexpr uint32_t TABB_X = 0x00020100u;   // x channel: code -> {0,1,2,0}
constexpr uint32_t TABB_M = 0x01000000u;   // m channel: code -> {0,0,0,1}
constexpr int MAX_DIGITS = 8;

enum { VEC_W = 0, VEC_D = 1, VEC_HET = 2, VEC_D2 = 3, VEC_WLO = 4, NVEC = 5 };
constexpr int W_EXTRA_BITS = 20;   // the per-sample W vector carries 20 more fractional bits than the passes
// scalars[]: 0 = sum of d_l (EIGMIX SumDenominator / KING-homo sum p(1-p)), 1 = sum d2_l
// iscalars[]: 0 = nLocus (GCTA), 1 = total missing genotypes (valid samples only)

struct SnpTables {
    long long qU[4];   // fixed-point U
    long long qW[4];   // fixed-point W
    long long qWx[4];  // fixed-point W with W_EXTRA_BITS more fractional bits (per-sample vector only)
    double diag;       // sum over samples of U[g] (g - mu): this SNP's contribution to trace(C)
    long long qD;      // fixed-point d (GCTA: 0/1 unscaled; EIGMIX / KING-homo: 2^frac_bits scaled)
    long long qD2;     // KING-homo: (p(1-p))^2
    double d, d2;      // float64 d, d2 (for the global scalars)
    double maxU, maxW;
};

// est: SNPREL_GRM_EIGENSTRAT / GCTA / CORR / EIGMIX, or SNPREL_EST_KING_HOMO
// uniform [0, 1) draw of table entry (SNP l, genotype g): randomised rounding (snprel_set_rounding)
__device__ __forceinline__ double dither_u01(uint64_t seed, long long l, int g) {
    uint64_t x = seed ^ ((uint64_t)l * 4u + (uint64_t)g) * 0xD1342543DE82EF95ull;
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

// dither != 0: the U table is rounded at random, floor(v 2^f + u) with u ~ U[0, 1) drawn per
// (global SNP index snp_index, genotype) -- unbiased and independent across SNPs, which is what the
// Hoeffding bound of choose_format relies on; 0: round to nearest (the default, worst-case bound)
__device__ __forceinline__ void snp_tables(const SnpStat st, int est, int bayesian, int frac_bits,
                                           int frac_bits_w, int frac_bits_d, SnpTables &t,
                                           long long snp_index = 0, uint64_t dither = 0) {
    const double sc = exp2((double)frac_bits);
    const double scw = exp2((double)frac_bits_w);
    const double scd = exp2((double)frac_bits_d);
    const double scx = exp2((double)(frac_bits + W_EXTRA_BITS));
    double mu = st.num > 0 ? (double)st.sum / (double)st.num : 0.0;   // DivideGeno, src/genPCA.cpp:98-142
    double w = 0, d = 0, d2 = 0;
    long long qD = 0, qD2 = 0;
    if (est == SNPREL_GRM_EIGMIX) {
        w = 1.0;
        double af = 0.5 * mu;
        d = 4 * af * (1 - af);                    // src/genEIGMIX.cpp:116-119
        qD = llrint(d * scd);
    } else if (est == SNPREL_EST_KING_HOMO) {
        double p = st.num > 0 ? 0.5 * (double)st.sum / (double)st.num : 0.0;   // src/genKING.cpp:239-241
        d = p * (1 - p);
        d2 = d * d;
        qD = llrint(d * scd);
        qD2 = llrint(d2 * scd);
    } else {
        if (bayesian) {                           // src/genPCA.cpp:445-452
            double s = ((double)st.sum + 1.0) / (double)(2 * st.num + 2);
            double r = 1.0 / sqrt(s * (1 - s));
            w = r * r;
        } else {                                  // rsqrt_prod, src/genPCA.cpp:145-181
            double s = mu * 0.5;
            if (0 < s && s < 1) {
                double r = 1.0 / sqrt(s * (1 - s));
                w = r * r;
            }
        }
        bool poly = (0 < st.sum) && (st.sum < 2 * st.num);   // src/genPCA.cpp:1206
        d = poly ? 1.0 : 0.0;
        qD = poly ? 1 : 0;
    }
    t.maxU = 0;
    t.maxW = 0;
    const int n2 = (st.sum - st.n1) / 2, n0 = st.num - st.n1 - n2;
    const double cnt[3] = {(double)n0, (double)st.n1, (double)n2};
    t.diag = 0;
#pragma unroll
    for (int g = 0; g < 3; g++) {
        double u = (est == SNPREL_EST_KING_HOMO) ? 0.0 : w * ((double)g - mu);
        double ww = mu * u;
        t.qU[g] = dither ? (long long)floor(u * sc + dither_u01(dither, snp_index, g)) : llrint(u * sc);
        t.qW[g] = llrint(ww * scw);
        t.qWx[g] = llrint(ww * scx);
        t.maxU = fmax(t.maxU, fabs(u));
        t.maxW = fmax(t.maxW, fabs(ww));
        t.diag += cnt[g] * u * ((double)g - mu);
    }
    t.qU[3] = 0;
    t.qW[3] = 0;
    t.qWx[3] = 0;
    t.qD = qD;
    t.qD2 = qD2;
    t.d = d;
    t.d2 = d2;
}

__device__ __forceinline__ uint32_t digit_of(long long &q) {
    long long dgt = ((q + 128) & 255) - 128;   // balanced digit in [-128, 127]
    q = (q - dgt) >> 8;
    return (uint32_t)(dgt & 255);
}

// ---- plan statistics ---------------------------------------------------------
// out[0] max |table value|, out[1] int64-range bound, out[2] total missing,
// out[3] local share of the normaliser (trace(C) / nLocus / sum d / sum d2)
__global__ void plan_kernel(const SnpStat *__restrict__ st, int64_t n_snp, int64_t n_samp, int est,
                            int bayesian, double *__restrict__ out) {
    double mx = 0, mxw = 0, sb = 0, tm = 0, sc = 0;
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < n_snp;
         l += (int64_t)gridDim.x * blockDim.x) {
        SnpTables t;
        snp_tables(st[l], est, bayesian, 0, 0, 0, t);
        const bool realD = (est == SNPREL_GRM_EIGMIX || est == SNPREL_EST_KING_HOMO);
        double dmax = realD ? fmax(t.d, t.d2) : 0.0;
        mx = fmax(mx, t.maxU);
        mxw = fmax(mxw, t.maxW);
        sb += 2 * t.maxU + 2 * t.maxW + 2 * dmax;
        tm += (double)(n_samp - st[l].num);
        if (est == SNPREL_GRM_EIGENSTRAT) sc += t.diag;
        else if (est == SNPREL_GRM_EIGMIX) sc += t.d;
        else if (est == SNPREL_EST_KING_HOMO) sc += t.d2;
        else sc += t.d;   // GCTA: number of polymorphic SNPs
    }
    __shared__ double s0[256], s1[256], s2[256], s3[256], s4[256];
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
    }
}

// per-sample genotype sum and missing count (error weight of the fixed-point format)
constexpr int SC_SNPS = 2048;
__global__ void __launch_bounds__(128)
sample_count_kernel(const uint8_t *__restrict__ geno, int64_t n_snp, int64_t row_bytes, int64_t npad,
                    int *__restrict__ cnt /*[2][npad]: sum x, #missing*/) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= row_bytes) return;
    const int64_t l0 = (int64_t)blockIdx.y * SC_SNPS;
    const int nl = (int)min((int64_t)SC_SNPS, n_snp - l0);
    const uint8_t *p = geno + l0 * row_bytes + b;
    // byte-sliced counters: each 2-bit field of a byte counted in its own 8-bit lane
    uint32_t c1 = 0, c2 = 0, c3 = 0;   // four 8-bit lanes each: #code1, #code2, #code3
    int s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0}, s3[4] = {0, 0, 0, 0};
    for (int s = 0; s < nl; s++) {
        uint32_t v = p[(int64_t)s * row_bytes];
        // spread the four 2-bit codes of the byte to the low bits of four byte lanes
        uint32_t w = (v | (v << 6) | (v << 12) | (v << 18)) & 0x03030303u;
        uint32_t lo = w & 0x01010101u, hi = (w >> 1) & 0x01010101u;
        c1 += lo & ~hi;
        c2 += hi & ~lo;
        c3 += lo & hi;
        if ((s & 127) == 127) {   // flush before an 8-bit lane can overflow
#pragma unroll
            for (int k = 0; k < 4; k++) {
                s1[k] += (c1 >> (8 * k)) & 255;
                s2[k] += (c2 >> (8 * k)) & 255;
                s3[k] += (c3 >> (8 * k)) & 255;
            }
            c1 = c2 = c3 = 0;
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        s1[k] += (c1 >> (8 * k)) & 255;
        s2[k] += (c2 >> (8 * k)) & 255;
        s3[k] += (c3 >> (8 * k)) & 255;
        int64_t i = b * 4 + k;
        int sx = s1[k] + 2 * s2[k];
        if (sx) atomicAdd(cnt + i, sx);
        if (s3[k]) atomicAdd(cnt + npad + i, s3[k]);
    }
}

// ---- digit tables: tab[pass][snp] ------------------------------------------
// pass order: U digits (nU), W digits (nW), D digits (nD), D2 digits (nD2)
__global__ void tables_kernel(const SnpStat *__restrict__ st, int64_t n_snp, int64_t cap, int est,
                              int bayesian, int frac_bits, int frac_bits_w, int frac_bits_d, int nU, int nW, int nD, int nD2,
                              uint32_t *__restrict__ tab, double *__restrict__ scalars /*[gridDim.x][2] partials*/,
                              long long *__restrict__ iscalars, int *__restrict__ overflow, uint64_t dither) {
    int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double d = 0, d2 = 0;
    long long poly = 0;
    if (l < n_snp) {
        SnpTables t;
        snp_tables(st[l], est, bayesian, frac_bits, frac_bits_w, frac_bits_d, t, (long long)l, dither);
        d = t.d;
        d2 = t.d2;
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
    __shared__ long long s2[256];
    s0[threadIdx.x] = d;
    s1[threadIdx.x] = d2;
    s2[threadIdx.x] = poly;
    __syncthreads();
    for (int o = blockDim.x / 2; o; o >>= 1) {
        if (threadIdx.x < o) {
            s0[threadIdx.x] += s0[threadIdx.x + o];
            s1[threadIdx.x] += s1[threadIdx.x + o];
            s2[threadIdx.x] += s2[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        // float64 block partials are summed in block order by the host (run-to-run deterministic)
        scalars[2 * blockIdx.x + 0] = s0[0];
        scalars[2 * blockIdx.x + 1] = s1[0];
        atomicAdd(reinterpret_cast<unsigned long long *>(iscalars), (unsigned long long)s2[0]);
    }
}

// ---- per-sample sums: vec[v][i] = sum_l T_v,l[g_il]  (exact int64) ---------
constexpr int SS_SNPS = 512;   // SNPs per block
__global__ void __launch_bounds__(128)
sample_sum_kernel(const uint8_t *__restrict__ geno, const SnpStat *__restrict__ st, int64_t n_snp,
                  int64_t row_bytes, int64_t npad, int est, int bayesian, int frac_bits, int frac_bits_w, int frac_bits_d,
                  long long *__restrict__ vec) {
    __shared__ long long tW[SS_SNPS][4];    // hi part of the extended-precision W (units 2^-frac_bits)
    __shared__ int tWlo[SS_SNPS][4];        // lo part (W_EXTRA_BITS bits, non-negative)
    __shared__ long long tD[SS_SNPS], tD2[SS_SNPS];
    const int64_t l0 = (int64_t)blockIdx.y * SS_SNPS;
    const int nl = (int)min((int64_t)SS_SNPS, n_snp - l0);
    for (int s = threadIdx.x; s < nl; s += blockDim.x) {
        SnpTables t;
        snp_tables(st[l0 + s], est, bayesian, frac_bits, frac_bits_w, frac_bits_d, t);
        for (int g = 0; g < 4; g++) {
            tW[s][g] = t.qWx[g] >> W_EXTRA_BITS;
            tWlo[s][g] = (int)(t.qWx[g] & ((1ll << W_EXTRA_BITS) - 1));
        }
        tD[s] = t.qD;
        tD2[s] = t.qD2;
    }
    __syncthreads();
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // byte column (4 samples)
    if (b >= row_bytes) return;
    long long aw[4] = {0, 0, 0, 0}, awl[4] = {0, 0, 0, 0}, ad[4] = {0, 0, 0, 0}, ad2[4] = {0, 0, 0, 0};
    int ah[4] = {0, 0, 0, 0};
    const uint8_t *p = geno + l0 * row_bytes + b;
    // eight independent byte loads in flight per thread: the serial load -> table lookup -> add
    // chain was latency bound (12.6 ms for 2.5 GB in profiles/r01_launches_pca.csv)
    constexpr int UNR = 8;
    for (int s0 = 0; s0 < nl; s0 += UNR) {
        uint32_t vv[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) vv[u] = (s0 + u < nl) ? p[(int64_t)(s0 + u) * row_bytes] : 0xFFu;
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            const int s = min(s0 + u, nl - 1);
            const bool live = s0 + u < nl;
            const uint32_t v = vv[u];
            long long dd = live ? tD[s] : 0, dd2 = live ? tD2[s] : 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                uint32_t code = (v >> (2 * k)) & 3;   // a dead slot reads as missing: table entry 3 is 0
                aw[k] += tW[s][code];
                awl[k] += tWlo[s][code];
                if (code == 3) {
                    ad[k] += dd;
                    ad2[k] += dd2;
                }
                ah[k] += (code == 1);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int64_t i = b * 4 + k;
        unsigned long long *o = reinterpret_cast<unsigned long long *>(vec);
        // W vector: hi part in units of 2^-frac_bits, lo part carries W_EXTRA_BITS more bits
        if (aw[k]) atomicAdd(o + VEC_W * npad + i, (unsigned long long)aw[k]);
        if (awl[k]) atomicAdd(o + VEC_WLO * npad + i, (unsigned long long)awl[k]);
        if (ad[k]) atomicAdd(o + VEC_D * npad + i, (unsigned long long)ad[k]);
        if (ah[k]) atomicAdd(o + VEC_HET * npad + i, (unsigned long long)(long long)ah[k]);
        if (ad2[k]) atomicAdd(o + VEC_D2 * npad + i, (unsigned long long)ad2[k]);
    }
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

// scr_cnt[2][npad]: this rank's per-sample genotype sums and missing counts
static void sample_counts(snprel_ctx *c) {
    const int64_t npad = c->n_samp_pad;
    c->scr_cnt.alloc((size_t)2 * npad);
    c->scr_cnt.zero(c->stream);
    if (c->n_snp > 0) {
        dim3 grid((unsigned)((c->row_bytes + 127) / 128), (unsigned)((c->n_snp + SC_SNPS - 1) / SC_SNPS));
        sample_count_kernel<<<grid, 128, 0, c->stream>>>(c->geno2b.p, c->n_snp, c->row_bytes, npad, c->scr_cnt.p);
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
            return;
        }
    }
    ensure_stats(c);
    const int64_t npad = c->n_samp_pad;
    DevBuf<double> &out = c->scr_plan;
    out.alloc(5);
    out.zero(c->stream);
    if (c->n_snp > 0) {
        int blocks = (int)std::min<int64_t>((c->n_snp + 255) / 256, 1024);
        plan_kernel<<<blocks, 256, 0, c->stream>>>(c->stat.p, c->n_snp, c->n_samp, est,
                                                   plan->bayesian, out.p);
        KERNEL_CHECK(c);
    }
    sample_counts(c);
    double h[5];
    c->host_cnt.resize((size_t)2 * npad);
    CUDA_CHECK(cudaMemcpyAsync(h, out.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(c->host_cnt.data(), c->scr_cnt.p, (size_t)2 * npad * sizeof(int),
                               cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    long long ew = 0, mm = 0;
    for (int64_t i = 0; i < c->n_samp; i++) {
        long long sx = c->host_cnt[i], ms = c->host_cnt[npad + i];
        ew = std::max(ew, sx);
        mm = std::max(mm, ms);
    }
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
//        <= 2^-(f+1) * err_weight  (U table against the x channel)
//         + 2^-(fw+1) * max_missing (W table against the m channel),   wanted <= tol * scale
//   denominator plane (EIGMIX / KING-homo): error <= 2^-(fd+1) * 2 max_missing, wanted <= tol * scale
// Tensor passes run two at a time, so a table with n digits costs n/2 double launches plus
// (n odd) one single launch that runs at ~0.8 of a double.
static double launch_cost(int n) { return (n / 2) + (n % 2) * 0.8; }

// round_mode 1 (snprel_set_rounding): the U table is rounded at random, so for a fixed pair (i, j)
// the error sum_l e_l[g_il] x_jl is a sum of independent zero-mean terms of width x_jl 2^-f, and by
// Hoeffding it exceeds 2^-f sqrt(1/2 sum_l x_jl^2 ln(2/delta)) with probability < delta; delta is
// 1e-12 divided by the number of pairs (union bound) and sum x^2 <= 2 sum x = 2 err_weight.  This
// replaces the worst-case 2^-(f+1) err_weight and saves one tensor pass per table at C2
// (tools/fixed_point_model.py: actual error 1.8e-11, bound 6.5e-11 with U in 4 digits).
static double u_table_error(const snprel_plan &plan, int fa, int round_mode, double n_samp) {
    if (round_mode != 1) return std::ldexp(plan.err_weight, -(fa + 1));
    const double pairs = std::max(1.0, 0.5 * n_samp * (n_samp + 1.0));
    return std::ldexp(std::sqrt(0.5 * 2.0 * plan.err_weight * std::log(2.0 * pairs / 1e-12)), -fa);
}

static void choose_format(int est, snprel_plan &plan, int &nU, int &nW, int &nD, int round_mode = 0,
                          double n_samp = 0) {
    const double tol = (plan.tol > 0 ? plan.tol : 1e-10) * 0.9;   // 10 % left for float64 rounding in the epilogue
    const bool homo = est == SNPREL_EST_KING_HOMO;
    const bool any_missing = plan.total_missing > 0;
    const int head = 61 - (int)std::ceil(std::log2(std::max(plan.sum_bound, 1.0)));   // int64 plane headroom
    if (head < 16) fail("fixed-point accumulator cannot hold this data set (sum bound %.3g)", plan.sum_bound);
    double scale = plan.scale;
    if (est == SNPREL_GRM_GCTA) scale -= 4.0 * (double)plan.max_missing;   // 2 (nLocus - D_ij), D_ij <= 2 max_missing
    if (est == SNPREL_GRM_EIGMIX) scale -= 2.0 * (double)plan.max_missing;
    // the subtraction is a worst case (every missing SNP at full weight for both samples);
    // never let it shrink the normaliser below 5 % of its complete-data value
    scale = std::max(scale, 0.05 * plan.scale);
    const double budget = tol * std::max(scale, 1e-300);
    // the extended-precision W vector needs max|W| * 2^(f + W_EXTRA_BITS) inside int64
    const int wcap = 61 - W_EXTRA_BITS - (int)std::ceil(std::log2(std::max(std::max(plan.max_abs, plan.max_abs_w), 1.0)));
    const int fmax = std::min(head, wcap);
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
            bool feasible = false;
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
}

void grm_accumulate(snprel_ctx *c, int est, const snprel_plan *plan_in) {
    if (est == SNPREL_GRM_CORR) est = SNPREL_GRM_GCTA;
    ensure_stats(c);
    snprel_plan plan = *plan_in;
    const bool homo = est == SNPREL_EST_KING_HOMO;
    int nU = 0, nW = 0, nD = 0;
    choose_format(est, plan, nU, nW, nD, c->round_mode, (double)c->n_samp);
    const int f = plan.frac_bits, fw = plan.frac_bits_w, fd = plan.frac_bits_d;
    int nD2 = homo ? nD : 0;
    const int npass = nU + nW + nD + nD2;
    const int64_t cap = c->snp_cap, npad = c->n_samp_pad;

    // Digit tables, per-sample vectors and global scalars depend on the genotypes and the
    // fixed-point format only, not on the row window: a tiled N x N run builds them once.
    snprel_ctx::PrepCache &pc = c->prep_cache;
    const bool prep_hit = pc.version == c->geno_version && pc.est == est && pc.bayesian == plan.bayesian &&
                          pc.round_mode == c->round_mode &&
                          pc.f == f && pc.fw == fw && pc.fd == fd && pc.nU == nU && pc.nW == nW && pc.nD == nD &&
                          pc.nD2 == nD2 && c->scr_tab.p && c->samp_sum.p && c->scr_cnt.p;
    DevBuf<uint32_t> &tab = c->scr_tab;
    if (!prep_hit) {
        // scr_cnt was summed over the ranks together with an earlier format's vectors while the
        // cached plan statistics kept it from being rebuilt: restore this rank's own counts
        if (pc.reduced && pc.version == c->geno_version) sample_counts(c);
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
            // (the draw is keyed by the LOCAL SNP index: shards of a multi-GPU run use different SNPs, so
            //  different local indices on different ranks still give independent draws per SNP)
            const uint64_t dither = (c->round_mode == 1 && !homo) ? (0xD17E5ull + (uint64_t)c->device * 0x9E3779B97F4A7C15ull) | 1ull : 0ull;
            tables_kernel<<<tblocks, 256, 0, c->stream>>>(c->stat.p, c->n_snp, cap, est, plan.bayesian, f, fw, fd, nU,
                                                          nW, nD, nD2, tab.p, c->scr_part.p, c->iscalars.p, ovf.p, dither);
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
            dim3 grid((unsigned)((c->row_bytes + 127) / 128), (unsigned)((c->n_snp + SS_SNPS - 1) / SS_SNPS));
            sample_sum_kernel<<<grid, 128, 0, c->stream>>>(c->geno2b.p, c->stat.p, c->n_snp, c->row_bytes, npad, est,
                                                           plan.bayesian, f, fw, fd, c->samp_sum.p);
            KERNEL_CHECK(c);
        }
        pc.version = c->geno_version;
        pc.est = est;
        pc.bayesian = plan.bayesian;
        pc.f = f;
        pc.fw = fw;
        pc.fd = fd;
        pc.nU = nU;
        pc.nW = nW;
        pc.nD = nD;
        pc.nD2 = nD2;
        pc.round_mode = c->round_mode;
        pc.reduced = false;
    }

    // Gram planes: 0 = numerator, 1 = missing-pair denominator (or KING-homo d), 2 = KING-homo d2
    const int nplanes = homo ? 2 : ((nD > 0) ? 2 : 1);
    c->acc.alloc((size_t)nplanes * row_window(c).rows * npad);
    c->acc.zero(c->stream);
    c->acc_planes = nplanes;

    std::vector<GramPass> passes;
    int pass = 0;
    for (int k = 0; k < nU; k++, pass++) passes.push_back({tab.p + (int64_t)pass * cap, TABB_X, 0, 8 * k});
    for (int k = 0; k < nW; k++, pass++) passes.push_back({tab.p + (int64_t)pass * cap, TABB_M, 0, 8 * k + (f - fw)});
    for (int k = 0; k < nD; k++, pass++)
        passes.push_back({tab.p + (int64_t)pass * cap, TABB_M, homo ? 0 : 1, 8 * k});
    for (int k = 0; k < nD2; k++, pass++) passes.push_back({tab.p + (int64_t)pass * cap, TABB_M, 1, 8 * k});

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
    c->reduce_list.push_back({c->acc.p, (int64_t)c->acc.n, 0});
    if (!pc.reduced) {   // (kept from an earlier window: already summed over the ranks)
        c->reduce_list.push_back({c->samp_sum.p, (int64_t)c->samp_sum.n, 0});
        c->reduce_list.push_back({c->scalars.p, (int64_t)c->scalars.n, 2});
        c->reduce_list.push_back({c->iscalars.p, (int64_t)c->iscalars.n, 0});
        c->reduce_list.push_back({c->scr_cnt.p, (int64_t)c->scr_cnt.n, 1});   // per-sample genotype sums / missing counts
    }
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
// numerator C_ij = (acc0[i][j] - vecW_hi[i]) * 2^-f - vecW_lo[i] * 2^-(f+20), then the estimator's
// normalisation.  mode 0: Eigenstrat (scale = (n-1)/trace); 1: GCTA; 3: EIGMIX (mul = 1 ibd / 2 GRM)
struct FinalArgs {
    const long long *acc;     // [planes][win.rows][npad]
    const long long *vec;     // [NVEC][npad]
    const int *nmiss;         // [npad] missing genotypes per sample
    long long n_snp_total;
    double inv_scale, inv_scale_lo;   // 2^-f, 2^-(f + W_EXTRA_BITS)
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
    const long long q = a.acc[(i - a.win.r0) * a.npad + j] - a.vec[VEC_W * a.npad + i];
    return (double)q * a.inv_scale - (double)a.vec[VEC_WLO * a.npad + i] * a.inv_scale_lo;
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
    long long nlocus;
};
static Globals read_globals(snprel_ctx *c) {
    double hs[4];
    long long hi[4];
    d2h(c, hs, c->scalars.p, 4);
    d2h(c, hi, c->iscalars.p, 4);
    return Globals{hs[0], hs[1], hi[0]};
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
    DevBuf<double> &o = c->scr_out;
    o.alloc(win_out_count(c, packed));
    FinalArgs a{};
    a.acc = c->acc.p;
    a.vec = c->samp_sum.p;
    a.nmiss = c->scr_cnt.p + c->n_samp_pad;
    a.n_snp_total = (long long)c->plan.n_snp;
    a.inv_scale = std::ldexp(1.0, -c->plan.frac_bits);
    a.inv_scale_lo = std::ldexp(1.0, -(c->plan.frac_bits + W_EXTRA_BITS));
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
        Globals g = read_globals(c);
        a.mode = 1;
        a.nlocus = g.nlocus;
    } else {
        Globals g = read_globals(c);
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
    d2h(c, out, o.p, win_out_count(c, packed));
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
    GramPass p{tab.p, tb, 0, 0};
    c->hot_launches = 0;
    gram_tc_run(c, &p, 1, acc.p, false);
    CUDA_CHECK(cudaMemcpy2DAsync(out, (size_t)n * 8, acc.p, (size_t)npad * 8, (size_t)n * 8, (size_t)n,
                                 cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

}  // namespace snprel
