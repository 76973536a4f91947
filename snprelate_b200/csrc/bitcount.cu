// Packed-bit pair kernels: IBS (CIBSCount::thread_ibs_num, src/genIBS.cpp:154-273),
// KING-robust (CKINGRobust::thread_ibs_num, src/genKING.cpp:292-426) and IndivBeta
// (CIndivBeta::thread_ibs_num, src/genBeta.cpp:65-178) counters, plus their
// epilogues (gnrIBSAve/gnrIBSNum src/genIBS.cpp:441-550, gnrIBD_KING_Robust
// src/genKING.cpp:576-679, gnrIBD_Beta / CalcIndivBetaGRM src/genBeta.cpp:263-460).
//
// Input: bit planes [word64][sample] of uint4 {p1.lo,p1.hi,p2.lo,p2.hi}
// (geno.cu:planes_kernel).  One CTA owns a 64 x 64 tile of sample pairs, streams
// the two 64-sample panels through shared memory with cp.async (16-byte chunks,
// double buffered) and keeps a 4 x 4 block of pairs per thread in registers.
// All arithmetic is integer (XOR/AND + __popc); counters are exact uint32 like
// the reference's (TIBS / TS_KINGRobust / TS_Beta).
#include "common.cuh"

namespace snprel {

constexpr int BT = 64;        // pairs tile edge
constexpr int BKW = 9;        // 64-SNP words per pipeline stage (three carry-save groups of 3)
constexpr int BTHREADS = 256;
constexpr int PAIR_DIRECT_DEFAULT = 0;   // streams that bypass the carry-save adder by default (see pair_update3)

template <int EST> struct EstTraits;
template <> struct EstTraits<SNPREL_EST_IBS> { static constexpr int NC = 3; };          // ibs0, ibs2, mask
template <> struct EstTraits<SNPREL_EST_KING_ROBUST> { static constexpr int NC = 5; };  // ibs0, mask, het, Aa1, Aa2
template <> struct EstTraits<SNPREL_EST_BETA> { static constexpr int NC = 3; };         // het, ibs2, mask

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// the bit streams whose populations are the estimator's counters, for 32 SNPs of one pair.
// vb = validity of the column sample (b1 | ~b2), hb = its heterozygosity (b1 & ~b2): both depend on
// the column word only and are computed once per word, outside the loop over the four row samples,
// so that the pair's mask is ONE three-input LOP3, (a1 | ~a2) & vb.
// one three-input logic op with an explicit truth table (index bit = a<<2 | b<<1 | c); spelled in
// PTX because nvcc re-associates C-level boolean expressions into a form that needs 8.3 LOP3 per word
// pair where 7 suffice (profiles/r01_pair_count_sass_mix.md)
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
}
constexpr int TA = 0xF0, TB = 0xCC, TC = 0xAA;   // truth-table columns of the three inputs

template <int EST>
__device__ __forceinline__ void pair_streams(uint32_t (&st)[EstTraits<EST>::NC], uint32_t a1, uint32_t a2,
                                             uint32_t b1, uint32_t b2, uint32_t vb, uint32_t hb) {
    // validity: a genotype is missing iff (plane1, plane2) == (0, 1)
    const uint32_t mask = lop3<(TA | (~TB & 0xFF)) & TC>(a1, a2, vb);            // (a1 | ~a2) & vb
    if (EST == SNPREL_EST_IBS) {
        const uint32_t t = lop3<(TA ^ TB) & TC>(a1, b1, mask);                   // plane 1 differs, both valid
        st[0] = lop3<TA & (TB ^ TC)>(t, a2, b2);                                 // ibs0: (0,0) vs (1,1)
        const uint32_t u = lop3<(~(TA ^ TB) & 0xFF) & TC>(a1, b1, mask);         // plane 1 equal, both valid
        st[1] = lop3<TA & (~(TB ^ TC) & 0xFF)>(u, a2, b2);                       // ibs2: identical
        st[2] = mask;
    } else if (EST == SNPREL_EST_KING_ROBUST) {
        const uint32_t t = lop3<(TA ^ TB) & TC>(a1, b1, mask);
        st[0] = lop3<TA & (TB ^ TC)>(t, a2, b2);                                 // ibs0
        st[1] = mask;                                                            // nLoci
        const uint32_t pa = lop3<TA ^ TB ^ TC>(a1, a2, b1);
        st[2] = lop3<(TA ^ TB) & TC>(pa, b2, mask);                              // het: exactly one of the two is Aa
        st[3] = lop3<TA & (~TB & 0xFF) & TC>(a1, a2, vb);                        // N1_Aa (row sample)
        st[4] = lop3<TA & (TB | (~TC & 0xFF))>(hb, a1, a2);                      // N2_Aa (column sample)
    } else {
        const uint32_t h1 = lop3<(TA ^ TB) | TC>(a1, a2, hb);                    // either heterozygous
        st[0] = h1 & mask;
        const uint32_t v = mask & ~h1;
        st[1] = lop3<TA & (~(TB ^ TC) & 0xFF)>(v, a1, b1);                       // same homozygote
        st[2] = mask;
    }
}

// POPC issues on the quarter-rate XU pipe and bounded the first version of this kernel (96 % busy,
// profiles/r01_ibs_pair_count_full.txt).  Three words of every stream go through a carry-save
// adder first (sum = a^b^c, carry = maj(a,b,c): two LOP3 on the full-rate ALU pipe), so three
// populations cost two POPCs: pop(a)+pop(b)+pop(c) = pop(sum) + 2 pop(carry).
// DIRECT: the last DIRECT streams skip the carry-save adder and are popcounted word by word (3 POPC + 3 IMAD
// instead of 2 LOP3 + 2 POPC + 2 IMAD per three words): it moves load from the ALU pipe to the XU / FMA pipes.
template <int EST, int DIRECT>
__device__ __forceinline__ void pair_update3(uint32_t (&acc)[EstTraits<EST>::NC], const uint32_t (&a1)[3],
                                             const uint32_t (&a2)[3], const uint32_t (&b1)[3],
                                             const uint32_t (&b2)[3], const uint32_t (&vb)[3],
                                             const uint32_t (&hb)[3]) {
    constexpr int NC = EstTraits<EST>::NC;
    uint32_t s0[NC], s1[NC], s2[NC];
    pair_streams<EST>(s0, a1[0], a2[0], b1[0], b2[0], vb[0], hb[0]);
    pair_streams<EST>(s1, a1[1], a2[1], b1[1], b2[1], vb[1], hb[1]);
    pair_streams<EST>(s2, a1[2], a2[2], b1[2], b2[2], vb[2], hb[2]);
#pragma unroll
    for (int k = 0; k < NC; k++) {
        if (k >= NC - DIRECT) {
            uint32_t t, u;
            asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(t) : "r"((uint32_t)__popc(s0[k])), "r"(acc[k]));
            asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(u) : "r"((uint32_t)__popc(s1[k])), "r"(t));
            asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(acc[k]) : "r"((uint32_t)__popc(s2[k])), "r"(u));
            continue;
        }
        uint32_t sum = s0[k] ^ s1[k] ^ s2[k];
        uint32_t carry = (s0[k] & s1[k]) | (s2[k] & (s0[k] ^ s1[k]));
        // acc += pop(sum) + 2 pop(carry) as two IMADs: the FMA pipe is idle in this kernel while the
        // ALU pipe (LOP3, IADD3) is the bound (profiles/r01_pair_count_sass_mix.md)
        uint32_t t;
        asm("mad.lo.u32 %0, %1, 2, %2;" : "=r"(t) : "r"((uint32_t)__popc(carry)), "r"(acc[k]));
        asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(acc[k]) : "r"((uint32_t)__popc(sum)), "r"(t));
    }
}

template <int EST, int DIRECT>
__global__ void __launch_bounds__(BTHREADS, EstTraits<EST>::NC <= 3 ? 2 : 1)
pair_count_kernel(const uint4 *__restrict__ planes, uint32_t *__restrict__ cnt, int64_t n_pad,
                  int64_t n_words, int words_per_split, RowWin win) {
    constexpr int NC = EstTraits<EST>::NC;
    const int ti = blockIdx.y + (int)(win.r0 / BT), tj = blockIdx.x;
    if (ti > tj) return;   // upper block triangle only
    __shared__ uint4 sA[2][BKW][BT];
    __shared__ uint4 sB[2][BKW][BT];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t w_begin = (int64_t)blockIdx.z * words_per_split;
    const int64_t w_end = min(n_words, w_begin + (int64_t)words_per_split);
    const int64_t i0 = (int64_t)ti * BT, j0 = (int64_t)tj * BT;

    uint32_t acc[4][4][NC];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int q = 0; q < 4; q++)
#pragma unroll
            for (int k = 0; k < NC; k++) acc[r][q][k] = 0;

    auto load_stage = [&](int buf, int64_t w0) {
        // BKW*BT = 576 uint4 per panel; words past the end read as all-missing (plane1 = 0, plane2 = 1)
        for (int e = tid; e < BKW * BT; e += BTHREADS) {
            int w = e >> 6, s = e & 63;
            int64_t gw = w0 + w;
            if (gw < w_end) {
                cp_async16(&sA[buf][w][s], planes + gw * n_pad + i0 + s);
                cp_async16(&sB[buf][w][s], planes + gw * n_pad + j0 + s);
            } else {
                sA[buf][w][s] = make_uint4(0u, 0u, 0xFFFFFFFFu, 0xFFFFFFFFu);
                sB[buf][w][s] = make_uint4(0u, 0u, 0xFFFFFFFFu, 0xFFFFFFFFu);
            }
        }
    };

    int buf = 0;
    if (w_begin < w_end) load_stage(0, w_begin);
    cp_async_commit();
    for (int64_t w0 = w_begin; w0 < w_end; w0 += BKW) {
        if (w0 + BKW < w_end) load_stage(buf ^ 1, w0 + BKW);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
#pragma unroll 1
        for (int w = 0; w < BKW; w += 3) {
            uint4 a[4][3];
#pragma unroll
            for (int k = 0; k < 3; k++)
#pragma unroll
                for (int r = 0; r < 4; r++) a[r][k] = sA[buf][w + k][ty + 16 * r];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint4 b[3];   // one column sample at a time keeps the live set under 128 registers
#pragma unroll
                for (int k = 0; k < 3; k++) b[k] = sB[buf][w + k][tx + 16 * q];
                {   // low 32 SNPs of the three words
                    const uint32_t b1[3] = {b[0].x, b[1].x, b[2].x}, b2[3] = {b[0].z, b[1].z, b[2].z};
                    const uint32_t vb[3] = {b1[0] | ~b2[0], b1[1] | ~b2[1], b1[2] | ~b2[2]};
                    const uint32_t hb[3] = {b1[0] & ~b2[0], b1[1] & ~b2[1], b1[2] & ~b2[2]};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const uint32_t a1[3] = {a[r][0].x, a[r][1].x, a[r][2].x}, a2[3] = {a[r][0].z, a[r][1].z, a[r][2].z};
                        pair_update3<EST, DIRECT>(acc[r][q], a1, a2, b1, b2, vb, hb);
                    }
                }
                {   // high 32 SNPs
                    const uint32_t b1[3] = {b[0].y, b[1].y, b[2].y}, b2[3] = {b[0].w, b[1].w, b[2].w};
                    const uint32_t vb[3] = {b1[0] | ~b2[0], b1[1] | ~b2[1], b1[2] | ~b2[2]};
                    const uint32_t hb[3] = {b1[0] & ~b2[0], b1[1] & ~b2[1], b1[2] & ~b2[2]};
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const uint32_t a1[3] = {a[r][0].y, a[r][1].y, a[r][2].y}, a2[3] = {a[r][0].w, a[r][1].w, a[r][2].w};
                        pair_update3<EST, DIRECT>(acc[r][q], a1, a2, b1, b2, vb, hb);
                    }
                }
            }
        }
        __syncthreads();
        buf ^= 1;
    }
    cp_async_wait<0>();

    const int64_t plane = win.rows * n_pad;
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int64_t i = i0 + ty + 16 * r, j = j0 + tx + 16 * q;
            if (j < i) continue;   // only the upper triangle is kept (CdMatTri)
#pragma unroll
            for (int k = 0; k < NC; k++)
                if (acc[r][q][k]) atomicAdd(cnt + k * plane + (i - win.r0) * n_pad + j, acc[r][q][k]);
        }
}

template <int EST>
static void launch_pair_count(snprel_ctx *c) {
    const int64_t npad = c->n_samp_pad;
    const RowWin win = row_window(c);
    const int64_t tiles = (c->n_samp + BT - 1) / BT;
    const int64_t wtiles = (win.r1 - win.r0 + BT - 1) / BT;          // tile rows of the window
    const int64_t n_words = c->plane_words;
    // split the SNP words when the tile grid alone cannot fill the chip
    int64_t ntile = std::max<int64_t>(1, wtiles * (2 * (tiles - win.r0 / BT) - wtiles + 1) / 2);
    int64_t want = (int64_t)c->num_sms * 4;
    int64_t splits = std::max<int64_t>(1, std::min<int64_t>((want + ntile - 1) / ntile,
                                                           (n_words + BKW - 1) / BKW));
    splits = std::min<int64_t>(splits, 65535);
    int64_t wps = round_up((n_words + splits - 1) / splits, BKW);
    splits = (n_words + wps - 1) / wps;
    dim3 grid((unsigned)tiles, (unsigned)wtiles, (unsigned)splits);
    // streams popcounted without the carry-save adder (experiment knob: debug flags 0x400 / 0x800 / 0xC00 = 1 / 2 / 3)
    const int direct = (c->debug_flags & 0xC00u) ? (int)((c->debug_flags >> 10) & 3u) : PAIR_DIRECT_DEFAULT;
    switch (direct) {
        case 1: pair_count_kernel<EST, 1><<<grid, BTHREADS, 0, c->stream>>>(c->planes.p, c->cnt.p, npad, n_words, (int)wps, win); break;
        case 2: pair_count_kernel<EST, 2><<<grid, BTHREADS, 0, c->stream>>>(c->planes.p, c->cnt.p, npad, n_words, (int)wps, win); break;
        case 3: pair_count_kernel<EST, 3><<<grid, BTHREADS, 0, c->stream>>>(c->planes.p, c->cnt.p, npad, n_words, (int)wps, win); break;
        default: pair_count_kernel<EST, 0><<<grid, BTHREADS, 0, c->stream>>>(c->planes.p, c->cnt.p, npad, n_words, (int)wps, win); break;
    }
    KERNEL_CHECK(c);
}

static void finish_accumulate(snprel_ctx *c, int est) {
    c->accum_win_r0 = row_window(c).r0;
    c->accum_win_rows = row_window(c).rows;
    c->accum_est = est;
    c->accum_reduced = false;
    c->reduce_list.clear();
    c->reduce_list.push_back({c->cnt.p, (int64_t)c->cnt_planes * row_window(c).rows * c->n_samp_pad, 1, c->n_samp_pad,
                              row_window(c).rows, row_window(c).r0});
}

void bitcount_accumulate(snprel_ctx *c, int est) {
    if (est != SNPREL_EST_IBS && est != SNPREL_EST_KING_ROBUST && est != SNPREL_EST_BETA)
        fail("internal: bad packed-bit estimator %d", est);
    if (est == SNPREL_EST_KING_ROBUST && c->n_snp >= 1073741824ll)
        fail("The number of SNPs should be less than 1,073,741,824.");   // src/genKING.cpp:598
    if (c->count_engine == 1) {   // same counters from the tensor pipe (count_tc.cu), opt-in
        tensor_count_accumulate(c, est);
        finish_accumulate(c, est);
        return;
    }
    ensure_planes(c);
    int nc = est == SNPREL_EST_KING_ROBUST ? 5 : 3;
    c->cnt.alloc((size_t)nc * row_window(c).rows * c->n_samp_pad);
    c->cnt_planes = nc;
    c->cnt.zero(c->stream);
    CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
    switch (est) {
        case SNPREL_EST_IBS: launch_pair_count<SNPREL_EST_IBS>(c); break;
        case SNPREL_EST_KING_ROBUST: launch_pair_count<SNPREL_EST_KING_ROBUST>(c); break;
        case SNPREL_EST_BETA: launch_pair_count<SNPREL_EST_BETA>(c); break;
        default: fail("internal: bad packed-bit estimator %d", est);
    }
    CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->hot_ms = ms;
    c->hot_launches = 1;
    c->hot_units = 0.5 * (double)c->n_samp * (double)c->n_samp * (double)c->n_snp;
    finish_accumulate(c, est);
}

// ---------------------------------------------------------------------------
// epilogues
// ---------------------------------------------------------------------------

// out index helpers: full symmetric n x n (column-major == row-major) or the
// row-packed upper triangle idx(r,c) = c + r(2n-r-1)/2 (src/dGenGWAS.h:556-561)
__device__ __forceinline__ void store_sym(double *out, int packed, int64_t n, int64_t i,
                                          int64_t j, double v, long long pbase = 0) {
    if (packed) {
        out[tri_idx(n, i, j) - pbase] = v;
    } else {
        out[i * n + j] = v;
        out[j * n + i] = v;
    }
}

__global__ void ibs_num_kernel(const uint32_t *__restrict__ cnt, int32_t *__restrict__ o0,
                               int32_t *__restrict__ o1, int32_t *__restrict__ o2, int64_t n,
                               int64_t npad, RowWin win, int packed) {
    int64_t i = win.r0 + blockIdx.x, j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= n || j < i) return;
    int64_t plane = win.rows * npad, k = (i - win.r0) * npad + j;
    uint32_t n0 = cnt[k], n2 = cnt[plane + k], nm = cnt[2 * plane + k];
    uint32_t n1 = nm - n0 - n2;
    if (packed) {
        long long t = tri_idx(n, i, j) - win.pbase;
        o0[t] = (int32_t)n0;
        o1[t] = (int32_t)n1;
        o2[t] = (int32_t)n2;
    } else {
        o0[i * n + j] = o0[j * n + i] = (int32_t)n0;
        o1[i * n + j] = o1[j * n + i] = (int32_t)n1;
        o2[i * n + j] = o2[j * n + i] = (int32_t)n2;
    }
}

__global__ void ibs_ave_kernel(const uint32_t *__restrict__ cnt, double *__restrict__ out,
                               int packed, int64_t n, int64_t npad, RowWin win) {
    int64_t i = win.r0 + blockIdx.x, j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= n || j < i) return;
    int64_t plane = win.rows * npad, k = (i - win.r0) * npad + j;
    uint32_t n0 = cnt[k], n2 = cnt[plane + k], nm = cnt[2 * plane + k];
    uint32_t n1 = nm - n0 - n2;
    // (0.5*IBS1 + IBS2) / (IBS0 + IBS1 + IBS2), src/genIBS.cpp:472-473
    double v = (0.5 * (double)n1 + (double)n2) / (double)(n0 + n1 + n2);
    store_sym(out, packed, n, i, j, v, win.pbase);
}

__global__ void king_robust_kernel(const uint32_t *__restrict__ cnt,
                                   const int32_t *__restrict__ fam, double *__restrict__ oibs0,
                                   double *__restrict__ okin, int packed, int64_t n, int64_t npad,
                                   RowWin win) {
    int64_t i = win.r0 + blockIdx.x, j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= n || j < i) return;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    if (i == j) {   // src/genKING.cpp:628
        store_sym(oibs0, packed, n, i, j, 0.0, win.pbase);
        store_sym(okin, packed, n, i, j, 0.5, win.pbase);
        return;
    }
    int64_t plane = win.rows * npad, k = (i - win.r0) * npad + j;
    uint32_t ibs0 = cnt[k], nloci = cnt[plane + k], het = cnt[2 * plane + k];
    uint32_t n1 = cnt[3 * plane + k], n2 = cnt[4 * plane + k];
    uint32_t sumsq = het + 4u * ibs0;   // sum (g_i - g_j)^2, src/genKING.cpp:421
    double r0 = nloci > 0 ? (double)ibs0 / (double)nloci : nan;
    int f1 = fam ? fam[i] : SNPREL_NA_INT, f2 = fam ? fam[j] : SNPREL_NA_INT;
    double v;
    if (f1 == f2 && f1 != SNPREL_NA_INT)
        v = 0.5 - (double)sumsq / (2.0 * (double)(n1 + n2));
    else
        v = 0.5 - (double)sumsq / (4.0 * (double)min(n1, n2));
    if (!isfinite(v)) v = nan;
    store_sym(oibs0, packed, n, i, j, r0, win.pbase);
    store_sym(okin, packed, n, i, j, v, win.pbase);
}

// IBD::Est_PLINK_Kinship (src/genIBD.cpp:341-383) per pair on the IBS counters; the diagonal is
// k0 = k1 = 0 (src/genIBS.cpp:596,614).  e = {E00, E01, E02, E11, E12} (E22 = 1).
struct MomTab { double e00, e01, e02, e11, e12; };

__global__ void ibd_mom_kernel(const uint32_t *__restrict__ cnt, double *__restrict__ ok0,
                               double *__restrict__ ok1, MomTab t, int constraint, int packed,
                               int64_t n, int64_t npad, RowWin win) {
    int64_t i = win.r0 + blockIdx.x, j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= n || j < i) return;
    if (i == j) {
        store_sym(ok0, packed, n, i, j, 0.0, win.pbase);
        store_sym(ok1, packed, n, i, j, 0.0, win.pbase);
        return;
    }
    int64_t plane = win.rows * npad, k = (i - win.r0) * npad + j;
    uint32_t n0 = cnt[k], n2 = cnt[plane + k], nm = cnt[2 * plane + k];
    uint32_t n1 = nm - n0 - n2;
    double tot = (double)(int)nm;
    double e00 = t.e00 * tot, e01 = t.e01 * tot, e11 = t.e11 * tot;
    double e02 = t.e02 * tot, e12 = t.e12 * tot, e22 = 1.0 * tot;
    double k0 = __ddiv_rn((double)n0, e00);
    double k1 = __ddiv_rn(__dsub_rn((double)n1, __dmul_rn(k0, e01)), e11);
    double k2 = __ddiv_rn(__dsub_rn(__dsub_rn((double)n2, __dmul_rn(k0, e02)), __dmul_rn(k1, e12)), e22);
    if (k0 > 1) { k0 = 1; k1 = k2 = 0; }
    if (k1 > 1) { k1 = 1; k0 = k2 = 0; }
    if (k2 > 1) { k2 = 1; k0 = k1 = 0; }
    if (k0 < 0) { double S = k1 + k2; k1 /= S; k2 /= S; k0 = 0; }
    if (k1 < 0) { double S = k0 + k2; k0 /= S; k2 /= S; k1 = 0; }
    if (k2 < 0) { double S = k0 + k1; k0 /= S; k1 /= S; k2 = 0; }
    if (constraint) {
        k2 = __dsub_rn(__dsub_rn(1.0, k0), k1);
        double pihat = __dadd_rn(k1 / 2, k2);
        if (__dmul_rn(pihat, pihat) < k2) {
            k0 = __dmul_rn(1 - pihat, 1 - pihat);
            k1 = __dmul_rn(__dmul_rn(2.0, pihat), 1 - pihat);
        }
    }
    store_sym(ok0, packed, n, i, j, k0, win.pbase);
    store_sym(ok1, packed, n, i, j, k1, win.pbase);
}

// copy counter planes to full symmetric int32 matrices (row = first sample)
__global__ void counts_sym_kernel(const uint32_t *__restrict__ cnt, int32_t *__restrict__ out,
                                  int est, int nplanes_out, int64_t n, int64_t npad) {
    int64_t i = blockIdx.x, j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= n || j < i) return;
    int64_t plane = npad * npad, k = i * npad + j, nn = n * n;
    if (est == SNPREL_EST_KING_ROBUST) {
        uint32_t ibs0 = cnt[k], nloci = cnt[plane + k], het = cnt[2 * plane + k];
        uint32_t n1 = cnt[3 * plane + k], n2 = cnt[4 * plane + k];
        uint32_t v[5] = {ibs0, nloci, het + 4u * ibs0, n1, n2};
        uint32_t vt[5] = {ibs0, nloci, het + 4u * ibs0, n2, n1};   // roles swap below the diagonal
        for (int p = 0; p < 5; p++) {
            out[p * nn + i * n + j] = (int32_t)v[p];
            out[p * nn + j * n + i] = (int32_t)vt[p];
        }
    } else {   // beta: ibscnt = het + 2*ibs2, num
        uint32_t ibscnt = cnt[k] + 2u * cnt[plane + k], num = cnt[2 * plane + k];
        out[i * n + j] = out[j * n + i] = (int32_t)ibscnt;
        out[nn + i * n + j] = out[nn + j * n + i] = (int32_t)num;
    }
}

// IndivBeta pass 1: raw beta into the upper triangle of a full n x n buffer and the
// per-row sums / minima of the off-diagonal (diag handled separately)
__global__ void beta_raw_kernel(const uint32_t *__restrict__ cnt, double *__restrict__ raw,
                                double *__restrict__ rowsum, double *__restrict__ rowmin,
                                int diag_inbreeding, int64_t n, int64_t npad) {
    int64_t i = blockIdx.x;
    int64_t plane = npad * npad;
    double s = 0, mn = __longlong_as_double(0x7ff0000000000000ll);
    for (int64_t j = i + threadIdx.x; j < n; j += blockDim.x) {
        int64_t k = i * npad + j;
        double ibscnt = (double)(cnt[k] + 2u * cnt[plane + k]), num = (double)cnt[2 * plane + k];
        double v;
        if (j == i) {
            v = diag_inbreeding ? ibscnt / num - 1 : (0.5 * ibscnt) / num;
        } else {
            v = (0.5 * ibscnt) / num;
            s += v;
        }
        mn = fmin(mn, v);   // the GRM flavour takes the minimum over diag + off-diag
        raw[i * n + j] = v;
    }
    __shared__ double ss[256], sm[256];
    ss[threadIdx.x] = s;
    sm[threadIdx.x] = mn;
    __syncthreads();
    for (int o = blockDim.x / 2; o; o >>= 1) {
        if (threadIdx.x < o) {
            ss[threadIdx.x] += ss[threadIdx.x + o];
            sm[threadIdx.x] = fmin(sm[threadIdx.x], sm[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        rowsum[i] = ss[0];
        rowmin[i] = sm[0];
    }
}

// IndivBeta pass 2: (beta - shift) * scale, GRM flavour diag * 0.5 + 1
__global__ void beta_final_kernel(const double *__restrict__ raw, double *__restrict__ out,
                                  int packed, double shift, double scale, int grm_flavour,
                                  int64_t n) {
    int64_t i = blockIdx.x, j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= n || j < i) return;
    double v = (raw[i * n + j] - shift) * scale;
    if (grm_flavour && i == j) v = v * 0.5 + 1;
    store_sym(out, packed, n, i, j, v);
}

static void need_accum(snprel_ctx *c, int est) {
    const RowWin w = row_window(c);   // reduced across ranks already, for this very window
    if (c->accum_est == est && c->accum_reduced && c->accum_win_r0 == w.r0 && c->accum_win_rows == w.rows) return;
    bitcount_accumulate(c, est);
}

static dim3 tri_grid(int64_t n) { return dim3((unsigned)n, (unsigned)((n + 127) / 128)); }
// rows of the current window x column chunks
static dim3 win_grid(snprel_ctx *c) {
    RowWin w = row_window(c);
    return dim3((unsigned)(w.r1 - w.r0), (unsigned)((c->n_samp + 127) / 128));
}
static void need_packed_in_window(snprel_ctx *c, int packed, const char *who) {
    if (!full_window(c) && !packed) fail("%s: a row window returns the packed upper triangle only (useMatrix)", who);
}
static size_t win_out_count(snprel_ctx *c, int packed) {
    if (!full_window(c)) return window_packed_count(c);
    int64_t n = c->n_samp;
    return packed ? (size_t)n * (n + 1) / 2 : (size_t)n * n;
}

static void check_grid_rows(int64_t n) {
    if (n > 2147483647ll) fail("too many samples for the epilogue grid");
}

template <class T>
static void d2h(snprel_ctx *c, T *host, const T *dev, size_t count) {
    CUDA_CHECK(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

static size_t out_count(int64_t n, int packed) {
    return packed ? (size_t)n * (n + 1) / 2 : (size_t)n * n;
}

void ibs_num_finish(snprel_ctx *c, int32_t *i0, int32_t *i1, int32_t *i2) {
    if (!i0 || !i1 || !i2) fail("snprel_ibs_num: NULL output");
    need_accum(c, SNPREL_EST_IBS);
    int64_t n = c->n_samp;
    check_grid_rows(n);
    const int packed = full_window(c) ? 0 : 1;   // a row window yields packed int32 slices
    const size_t oc = win_out_count(c, packed);
    DevBuf<int32_t> o;
    o.alloc(3 * oc);
    ibs_num_kernel<<<win_grid(c), 128, 0, c->stream>>>(c->cnt.p, o.p, o.p + oc, o.p + 2 * oc, n, c->n_samp_pad,
                                                       row_window(c), packed);
    KERNEL_CHECK(c);
    d2h(c, i0, o.p, oc);
    d2h(c, i1, o.p + oc, oc);
    d2h(c, i2, o.p + 2 * oc, oc);
}

void ibs_ave_finish(snprel_ctx *c, double *out, int packed) {
    if (!out) fail("snprel_ibs_ave: NULL output");
    need_packed_in_window(c, packed, "snprel_ibs_ave");
    need_accum(c, SNPREL_EST_IBS);
    int64_t n = c->n_samp;
    check_grid_rows(n);
    output_wait(c);                       // the scratch may still be on its way to the host
    DevBuf<double> &o = c->scr_out;       // persistent: no 10 GB cudaMalloc / cudaFree per row window
    o.alloc(win_out_count(c, packed));
    ibs_ave_kernel<<<win_grid(c), 128, 0, c->stream>>>(c->cnt.p, o.p, packed, n, c->n_samp_pad, row_window(c));
    KERNEL_CHECK(c);
    deliver(c, out, o.p, win_out_count(c, packed));
}

void king_robust_finish(snprel_ctx *c, const int32_t *fam, double *ibs0, double *kin, int packed) {
    if (!ibs0 || !kin) fail("snprel_king_robust: NULL output");
    need_packed_in_window(c, packed, "snprel_king_robust");
    need_accum(c, SNPREL_EST_KING_ROBUST);
    int64_t n = c->n_samp;
    check_grid_rows(n);
    DevBuf<int32_t> dfam;
    if (fam) {
        dfam.alloc((size_t)n);
        CUDA_CHECK(cudaMemcpyAsync(dfam.p, fam, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice,
                                   c->stream));
    }
    output_wait(c);
    DevBuf<double> &o = c->scr_out;
    size_t oc = win_out_count(c, packed);
    o.alloc(2 * oc);
    king_robust_kernel<<<win_grid(c), 128, 0, c->stream>>>(c->cnt.p, fam ? dfam.p : nullptr, o.p, o.p + oc,
                                                           packed, n, c->n_samp_pad, row_window(c));
    KERNEL_CHECK(c);
    if (fam) CUDA_CHECK(cudaStreamSynchronize(c->stream));    // dfam is freed when this function returns
    deliver(c, ibs0, o.p, oc, false);
    deliver(c, kin, o.p + oc, oc, true);
}

// gnrIBD_PLINK (src/genIBS.cpp:558-639) from the IBS counters and the summed per-SNP terms
void ibd_mom_finish(snprel_ctx *c, const double *sums, int constraint, double *k0, double *k1, int packed) {
    if (!sums || !k0 || !k1) fail("snprel_ibd_mom: NULL argument");
    need_packed_in_window(c, packed, "snprel_ibd_mom");
    need_accum(c, SNPREL_EST_IBS);
    int64_t n = c->n_samp;
    check_grid_rows(n);
    const double nv = sums[5];
    MomTab t = {sums[0] / nv, sums[1] / nv, sums[2] / nv, sums[3] / nv, sums[4] / nv};
    DevBuf<double> o;
    size_t oc = win_out_count(c, packed);
    o.alloc(2 * oc);
    ibd_mom_kernel<<<win_grid(c), 128, 0, c->stream>>>(c->cnt.p, o.p, o.p + oc, t, constraint, packed, n,
                                                        c->n_samp_pad, row_window(c));
    KERNEL_CHECK(c);
    d2h(c, k0, o.p, oc);
    d2h(c, k1, o.p + oc, oc);
}

static void counts_finish(snprel_ctx *c, int est, int np, int32_t *out) {
    if (!out) fail("NULL output");
    if (!full_window(c)) fail("raw counter matrices are not available inside a row window");
    need_accum(c, est);
    int64_t n = c->n_samp;
    check_grid_rows(n);
    DevBuf<int32_t> o;
    o.alloc((size_t)np * n * n);
    counts_sym_kernel<<<tri_grid(n), 128, 0, c->stream>>>(c->cnt.p, o.p, est, np, n, c->n_samp_pad);
    KERNEL_CHECK(c);
    d2h(c, out, o.p, (size_t)np * n * n);
}

void king_robust_counts_finish(snprel_ctx *c, int32_t *out5) {
    counts_finish(c, SNPREL_EST_KING_ROBUST, 5, out5);
}
void beta_counts_finish(snprel_ctx *c, int32_t *out2) { counts_finish(c, SNPREL_EST_BETA, 2, out2); }

// gnrIBD_Beta (src/genBeta.cpp:361-460) when grm_flavour == 0,
// CalcIndivBetaGRM (src/genBeta.cpp:308-357) when grm_flavour != 0
void indiv_beta_finish(snprel_ctx *c, int inbreeding, int grm_flavour, double *out, int packed,
                       double *avg_out) {
    if (!out) fail("snprel_indiv_beta: NULL output");
    if (!full_window(c)) fail("snprel_indiv_beta: needs the whole matrix (global average), not a row window");
    need_accum(c, SNPREL_EST_BETA);
    int64_t n = c->n_samp;
    check_grid_rows(n);
    DevBuf<double> raw, rs, rm, o;
    raw.alloc((size_t)n * n);
    rs.alloc((size_t)n);
    rm.alloc((size_t)n);
    beta_raw_kernel<<<(unsigned)n, 256, 0, c->stream>>>(c->cnt.p, raw.p, rs.p, rm.p,
                                                        grm_flavour ? 1 : inbreeding, n,
                                                        c->n_samp_pad);
    KERNEL_CHECK(c);
    std::vector<double> hs((size_t)n), hm((size_t)n);
    d2h(c, hs.data(), rs.p, (size_t)n);
    d2h(c, hm.data(), rm.p, (size_t)n);
    double avg = 0, mn = hm[0];
    for (int64_t i = 0; i < n; i++) {
        avg += hs[i];
        if (hm[i] < mn) mn = hm[i];
    }
    avg /= (double)((long long)n * (n - 1) / 2);
    if (avg_out) *avg_out = avg;
    double shift, scale;
    if (grm_flavour) {
        shift = mn;
        scale = 2.0 / (1 - mn);
    } else {
        shift = avg;
        scale = 1.0 / (1 - avg);
    }
    o.alloc(out_count(n, packed));
    beta_final_kernel<<<tri_grid(n), 128, 0, c->stream>>>(raw.p, o.p, packed, shift, scale,
                                                          grm_flavour, n);
    KERNEL_CHECK(c);
    d2h(c, out, o.p, out_count(n, packed));
}

}  // namespace snprel
