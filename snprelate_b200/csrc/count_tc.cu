// Opt-in tensor-pipe engine for the pair counters of bitcount.cu (snprel_set_count_engine):
// IBS (CIBSCount, src/genIBS.cpp:154-273), KING-robust (CKINGRobust, src/genKING.cpp:292-426) and
// IndivBeta (CIndivBeta, src/genBeta.cpp:65-178) counters as exact {-1,0,1} table Grams on K1
// (gram_tc.cu), converted back to the packed-bit kernels' uint32 counter planes so that every
// epilogue and the multi-GPU reduction stay unchanged.
#include "common.cuh"

namespace snprel {

// ---------------------------------------------------------------------------
// Pair counters on the tensor pipe (opt-in, snprel_set_count_engine): the IBS / KING-robust /
// IndivBeta counters of bitcount.cu are sums over SNPs of products of per-sample indicator
// values, i.e. table Grams with constant tables of entries in {-1, 0, 1}.  With v = valid,
// h = heterozygous, s = (1,-1,1), u = (1,0,-1), w = (1,0,1) over genotype 0/1/2 (0 for missing)
// and c_ab = #SNPs with (g_i, g_j) = (a, b):
//     v.v = nLoci          s.s = nLoci - 2 IBS1        u.u = c00 + c22 - IBS0      w.w = c00 + c22 + IBS0
//     h.v = N1_Aa (row sample het, column sample valid)          v.h = N2_Aa
// Four symmetric rank-1 forms are the minimum that span {nLoci, IBS0, IBS2}; every product is
// exact in int32/int64, so the counters are bit-identical to the packed-bit kernels' (tests).
// The packed-bit kernels stay the default (north star); this engine is 8-9x faster because the
// bit kernels are bound by the integer ALU, not by HBM.
// ---------------------------------------------------------------------------
constexpr uint32_t CT_V = 0x00010101u;   // valid
constexpr uint32_t CT_S = 0x0001FF01u;   // (1, -1, 1)
constexpr uint32_t CT_U = 0x00FF0001u;   // (1, 0, -1)
constexpr uint32_t CT_W = 0x00010001u;   // (1, 0, 1)
constexpr uint32_t CT_H = 0x00000100u;   // heterozygous

__global__ void fill_u32_kernel(uint32_t *__restrict__ p, int64_t n, uint32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// planes (int64, [np][rows][npad]) -> the packed-bit kernels' uint32 counter planes
__global__ void counts_from_planes_kernel(const long long *__restrict__ acc, uint32_t *__restrict__ cnt, int est,
                                          int64_t n, int64_t npad, RowWin win) {
    int64_t i = win.r0 + blockIdx.x, j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= n || j < i || i >= win.r1) return;
    const int64_t plane = win.rows * npad, k = (i - win.r0) * npad + j;
    const long long vv = acc[k];
    if (est == SNPREL_EST_BETA) {          // ibscnt = het_any + 2 same-homozygote = nLoci + u.u
        cnt[k] = (uint32_t)(vv + acc[plane + k]);
        cnt[plane + k] = 0;
        cnt[2 * plane + k] = (uint32_t)vv;
        return;
    }
    const long long ss = acc[plane + k], uu = acc[2 * plane + k], ww = acc[3 * plane + k];
    const long long n1 = (vv - ss) / 2, n0 = (ww - uu) / 2, n2 = vv - n1 - n0;
    if (est == SNPREL_EST_IBS) {           // ibs0, ibs2, mask
        cnt[k] = (uint32_t)n0;
        cnt[plane + k] = (uint32_t)n2;
        cnt[2 * plane + k] = (uint32_t)vv;
    } else {                               // ibs0, mask, het, N1_Aa, N2_Aa
        cnt[k] = (uint32_t)n0;
        cnt[plane + k] = (uint32_t)vv;
        cnt[2 * plane + k] = (uint32_t)n1;
        cnt[3 * plane + k] = (uint32_t)acc[4 * plane + k];
        cnt[4 * plane + k] = (uint32_t)acc[5 * plane + k];
    }
}

void tensor_count_accumulate(snprel_ctx *c, int est) {
    ensure_stats(c);   // (pads the SNP tail)
    const int64_t cap = c->snp_cap, npad = c->n_samp_pad;
    const RowWin win = row_window(c);
    struct CP { uint32_t a, b; };
    std::vector<CP> cp;
    if (est == SNPREL_EST_BETA) cp = {{CT_V, CT_V}, {CT_U, CT_U}};
    else {
        cp = {{CT_V, CT_V}, {CT_S, CT_S}, {CT_U, CT_U}, {CT_W, CT_W}};
        if (est == SNPREL_EST_KING_ROBUST) {
            cp.push_back({CT_H, CT_V});
            cp.push_back({CT_V, CT_H});
        }
    }
    const int np = (int)cp.size(), nc = est == SNPREL_EST_KING_ROBUST ? 5 : 3;
    // constant digit tables (one word per SNP row; padding rows are all-missing, entry 3 is 0)
    const uint32_t kinds[5] = {CT_V, CT_S, CT_U, CT_W, CT_H};
    DevBuf<uint32_t> &tab = c->scr_ctab;
    tab.alloc((size_t)5 * cap);
    for (int t = 0; t < 5; t++) {
        fill_u32_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, c->stream>>>(tab.p + (int64_t)t * cap, cap, kinds[t]);
        KERNEL_CHECK(c);
    }
    auto tab_of = [&](uint32_t kind) {
        for (int t = 0; t < 5; t++)
            if (kinds[t] == kind) return (const uint32_t *)(tab.p + (int64_t)t * cap);
        return (const uint32_t *)nullptr;
    };
    // own planes: KING-homo keeps its two float-sum Gram planes in c->acc while these counters run
    DevBuf<long long> &cacc = c->scr_cacc;
    cacc.alloc((size_t)np * win.rows * npad);
    cacc.zero(c->stream);
    std::vector<GramPass> passes;
    for (int p = 0; p < np; p++) passes.push_back({tab_of(cp[p].a), gram_const_table(c, cp[p].b), p, 0, 1, nullptr});
    c->cnt.alloc((size_t)nc * win.rows * npad);
    c->cnt_planes = nc;
    c->cnt.zero(c->stream);
    c->hot_launches = 0;
    CUDA_CHECK(cudaEventRecord(c->ev0, c->stream));
    gram_tc_run(c, passes.data(), np, cacc.p, true);
    dim3 grid((unsigned)(win.r1 - win.r0), (unsigned)((c->n_samp + 127) / 128));
    counts_from_planes_kernel<<<grid, 128, 0, c->stream>>>(cacc.p, c->cnt.p, est, c->n_samp, npad, win);
    KERNEL_CHECK(c);
    CUDA_CHECK(cudaEventRecord(c->ev1, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->hot_ms = ms;
    c->hot_units = 0.5 * (double)c->n_samp * (double)c->n_samp * (double)c->n_snp;
}

}  // namespace snprel
