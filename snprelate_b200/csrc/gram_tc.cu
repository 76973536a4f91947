// K1: the tcgen05 "table Gram" -- the tensor-core hot loop that replaces
// CProdMat_AlgArith::MulAdd (src/genPCA.cpp:229-312) together with its block
// preparation (TransposeGenotype / GenoSub / GenoMul, src/genPCA.h:93-108,
// src/genPCA.cpp:315-368) and the missing-pair denominator loops
// (src/genPCA.cpp:1201-1224, src/genEIGMIX.cpp:113-138).
//
// It computes, for up to two "passes" p that share one B table,
//
//     Acc_p[i][j] = sum over SNPs l of  tabA_p[l][g_il] * tabB[g_jl]      (exact, int32)
//
// where g is the 2-bit genotype code, tabA_p[l] is a per-SNP table of four int8
// digits and tabB a table of four small int8 values, and adds Acc_p << shift_p into
// an int64 fixed-point output plane with 64-bit atomics.  The per-SNP real weights
// (1/(p(1-p)), 2p, 4p(1-p) ...) live in the digits of tabA: grm.cu slices each
// table value into balanced base-256 digits, one pass per digit, so the sum over
// passes reproduces the float64 result to a proven bound while every tensor-core
// product and every accumulation is exact integer arithmetic (order independent,
// hence bit-identical for any tiling, split or GPU count).
//
// Mapping onto sm_100a:
//   * one CTA per (128 x 256 sample tile, SNP split); 10 warps.
//   * warp 1 is the loader: per stage it issues six TMA box copies (16 bytes x 128 SNP rows
//     each, cp.async.bulk.tensor.2d) of the packed 2-bit genotypes plus bulk copies of the
//     digit tables into a small shared-memory ring, signalled through mbarrier complete_tx.
//   * warps 2..9 are producers: each thread takes 64 packed genotypes of one SNP from
//     the ring, turns every 32-bit word (16 samples) into byte-permute
//     selectors (3 logic ops + 2 shifts) and emits int8 operand rows with PRMT
//     against the 4-entry tables -- the "unpack, centre and scale on the fly" step.
//     Operands are written MN-major (16 consecutive samples = one 16-byte core row),
//     the layout in which a packed genotype word expands without any transpose.
//   * warp 0 issues tcgen05.mma.kind::i8 (M=128, N=256, K=32) from shared-memory
//     descriptors into two 128x256 int32 accumulators that fill the 512 TMEM columns;
//     tcgen05.commit releases pipeline stages back to the producers through mbarriers.
//   * after the last SNP stage the producer warps become the epilogue: tcgen05.ld the
//     accumulators and issue the 64-bit atomics.
#include "common.cuh"
#include "tc_ptx.cuh"   // CUtensorMap types; the encoder is fetched through cudaGetDriverEntryPoint

namespace snprel {
namespace tc {

constexpr int TM = 128;            // A rows (samples) per tile
constexpr int TN = 256;            // B rows (samples) per tile
constexpr int SK = 128;            // SNPs per pipeline stage
constexpr int MMA_K = 32;          // SNPs per tcgen05.mma (int8)
constexpr int NSTAGE = 3;
constexpr int MAXP = 2;            // passes per launch = TMEM accumulators
constexpr int PROD_WARPS = 8;
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int FIRST_PROD_WARP = 2;                       // warp 0 = MMA issuer, warp 1 = TMA loader
constexpr int THREADS = 32 * (FIRST_PROD_WARP + PROD_WARPS);
constexpr int A_BYTES = TM * SK;   // one pass, one stage (16 KB)
constexpr int B_BYTES = TN * SK;   // one stage (32 KB)
constexpr int STAGE_BYTES = MAXP * A_BYTES + B_BYTES;
// ring of packed (2-bit) genotype boxes + digit tables, filled by TMA PF_DEPTH stages ahead:
// six boxes of 16 bytes x SK rows (A quads 0-1, B quads 0-3), then MAXP tables of SK words
constexpr int PF_DEPTH = 2;
constexpr int PF_BOX = SK * 16;
constexpr int PF_NBOX = (TM + TN) / 64;
constexpr int PF_TAB = SK * 4;
constexpr int PF_BYTES = PF_NBOX * PF_BOX + MAXP * PF_TAB;
constexpr int BAR_OFFSET = NSTAGE * STAGE_BYTES + PF_DEPTH * PF_BYTES;
constexpr int SMEM_BYTES = BAR_OFFSET + 1024;
constexpr int A_LBO = (TM / 16) * 128;   // byte stride between 8-SNP groups (K direction)
constexpr int B_LBO = (TN / 16) * 128;
constexpr int CORE_SBO = 128;            // byte stride between 16-sample cores (MN direction)
constexpr uint32_t TMEM_COLS = 512;

struct Params {
    const uint8_t *geno;
    long long row_bytes;
    const uint32_t *tabA[MAXP];
    uint32_t tabB;
    int npass;
    int plane[MAXP];
    int shift[MAXP];
    long long *out;
    long long ld;            // leading dimension of an output plane (n_samp_pad)
    long long plane_stride;  // elements per plane
    long long n_samp;
    const int2 *tiles;       // (tile_m, tile_n) work list
    int stages_total;
    int stages_per_split;
    int upper_only;
    uint32_t flags;          // bit0: swap LBO/SBO in the descriptors (bring-up probe)
    uint32_t sh32;           // always 32 (see mbar_arrive_after)
    int *error_flag;
};

// instruction descriptor for kind::i8: D=s32, A=s8, B=s8, both MN-major, M=128, N=256
__device__ __forceinline__ uint32_t make_idesc() {
    return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(TN >> 3) << 17) |
           ((uint32_t)(TM >> 4) << 24);
}

template <int NP>
__global__ void __launch_bounds__(THREADS, 1)
table_gram_kernel(const __grid_constant__ Params P, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + BAR_OFFSET;
    // barriers: full[NSTAGE], empty[NSTAGE], accum, pf_full[PF_DEPTH], pf_empty[PF_DEPTH]; then the TMEM slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGE + s); };
    const uint32_t accum_bar = bar_base + 8u * (2 * NSTAGE);
    auto pf_full = [&](int s) { return bar_base + 8u * (2 * NSTAGE + 1 + s); };
    auto pf_empty = [&](int s) { return bar_base + 8u * (2 * NSTAGE + 1 + PF_DEPTH + s); };
    constexpr int SLOT_IDX = 2 * NSTAGE + 1 + 2 * PF_DEPTH;
    const uint32_t tmem_slot = bar_base + 8u * SLOT_IDX;
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem + BAR_OFFSET + 8 * SLOT_IDX);
    const uint32_t pf_base = smem_base + NSTAGE * STAGE_BYTES;

    const int2 tile = P.tiles[blockIdx.x];
    const int st_begin = blockIdx.y * P.stages_per_split;
    const int st_end = min(P.stages_total, st_begin + P.stages_per_split);
    const int nst = st_end - st_begin;
    if (nst <= 0) return;

    if (warp == 0) {
        if (lane == 0) {
            for (int s = 0; s < NSTAGE; s++) {
                mbar_init(full_bar(s), PROD_THREADS);
                mbar_init(empty_bar(s), 1);
            }
            mbar_init(accum_bar, 1);
            for (int s = 0; s < PF_DEPTH; s++) {
                mbar_init(pf_full(s), 1);
                mbar_init(pf_empty(s), PROD_THREADS);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc(tmem_slot, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc();
        const bool swap = (P.flags & 1u) != 0;
        const uint32_t a_lbo = swap ? CORE_SBO : A_LBO, a_sbo = swap ? A_LBO : CORE_SBO;
        const uint32_t b_lbo = swap ? CORE_SBO : B_LBO, b_sbo = swap ? B_LBO : CORE_SBO;
        for (int it = 0; it < nst; it++) {
            const int s = it % NSTAGE;
            const uint32_t phase = (uint32_t)(it / NSTAGE) & 1u;
            mbar_wait(full_bar(s), phase, P.error_flag, 1);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t stage_addr = smem_base + s * STAGE_BYTES;
#pragma unroll
                for (int j = 0; j < SK / MMA_K; j++) {
                    // K step j covers 4 groups of 8 SNPs
                    uint64_t bdesc = make_desc(stage_addr + MAXP * A_BYTES + j * 4 * B_LBO, b_lbo, b_sbo);
#pragma unroll
                    for (int p = 0; p < NP; p++) {
                        uint64_t adesc =
                            make_desc(stage_addr + p * A_BYTES + j * 4 * A_LBO, a_lbo, a_sbo);
                        umma_i8(tmem_base + (uint32_t)(p * TN), adesc, bdesc, idesc,
                                (it > 0 || j > 0) ? 1u : 0u);
                    }
                }
                umma_commit(empty_bar(s));   // frees the stage once these MMAs have read it
            }
            __syncwarp();
        }
        if (lane == 0) umma_commit(accum_bar);
        __syncwarp();
    } else if (warp == 1) {
        // ===================== TMA loader =====================
        if (lane == 0) {
            const int ax = tile.x * (TM / 4), bx = tile.y * (TN / 4);
            for (int it = 0; it < nst; it++) {
                const int sl = it % PF_DEPTH;
                const uint32_t ph = (uint32_t)(it / PF_DEPTH) & 1u;
                mbar_wait(pf_empty(sl), ph ^ 1u, P.error_flag, 4);
                const uint32_t slot = pf_base + (uint32_t)sl * PF_BYTES;
                const uint32_t bar = pf_full(sl);
                mbar_arrive_expect_tx(bar, PF_NBOX * PF_BOX + NP * PF_TAB);
                const int y = (st_begin + it) * SK;
#pragma unroll
                for (int q = 0; q < TM / 64; q++) tma_load_2d(slot + q * PF_BOX, &tmap, ax + q * 16, y, bar);
#pragma unroll
                for (int q = 0; q < TN / 64; q++)
                    tma_load_2d(slot + (TM / 64 + q) * PF_BOX, &tmap, bx + q * 16, y, bar);
#pragma unroll
                for (int q = 0; q < NP; q++)
                    bulk_load(slot + PF_NBOX * PF_BOX + q * PF_TAB, P.tabA[q] + (long long)y, PF_TAB, bar);
            }
        }
        __syncwarp();
    } else {
        // ===================== producers =====================
        const int p = threadIdx.x - 32 * FIRST_PROD_WARP;
        const int sl = p & (SK - 1);   // SNP within the stage
        const int half = p >> 7;       // which 64-sample quad of A / which 128-sample half of B
        const int kg = sl >> 3, r = sl & 7;
        uint32_t tb[1] = {P.tabB};
        const uint32_t a_off = kg * A_LBO + (half * 4) * CORE_SBO + r * 16;
        const uint32_t b_off = MAXP * A_BYTES + kg * B_LBO + (half * 8) * CORE_SBO + r * 16;
        // this thread's words inside a ring slot (16-byte box rows: conflict-free LDS.128)
        const uint32_t pa = half * PF_BOX + sl * 16;
        const uint32_t pb = (TM / 64 + 2 * half) * PF_BOX + sl * 16;
        const uint32_t pt = PF_NBOX * PF_BOX + sl * 4;

#pragma unroll 1
        for (int it = 0; it < nst; it++) {
            const int s = it % NSTAGE;
            const uint32_t phase = (uint32_t)(it / NSTAGE) & 1u;
            const int ps = it % PF_DEPTH;
            mbar_wait(pf_full(ps), (uint32_t)(it / PF_DEPTH) & 1u, P.error_flag, 5);
            const uint32_t slot = pf_base + (uint32_t)ps * PF_BYTES;
            const uint4 ca = ld_shared_v4(slot + pa);
            const uint4 cb0 = ld_shared_v4(slot + pb);
            const uint4 cb1 = ld_shared_v4(slot + pb + PF_BOX);
            uint32_t ct[NP];
#pragma unroll
            for (int q = 0; q < NP; q++) ct[q] = ld_shared_u32(slot + pt + q * PF_TAB);
            // the loader may refill this slot -- but only once every word above has really been read
            mbar_arrive_after(pf_empty(ps), ca.x ^ ca.y ^ ca.z ^ ca.w ^ cb0.x ^ cb0.y ^ cb0.z ^ cb0.w ^ cb1.x ^ cb1.y ^
                                                cb1.z ^ cb1.w ^ ct[0] ^ ct[NP - 1],
                              P.sh32);

            mbar_wait(empty_bar(s), phase ^ 1u, P.error_flag, 2);
            const uint32_t stage_addr = smem_base + s * STAGE_BYTES;
            const uint32_t aw[4] = {ca.x, ca.y, ca.z, ca.w};
#pragma unroll
            for (int w = 0; w < 4; w++) {
                uint32_t dst[NP];
#pragma unroll
                for (int q = 0; q < NP; q++) dst[q] = stage_addr + q * A_BYTES + a_off + w * CORE_SBO;
                expand_word<NP>(aw[w], ct, dst);
            }
            const uint32_t bw[8] = {cb0.x, cb0.y, cb0.z, cb0.w, cb1.x, cb1.y, cb1.z, cb1.w};
#pragma unroll
            for (int w = 0; w < 8; w++) {
                uint32_t dst[1] = {stage_addr + b_off + w * CORE_SBO};
                expand_word<1>(bw[w], tb, dst);
            }
            fence_proxy_async_smem();   // generic-proxy stores -> visible to the tensor core
            mbar_arrive(full_bar(s));
        }

        // ===================== epilogue =====================
        mbar_wait(accum_bar, 0, P.error_flag, 3);
        tc_fence_after();
        const int quarter = warp & 3;            // TMEM lanes this warp may touch
        const int colhalf = (warp - FIRST_PROD_WARP) >> 2;     // two warps share a lane quarter
        const int row = quarter * 32 + lane;
        const long long gi = (long long)tile.x * TM + (row & ~15) + core_pos_to_sample(row & 15);
#pragma unroll 1
        for (int q = 0; q < NP; q++) {
            long long *outp = P.out + (long long)P.plane[q] * P.plane_stride + gi * P.ld;
            const long long mul = 1ll << P.shift[q];
#pragma unroll 1
            for (int cc = 0; cc < 4; cc++) {
                const int col0 = colhalf * 128 + cc * 32;
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(q * TN + col0), v);
                if (gi < P.n_samp) {
#pragma unroll
                    for (int k = 0; k < 32; k++) {
                        int col = col0 + k;
                        long long gj = (long long)tile.y * TN + (col & ~15) + core_pos_to_sample(col & 15);
                        int val = (int)v[k];
                        if (val != 0 && gj < P.n_samp && (!P.upper_only || gj >= gi))
                            atomicAdd(reinterpret_cast<unsigned long long *>(outp + gj),
                                      (unsigned long long)((long long)val * mul));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace tc

// ---------------------------------------------------------------------------
// host driver: group passes that share a B table into launches of <= 2
// ---------------------------------------------------------------------------
void gram_tc_run(snprel_ctx *c, const GramPass *passes, int npass, long long *out_planes,
                 bool upper_only) {
    using namespace tc;
    if (npass <= 0) return;
    if (!(c->debug_flags & 4u)) {   // default: CTA-pair (cta_group::2) kernel, gram_tc2.cu
        gram_tc2_run(c, passes, npass, out_planes, upper_only);
        return;
    }
    geno_pad_tail(c);
    const int64_t n = c->n_samp, npad = c->n_samp_pad;
    const int tiles_m = (int)((n + TM - 1) / TM), tiles_n = (int)((n + TN - 1) / TN);
    std::vector<int2> &tiles = c->host_tiles;
    tiles.clear();
    for (int tm = 0; tm < tiles_m; tm++)
        for (int tn = 0; tn < tiles_n; tn++)
            if (!upper_only || (int64_t)tn * TN + TN - 1 >= (int64_t)tm * TM) tiles.push_back(make_int2(tm, tn));
    DevBuf<int2> &dtiles = c->scr_tiles;
    dtiles.alloc(tiles.size());
    CUDA_CHECK(cudaMemcpyAsync(dtiles.p, tiles.data(), tiles.size() * sizeof(int2),
                               cudaMemcpyHostToDevice, c->stream));
    c->scr_flags.alloc(2);
    int *derr = c->scr_flags.p + 1;
    CUDA_CHECK(cudaMemsetAsync(derr, 0, sizeof(int), c->stream));

    const int stages_total = (int)(round_up(std::max<int64_t>(c->n_snp, 1), SK) / SK);
    // split the SNP range when the tile list alone cannot fill the SMs
    int64_t want = (int64_t)c->num_sms * 2;
    int64_t splits = std::max<int64_t>(1, std::min<int64_t>((want + (int64_t)tiles.size() - 1) /
                                                               (int64_t)tiles.size(),
                                                           stages_total));
    if (c->debug_flags & 2u) splits = 1;   // test hook: one CTA walks the whole SNP range
    splits = std::min<int64_t>(splits, 65535);
    int sps = (int)((stages_total + splits - 1) / splits);
    splits = (stages_total + sps - 1) / sps;
    // int32 accumulator headroom: |digit| <= 128, |tabB| <= 2  ->  256 per SNP
    const int max_stages_i32 = (int)((2147483647ll / 256) / SK);
    if (sps > max_stages_i32) {
        sps = max_stages_i32;
        splits = (stages_total + sps - 1) / sps;
    }

    // tensor map over the packed genotype matrix: bytes x SNP rows, box = 16 bytes x SK rows
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) fail("cuTensorMapEncodeTiled is not available in this driver");
        encode = reinterpret_cast<EncodeFn>(fn);
    }
    alignas(64) CUtensorMap tmap;
    {
        const int64_t rows = round_up(std::max<int64_t>(c->n_snp, 1), SK);
        cuuint64_t gdim[2] = {(cuuint64_t)c->row_bytes, (cuuint64_t)rows};
        cuuint64_t gstride[1] = {(cuuint64_t)c->row_bytes};
        cuuint32_t box[2] = {16, (cuuint32_t)SK};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, c->geno2b.p, gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) fail("cuTensorMapEncodeTiled failed (%d)", (int)r);
    }

    static bool attr_done = false;
    if (!attr_done) {
        CUDA_CHECK(cudaFuncSetAttribute(table_gram_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        SMEM_BYTES));
        CUDA_CHECK(cudaFuncSetAttribute(table_gram_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        SMEM_BYTES));
        attr_done = true;
    }

    std::vector<char> used((size_t)npass, 0);
    for (int i = 0; i < npass; i++) {
        if (used[i]) continue;
        Params P{};
        P.geno = c->geno2b.p;
        P.row_bytes = c->row_bytes;
        P.tabB = passes[i].tabB;
        P.out = out_planes;
        P.ld = npad;
        P.plane_stride = npad * npad;
        P.n_samp = n;
        P.tiles = dtiles.p;
        P.stages_total = stages_total;
        P.stages_per_split = sps;
        P.upper_only = upper_only ? 1 : 0;
        P.flags = c->debug_flags;
        P.sh32 = 32;
        P.error_flag = derr;
        int np = 0;
        for (int j = i; j < npass && np < MAXP; j++) {
            if (used[j] || passes[j].tabB != passes[i].tabB) continue;
            P.tabA[np] = passes[j].tabA;
            P.plane[np] = passes[j].plane;
            P.shift[np] = passes[j].shift;
            used[j] = 1;
            np++;
        }
        P.npass = np;
        dim3 grid((unsigned)tiles.size(), (unsigned)splits);
        if (np == 2)
            table_gram_kernel<2><<<grid, THREADS, SMEM_BYTES, c->stream>>>(P, tmap);
        else
            table_gram_kernel<1><<<grid, THREADS, SMEM_BYTES, c->stream>>>(P, tmap);
        KERNEL_CHECK(c);
        c->hot_launches++;
    }
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    int herr = 0;
    CUDA_CHECK(cudaMemcpy(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost));
    if (herr) fail("table_gram_kernel: pipeline barrier %d timed out", herr);
}

}  // namespace snprel
