// K1: the tcgen05 "table Gram" -- the tensor-core hot loop that replaces
// CProdMat_AlgArith::MulAdd (src/genPCA.cpp:229-312) together with its block preparation
// (TransposeGenotype / GenoSub / GenoMul, src/genPCA.h:93-108, src/genPCA.cpp:315-368) and the
// missing-pair denominator loops (src/genPCA.cpp:1201-1224, src/genEIGMIX.cpp:113-138).
//
// For up to two "passes" p that share one B table it computes
//
//     Acc_p[i][j] = sum over SNPs l of  tabA_p[l][g_il] * tabB[g_jl]      (exact, int32)
//
// where g is the 2-bit genotype code, tabA_p[l] a per-SNP table of four int8 digits and tabB a
// table of four small int8 values, and adds Acc_p << shift_p into an int64 fixed-point plane
// with 64-bit atomics.  The per-SNP real weights (1/(p(1-p)), 2p, 4p(1-p) ...) live in the digits
// of tabA: grm.cu slices each table value into balanced base-256 digits, one pass per digit, so
// the sum over passes reproduces the float64 result to a proven bound while every tensor-core
// product and every accumulation is exact integer arithmetic (order independent, hence
// bit-identical for any tiling, split or GPU count).
//
// Mapping onto sm_100a: a cluster of two CTAs (one TPC, tcgen05 cta_group::2) owns a 256 x 256
// sample tile (upper triangle only) and an SNP split; 10 warps per CTA, 1 CTA per SM.
//   * warp 1, one lane -- TMA loader: per 128-SNP stage two cp.async.bulk.tensor.2d boxes
//     (32 bytes = 128 samples x 128 SNP rows, 32-byte swizzle) of the packed 2-bit genotypes plus
//     cp.async.bulk copies of the digit tables into a 3-deep shared-memory ring, mbarrier
//     complete_tx.  (Round 1 used four 16-byte boxes: every 32-byte L2 sector was fetched twice;
//     the wider box cut the L2->SM sectors by 40 % and the kernel time by 3.7 %, profiles/r02_notes.md.)
//   * warps 2..9 -- producers: each thread takes 64 packed genotypes of one SNP for this CTA's
//     128 A rows and for ITS HALF (128) of the B rows, turns every 32-bit word (16 samples) into
//     byte-permute selectors (3 logic ops + 2 shifts) and emits int8 operand rows with PRMT
//     against the 4-entry tables -- the fused "unpack, centre and scale" step.  Operands are
//     written MN-major (16 consecutive samples = one 16-byte core-matrix row), the layout in
//     which a packed SNP-major genotype word expands without any transpose.
//   * warp 0 of CTA 0, one lane -- MMA issuer: tcgen05.mma.cta_group::2.kind::i8 (M=256, N=256,
//     K=32) reads A and the B halves from both SMs' shared memory and accumulates 128 x 256 int32
//     per pass into each SM's TMEM (2 passes = all 512 columns); tcgen05.commit multicasts the
//     stage-empty / accumulator-ready arrivals to both CTAs.
//   * the stage-full barrier lives in CTA 0 and counts the producers of BOTH CTAs; CTA 1's
//     producers arrive on it remotely (mapa + mbarrier.arrive on the cluster address).
//   * after the last stage the producer warps become the epilogue: tcgen05.ld their 128 rows and
//     issue the 64-bit atomics.
// History (profiles/): a single-CTA 128 x 256 version was bounded by shared-memory traffic
// (64 KB written + 96 KB read by the tensor core per stage); the pair halves the B expansion and
// the B traffic per CTA (48 KB + 64 KB).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace snprel {
namespace tc2 {

using namespace tc;

constexpr int TM2 = 256;           // tile rows (A), 128 per CTA
constexpr int TN2 = 256;           // tile cols (B), 128 per CTA
constexpr int HM = 128, HN = 128;  // per-CTA halves
constexpr int SK = 128;            // SNPs per stage
constexpr int MMA_K = 32;
constexpr int NSTAGE = 4;
constexpr int PROD_WARPS = 8;
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int FIRST_PROD_WARP = 2;
constexpr int THREADS = 32 * (FIRST_PROD_WARP + PROD_WARPS);
constexpr int A_BYTES = HM * SK;   // 16 KB per pass per stage
constexpr int B_BYTES = HN * SK;   // 16 KB per stage
// An item fills the two TMEM accumulators either with two passes over one B tile (NP=2, NB=1)
// or with ONE pass over two B tiles (NP=1, NB=2: a 256 x 512 tile, so that the odd pass of a
// table costs half an item instead of ~0.8).  Either way a stage holds three 16 KB operands.
constexpr int STAGE_BYTES = 3 * A_BYTES;   // 48 KB
constexpr int PF_BOX = SK * 16;
constexpr int PF_TAB = SK * 4;
constexpr int PF_MAX_DEPTH = 3;
constexpr int PF_RING_BYTES = 29184;       // 3 x (4 boxes + 2 A tables + 1 B table) or 2 x (6 boxes + 1 + 1)
constexpr int BAR_OFFSET = NSTAGE * STAGE_BYTES + PF_RING_BYTES;
constexpr int SMEM_BYTES = BAR_OFFSET + 1024;
constexpr int LBO = (HM / 16) * 128;   // 8 cores per 8-SNP group (A and B halves alike)
constexpr int SBO = 128;
constexpr uint32_t TMEM_COLS = 512;

// one digit pass: per-SNP A table (4 int8 digits by genotype code) against a per-SNP B table
struct PassDesc {
    const uint32_t *tabA;
    const uint32_t *tabB;
    int plane;   // output plane
    int shift;   // contribution = acc << shift
};

// One work item = one CTA pair: a 256-row tile x one or two 256-column B tiles x an SNP range.
// All items of a step go out in ONE launch (heterogeneous grid); the hardware block scheduler
// deals them to the 74 CTA-pair slots in list order, so the host orders them for L2 locality
// (tile-major, super-tile raster) and puts the small ones last (wave tail).
struct Item {
    int tm;              // A tile row, units of 256 samples
    int tn;              // first B tile column, units of 256 samples
    int st_begin, st_end;
    short ncols[2];      // MMA N of each B tile: 0 (absent / below the diagonal), 128 or 256
    short mode;          // 0: NP=2 passes x NB=1 B tile; 1: NP=1 pass x NB=2 B tiles
    short pad;
    int pass[2];
};

struct Params {
    const Item *items;
    const PassDesc *passes;
    long long *out;
    long long ld;
    long long plane_stride;
    long long n_samp;
    long long row0;          // first row of the output window (planes hold rows row0 ..)
    int upper_only;
    uint32_t sh32;
    int box32;               // genotype boxes are 32 bytes wide with the 32-byte TMA swizzle (else 16 bytes, plain)
    int *error_flag;
    long long *trace;        // optional [items][8] clock64 stamps of the leader CTA (snprel_debug_flags 1, tools/k1_trace.py)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default .release.cta semantics as in CUTLASS' ClusterBarrier::arrive(cta_id): an explicit
    // .release.cluster costs a MEMBAR.ALL.GPU + ERRBAR per thread and stage (33 % of all stall
    // samples in profiles/r01 notes); the data being published is this CTA's own shared memory,
    // already ordered for the async proxy by fence.proxy.async
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        " selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int *error_flag, int code) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) {
            if (error_flag) atomicExch(error_flag, code);
            __trap();
        }
    }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {
    asm volatile(
        "{\n .reg .b16 m;\n mov.b16 m, 3;\n"
        " tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n}"
        ::"r"(bar)
        : "memory");
}
// D=s32, A=s8, B=s8, MN-major, M=256 (pair), N = ncols (32 .. 256)
__device__ __forceinline__ uint32_t make_idesc2(int ncols) {
    return (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(ncols >> 3) << 17) |
           ((uint32_t)(TM2 >> 4) << 24);
}

template <int NP, int NB, bool TRACE>
__device__ __forceinline__ void gram_item(const Params &P, const Item &item, const CUtensorMap *tmap, uint8_t *smem) {
    static_assert(NP * NB == 2, "two TMEM accumulators of 256 columns");
    constexpr int PF_DEPTH = NB == 1 ? 3 : 2;
    constexpr int PF_NBOX = 2 + 2 * NB;                       // A quads 0-1, then two quads per B tile
    constexpr int PF_BYTES = PF_NBOX * PF_BOX + (NP + 1) * PF_TAB;   // boxes, NP A tables, the B table
    static_assert(PF_DEPTH * PF_BYTES <= PF_RING_BYTES, "ring does not fit");
    static_assert(PF_BYTES % 256 == 0, "TMA destinations are 128-byte aligned; the 32-byte swizzle pattern repeats every 256");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + BAR_OFFSET;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (NSTAGE + s); };
    const uint32_t accum_bar = bar_base + 8u * (2 * NSTAGE);
    auto pf_full = [&](int s) { return bar_base + 8u * (2 * NSTAGE + 1 + s); };
    auto pf_empty = [&](int s) { return bar_base + 8u * (2 * NSTAGE + 1 + PF_MAX_DEPTH + s); };
    constexpr int SLOT_IDX = 2 * NSTAGE + 1 + 2 * PF_MAX_DEPTH;
    const uint32_t tmem_slot = bar_base + 8u * SLOT_IDX;
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem + BAR_OFFSET + 8 * SLOT_IDX);
    const uint32_t pf_base = smem_base + NSTAGE * STAGE_BYTES;

    const int st_begin = item.st_begin;
    const int nst = item.st_end - item.st_begin;   // identical in both CTAs of the pair
    if (nst <= 0) return;
    // (clock stamps exist in the TRACE instantiation only: the shipped path carries no extra code)
    long long *tr = (TRACE && P.trace && cluster_ctarank() == 0) ? P.trace + (long long)(blockIdx.x >> 1) * 8 : nullptr;
    if (TRACE && tr && threadIdx.x == 0) {
        tr[0] = clock64();
        tr[7] = (long long)nst | ((long long)item.mode << 32) | ((long long)(item.ncols[0] + item.ncols[1]) << 40);
    }
    int ncols[NB];
#pragma unroll
    for (int b = 0; b < NB; b++) ncols[b] = item.ncols[b];
    PassDesc pd[NP];
#pragma unroll
    for (int q = 0; q < NP; q++) pd[q] = P.passes[item.pass[q]];

    if (warp == 0) {
        if (lane == 0) {
            for (int s = 0; s < NSTAGE; s++) {
                mbar_init(full_bar(s), 2 * PROD_THREADS);   // producers of both CTAs (used in CTA 0 only)
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
        tmem_alloc2(tmem_slot, TMEM_COLS);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();          // peer barriers are initialised before anyone arrives remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    if (TRACE && tr && threadIdx.x == 0) tr[1] = clock64();

    if (warp == 0) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0) {
            uint32_t idesc[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) idesc[b] = make_idesc2(ncols[b]);
            for (int it = 0; it < nst; it++) {
                const int s = it % NSTAGE;
                const uint32_t phase = (uint32_t)(it / NSTAGE) & 1u;
                mbar_wait_cluster(full_bar(s), phase, P.error_flag, 1);
                tc_fence_after();
                if (TRACE && tr && it == 0 && lane == 0) tr[2] = clock64();
                if (lane == 0) {
                    const uint32_t stage_addr = smem_base + s * STAGE_BYTES;
#pragma unroll
                    for (int j = 0; j < SK / MMA_K; j++) {
#pragma unroll
                        for (int b = 0; b < NB; b++) {
                            if (ncols[b] == 0) continue;
                            uint64_t bdesc = make_desc(stage_addr + (NP + b) * A_BYTES + j * 4 * LBO, LBO, SBO);
#pragma unroll
                            for (int p = 0; p < NP; p++) {
                                uint64_t adesc = make_desc(stage_addr + p * A_BYTES + j * 4 * LBO, LBO, SBO);
                                umma2_i8(tmem_base + (uint32_t)((p * NB + b) * TN2), adesc, bdesc, idesc[b],
                                         (it > 0 || j > 0) ? 1u : 0u);
                            }
                        }
                    }
                    umma2_commit_mc(empty_bar(s));
                }
                __syncwarp();
            }
            if (lane == 0) umma2_commit_mc(accum_bar);
            if (TRACE && tr && lane == 0) tr[3] = clock64();
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== TMA loader =====================
        if (lane == 0) {
            const int ax = (item.tm * TM2 + (int)rank * HM) / 4;
            uint32_t tx_bytes = 2 * PF_BOX + (NP + 1) * PF_TAB;
#pragma unroll
            for (int b = 0; b < NB; b++) tx_bytes += ncols[b] ? 2 * PF_BOX : 0;
            for (int it = 0; it < nst; it++) {
                const int sl = it % PF_DEPTH;
                const uint32_t ph = (uint32_t)(it / PF_DEPTH) & 1u;
                mbar_wait(pf_empty(sl), ph ^ 1u, P.error_flag, 4);
                const uint32_t slot = pf_base + (uint32_t)sl * PF_BYTES;
                const uint32_t bar = pf_full(sl);
                mbar_arrive_expect_tx(bar, tx_bytes);
                const int y = (st_begin + it) * SK;
                if (P.box32) {
                    tma_load_2d(slot, tmap, ax, y, bar);
                } else {
#pragma unroll
                    for (int q = 0; q < HM / 64; q++) tma_load_2d(slot + q * PF_BOX, tmap, ax + q * 16, y, bar);
                }
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    if (ncols[b] == 0) continue;
                    // this CTA's half of the B tile's ncols columns starts at sample rank * ncols / 2 of the
                    // tile, wherever that falls inside a 16-byte group: box coordinates are free.  Columns
                    // beyond the padded matrix are out of bounds for the tensor map and arrive as zeros
                    // (code 0); the epilogue drops them.
                    const int bx = ((item.tn + b) * TN2 + (int)rank * (ncols[b] >> 1)) / 4;
                    if (P.box32) {
                        tma_load_2d(slot + (2 + 2 * b) * PF_BOX, tmap, bx, y, bar);
                    } else {
#pragma unroll
                        for (int q = 0; q < HN / 64; q++)
                            tma_load_2d(slot + (2 + 2 * b + q) * PF_BOX, tmap, bx + q * 16, y, bar);
                    }
                }
#pragma unroll
                for (int q = 0; q < NP; q++)
                    bulk_load(slot + PF_NBOX * PF_BOX + q * PF_TAB, pd[q].tabA + (long long)y, PF_TAB, bar);
                bulk_load(slot + PF_NBOX * PF_BOX + NP * PF_TAB, pd[0].tabB + (long long)y, PF_TAB, bar);
            }
        }
        __syncwarp();
    } else {
        // ===================== producers =====================
        const int p = threadIdx.x - 32 * FIRST_PROD_WARP;
        const int sl = p & (SK - 1);   // SNP within the stage
        const int half = p >> 7;       // which 64-sample quad of this CTA's A half and of its B half
        const int kg = sl >> 3, r = sl & 7;
        const uint32_t a_off = kg * LBO + (half * 4) * SBO + r * 16;
        const uint32_t b_off = NP * A_BYTES + a_off;
        // 16-byte boxes: [quad][SNP][16 B].  32-byte boxes: [SNP][32 B] with the 32-byte swizzle (address
        // bit 4 ^= bit 7; the slots are 256-byte aligned), which keeps the LDS.128 of eight consecutive
        // SNPs on eight different 16-byte bank groups.
        const uint32_t sw = (uint32_t)(half ^ ((sl >> 2) & 1)) * 16u;
        const uint32_t pa = P.box32 ? (uint32_t)sl * 32u + sw : half * PF_BOX + sl * 16;
        const uint32_t pb = P.box32 ? 2 * PF_BOX + (uint32_t)sl * 32u + sw : (2 + half) * PF_BOX + sl * 16;
        const uint32_t pt = PF_NBOX * PF_BOX + sl * 4;
        int bwords[NB];                // 16-sample words of B tile b this thread expands (0..4)
#pragma unroll
        for (int b = 0; b < NB; b++) bwords[b] = min(4, max(0, (ncols[b] >> 5) - 4 * half));
        uint32_t full_remote[NSTAGE];
#pragma unroll
        for (int s = 0; s < NSTAGE; s++) full_remote[s] = mapa(full_bar(s), 0);   // leader's barrier

#pragma unroll 1
        for (int it = 0; it < nst; it++) {
            const int s = it % NSTAGE;
            const uint32_t phase = (uint32_t)(it / NSTAGE) & 1u;
            const int ps = it % PF_DEPTH;
            mbar_wait(pf_full(ps), (uint32_t)(it / PF_DEPTH) & 1u, P.error_flag, 5);
            const uint32_t slot = pf_base + (uint32_t)ps * PF_BYTES;
            const uint4 ca = ld_shared_v4(slot + pa);
            uint4 cb[NB];
#pragma unroll
            for (int b = 0; b < NB; b++) cb[b] = bwords[b] ? ld_shared_v4(slot + pb + 2 * b * PF_BOX) : make_uint4(0, 0, 0, 0);
            uint32_t ct[NP];
#pragma unroll
            for (int q = 0; q < NP; q++) ct[q] = ld_shared_u32(slot + pt + q * PF_TAB);
            uint32_t tb[1] = {ld_shared_u32(slot + pt + NP * PF_TAB)};
            uint32_t dep = ca.x ^ ca.y ^ ca.z ^ ca.w ^ ct[0] ^ ct[NP - 1] ^ tb[0];
#pragma unroll
            for (int b = 0; b < NB; b++) dep ^= cb[b].x ^ cb[b].y ^ cb[b].z ^ cb[b].w;
            mbar_arrive_after(pf_empty(ps), dep, P.sh32);

            mbar_wait(empty_bar(s), phase ^ 1u, P.error_flag, 2);
            const uint32_t stage_addr = smem_base + s * STAGE_BYTES;
            const uint32_t aw[4] = {ca.x, ca.y, ca.z, ca.w};
#pragma unroll
            for (int w = 0; w < 4; w++) {
                uint32_t dst[NP];
#pragma unroll
                for (int q = 0; q < NP; q++) dst[q] = stage_addr + q * A_BYTES + a_off + w * SBO;
                expand_word<NP>(aw[w], ct, dst);
            }
#pragma unroll
            for (int b = 0; b < NB; b++) {
                const uint32_t bw[4] = {cb[b].x, cb[b].y, cb[b].z, cb[b].w};
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    if (w < bwords[b]) {
                        uint32_t dst[1] = {stage_addr + b_off + b * B_BYTES + w * SBO};
                        expand_word<1>(bw[w], tb, dst);
                    }
                }
            }
            fence_proxy_async_smem();
            // constant-index select keeps full_remote[] in registers
            uint32_t fr = full_remote[0];
#pragma unroll
            for (int k = 1; k < NSTAGE; k++) fr = (s == k) ? full_remote[k] : fr;
            mbar_arrive_cluster(fr);
        }

        // ===================== epilogue (each CTA drains its own 128 rows) =====================
        mbar_wait(accum_bar, 0, P.error_flag, 3);
        tc_fence_after();
        if (TRACE && tr && threadIdx.x == 32 * FIRST_PROD_WARP) tr[4] = clock64();
        const int quarter = warp & 3;
        const int colhalf = (warp - FIRST_PROD_WARP) >> 2;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // A TMEM lane is a ROW: straight out of tcgen05.ld a warp's 32 lanes would hit 32 different rows of the
        // plane with every reduction (32 sectors per instruction, 65 536 sector reductions per CTA and item: the
        // 40 us per item that kept the tensor pipe at 90 %).  The block is transposed through the operand
        // stages (free once the last MMA has been committed), so that a warp's reduction covers 32
        // consecutive columns of ONE row.
        long long *stage = reinterpret_cast<long long *>(smem + (size_t)(warp - FIRST_PROD_WARP) * (32 * 33 * 8));
        const long long row_base = (long long)item.tm * TM2 + (long long)rank * HM + quarter * 32;
        auto flush = [&](const long long (&val)[32], long long *plane, long long tn_eff, int col0) {
#pragma unroll
            for (int k = 0; k < 32; k++) stage[lane * 33 + k] = val[k];
            __syncwarp();
            const int col = col0 + lane;
            const long long gj = tn_eff * TN2 + (col & ~15) + core_pos_to_sample(col & 15);
            const bool col_ok = gj < P.n_samp;
#pragma unroll 4
            for (int r = 0; r < 32; r++) {
                const long long v = stage[r * 33 + lane];
                const long long gr = row_base + (r & ~15) + core_pos_to_sample(r & 15);
                if (v != 0 && col_ok && gr < P.n_samp && (!P.upper_only || gj >= gr))
                    atomicAdd(reinterpret_cast<unsigned long long *>(plane + (gr - P.row0) * P.ld + gj), (unsigned long long)v);
            }
            __syncwarp();
        };
        // two digit passes that feed the same plane leave as ONE 64-bit atomic per entry
        const bool combine = NP == 2 && pd[0].plane == pd[NP - 1].plane;
        if (combine) {
            long long *plane = P.out + (long long)pd[0].plane * P.plane_stride;
            const long long mul0 = 1ll << pd[0].shift, mul1 = 1ll << pd[NP - 1].shift;
#pragma unroll 1
            for (int cc = 0; cc < 4; cc++) {
                const int col0 = (2 * cc + colhalf) * 32;
                if (col0 >= ncols[0]) break;
                uint32_t v0[32], v1[32];
                tmem_ld32(lane_base + (uint32_t)col0, v0);
                tmem_ld32(lane_base + (uint32_t)(TN2 + col0), v1);
                long long val[32];
#pragma unroll
                for (int k = 0; k < 32; k++) val[k] = (long long)(int)v0[k] * mul0 + (long long)(int)v1[k] * mul1;
                flush(val, plane, (long long)item.tn, col0);
            }
        } else {
#pragma unroll 1
            for (int a = 0; a < 2; a++) {        // accumulator a = pass (a / NB), B tile (a % NB)
                const int q = a / NB, bt = a % NB;
                long long *plane = P.out + (long long)pd[q].plane * P.plane_stride;
                const long long mul = 1ll << pd[q].shift;
#pragma unroll 1
                for (int cc = 0; cc < 4; cc++) {
                    const int col0 = (2 * cc + colhalf) * 32;
                    if (col0 >= ncols[bt]) break;
                    uint32_t v[32];
                    tmem_ld32(lane_base + (uint32_t)(a * TN2 + col0), v);
                    long long val[32];
#pragma unroll
                    for (int k = 0; k < 32; k++) val[k] = (long long)(int)v[k] * mul;
                    flush(val, plane, (long long)item.tn + bt, col0);
                }
            }
        }
    }
    if (TRACE && tr && threadIdx.x == 32 * FIRST_PROD_WARP) tr[5] = clock64();
    tc_fence_before();
    __syncthreads();
    cluster_sync();          // both CTAs are done with TMEM and with each other's barriers
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, TMEM_COLS);
    }
    if (TRACE && tr && threadIdx.x == 0) tr[6] = clock64();
}

template <bool TRACE>
__device__ __forceinline__ void table_gram_body(const Params &P, const CUtensorMap *tmap) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const Item item = P.items[blockIdx.x >> 1];
    if (item.mode == 0)
        gram_item<2, 1, TRACE>(P, item, tmap, smem);
    else
        gram_item<1, 2, TRACE>(P, item, tmap, smem);
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
table_gram_kernel3(const __grid_constant__ Params P, const __grid_constant__ CUtensorMap tmap) {
    table_gram_body<false>(P, &tmap);
}
// the same kernel with per-item clock stamps (snprel_debug_flags 1, tools/k1_trace.py)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
table_gram_kernel3_traced(const __grid_constant__ Params P, const __grid_constant__ CUtensorMap tmap) {
    table_gram_body<true>(P, &tmap);
}

}  // namespace tc2

// ---------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------
// device array of snp_cap copies of one table word (constant tables: the x / m channels, the
// {-1,0,1} tables of the tensor count engine), built once per (word, capacity)
__global__ void fill_table_kernel(uint32_t *__restrict__ p, int64_t n, uint32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void pull_staged_kernel(const uint4 *__restrict__ host_mapped, uint4 *__restrict__ dev, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dev[i] = host_mapped[i];
}
const uint32_t *gram_const_table(snprel_ctx *c, uint32_t word) {
    for (auto &t : c->const_tabs)
        if (t->word == word && (int64_t)t->buf.n >= c->snp_cap) return t->buf.p;
    std::unique_ptr<snprel_ctx::ConstTab> t(new snprel_ctx::ConstTab());
    t->word = word;
    t->buf.alloc((size_t)c->snp_cap);
    fill_table_kernel<<<(unsigned)((c->snp_cap + 255) / 256), 256, 0, c->stream>>>(t->buf.p, c->snp_cap, word);
    KERNEL_CHECK(c);
    c->const_tabs.push_back(std::move(t));
    return c->const_tabs.back()->buf.p;
}

// MMA N for a B tile with `nvalid` live columns: 128 or 256.  (The instruction set lists every
// multiple of 32 from 32 for M = 256 / cta_group::2 / kind::i8, but with this MN-major operand
// layout N = 32 and N = 64 raise "illegal instruction" on the B200 -- measured in round 2 -- so the
// ragged last tile column only drops to half width.)
static int tile_ncols(int64_t nvalid) {
    if (nvalid <= 0) return 0;
    return nvalid <= 128 ? 128 : 256;
}

void gram_tc_run(snprel_ctx *c, const GramPass *passes, int npass, long long *out_planes, bool upper_only,
                 int64_t snp_lo, int64_t snp_hi, bool sync) {
    using namespace tc2;
    if (npass <= 0) return;
    geno_pad_tail(c);
    const int64_t n = c->n_samp, npad = c->n_samp_pad;
    const int nt = (int)((n + TM2 - 1) / TM2);
    const RowWin win = row_window(c);
    const int tm_lo = (int)(win.r0 / TM2), tm_hi = (int)std::min<int64_t>(nt, (win.r1 + TM2 - 1) / TM2);
    if (tm_hi <= tm_lo) return;
    // SNP range of this launch (default: everything resident); boundaries are stage aligned
    if (snp_hi < 0) snp_hi = c->n_snp;
    if (snp_lo % SK) fail("internal: gram_tc_run range must start at a multiple of %d SNPs", SK);
    const int st_lo = (int)(snp_lo / SK);
    const int st_hi = (int)(round_up(std::max<int64_t>(snp_hi, snp_lo + 1), SK) / SK);
    const int stages_total = st_hi - st_lo;
    const int64_t slots = std::max(1, c->num_sms / 2);

    // ---- pass groups: two passes that share a B table fill the two accumulators of a 256 x 256
    // tile (mode 0); a pass left over runs alone over 256 x 512 tiles (mode 1)
    struct Group { int mode; int pass[2]; };
    std::vector<Group> groups;
    std::vector<PassDesc> pd((size_t)npass);
    for (int i = 0; i < npass; i++) pd[i] = PassDesc{passes[i].tabA, passes[i].tabB, passes[i].plane, passes[i].shift};
    {
        std::vector<char> used((size_t)npass, 0);
        for (int i = 0; i < npass; i++) {
            if (used[i]) continue;
            used[i] = 1;
            int mate = -1;
            for (int j = i + 1; j < npass && mate < 0; j++)   // same plane first: their atomics are combined
                if (!used[j] && passes[j].tabB == passes[i].tabB && passes[j].plane == passes[i].plane) mate = j;
            for (int j = i + 1; j < npass && mate < 0; j++)
                if (!used[j] && passes[j].tabB == passes[i].tabB) mate = j;
            if (mate >= 0) {
                used[mate] = 1;
                groups.push_back({0, {i, mate}});
            } else {
                groups.push_back({1, {i, i}});
            }
        }
    }

    // ---- SNP segments per group.  (a) int32 accumulator headroom: |digit| <= 128 and a bound on
    // sum_l |B_l[g_jl]| per 512-SNP chunk (measured by the caller, or max|B| x 512);  (b) enough
    // items to fill the chip when the tile grid alone is small.
    int64_t base_items = 0;
    for (int tm = tm_lo; tm < tm_hi; tm++) {
        const int ncol = nt - (upper_only ? tm : 0);
        for (auto &g : groups) base_items += g.mode == 0 ? ncol : (ncol + 1) / 2;
    }
    int parts = 1;
    if (!(c->debug_flags & 2u) && base_items < 8 * slots)
        parts = (int)std::min<int64_t>(std::min<int64_t>((8 * slots + base_items - 1) / base_items, 64),
                                       std::max(1, stages_total / 4));
    // Large tile grids: cut the SNP range into segments of ~1024 stages (131072 SNPs) and emit the items
    // SEGMENT-MAJOR inside every super-tile, so that everything co-resident (and the next few waves)
    // walks the same 8 MB slice of at most RB + CB genotype panels -- a working set the 126 MB L2 holds.
    // Without it the ~74 resident items drift apart along their 1M-SNP walks and every panel is
    // streamed from HBM again by every item (measured at 10k x 1M: 522 GB -> 64 GB of DRAM reads per
    // step, SM clock under the power cap 1547 -> 1677 MHz, kernel -4 %; profiles/r02_notes.md).
    bool seg_major = false;
    if (!(c->debug_flags & 2u) && parts == 1 && !(c->debug_flags & 256u)) {
        parts = std::max(1, (stages_total + 512) / 1024);
        seg_major = parts > 1;
    }
    if (c->debug_flags & 0xF000u) {     // experiment: forced number of segments
        parts = (int)((c->debug_flags >> 12) & 15u);
        seg_major = parts > 1;
    }
    std::vector<std::vector<int>> cuts(groups.size());   // stage boundaries, ascending, first 0, last stages_total
    const int CH = GRAM_CHUNK / SK;                        // stages per bound chunk
    for (size_t gi = 0; gi < groups.size(); gi++) {
        const GramPass &ps = passes[groups[gi].pass[0]];
        std::vector<int> &cu = cuts[gi];
        cu.push_back(st_lo);
        std::vector<int> want;                             // wished part boundaries, multiples of a chunk
        for (int k = 1; k < parts; k++) {
            const int bnd = (st_lo + (int)((long long)k * stages_total / parts)) / CH * CH;
            if (bnd > st_lo && bnd < st_hi && (want.empty() || bnd > want.back())) want.push_back(bnd);
        }
        size_t wi = 0;
        long long acc = 0;
        for (int st = st_lo / CH * CH; st < st_hi; st += CH) {
            const size_t ck = (size_t)(st / CH);
            long long cb = (ps.b_chunk_bound && ck < ps.b_chunk_bound->size())
                               ? std::min<long long>((*ps.b_chunk_bound)[ck], (long long)ps.b_abs_max * GRAM_CHUNK)
                               : (long long)ps.b_abs_max * GRAM_CHUNK;
            cb *= 128;
            if (cb > 2147483647ll) fail("internal: one SNP chunk overflows the int32 accumulator");
            const bool part_cut = wi < want.size() && st >= want[wi];
            if (st > cu.back() && (acc + cb > 2147483647ll || part_cut)) {
                cu.push_back(st);
                acc = 0;
            }
            while (wi < want.size() && st >= want[wi]) wi++;
            acc += cb;
        }
        cu.push_back(st_hi);
    }

    // ---- the item list: tile-major (all pass groups of a tile are neighbours, so the co-resident
    // CTA pairs share genotype panels in L2), tiles in a super-tile raster
    std::vector<Item> items;
    auto width = [&](int t) { return tile_ncols(std::min<int64_t>(TN2, n - (int64_t)t * TN2)); };
    int only_group = -1, only_seg = -1;
    auto emit = [&](int tm, int tn) {
        const int first_col = upper_only ? tm : 0;
        for (size_t gi = 0; gi < groups.size(); gi++) {
            if (only_group >= 0 && (int)gi != only_group) continue;
            const Group &g = groups[gi];
            Item it{};
            it.tm = tm;
            it.mode = (short)g.mode;
            it.pass[0] = g.pass[0];
            it.pass[1] = g.pass[1];
            if (g.mode == 0) {
                it.tn = tn;
                it.ncols[0] = (short)width(tn);
                it.ncols[1] = 0;
            } else {
                const int tn2 = tn / 2;
                if (tn != std::max(2 * tn2, first_col)) continue;   // the pair is emitted at its first live column
                it.tn = 2 * tn2;
                it.ncols[0] = (short)(2 * tn2 >= first_col ? width(2 * tn2) : 0);
                it.ncols[1] = (short)(2 * tn2 + 1 < nt ? width(2 * tn2 + 1) : 0);
            }
            for (size_t k = 0; k + 1 < cuts[gi].size(); k++) {
                if (only_seg >= 0 && (int)k != only_seg) continue;
                it.st_begin = cuts[gi][k];
                it.st_end = cuts[gi][k + 1];
                items.push_back(it);
            }
        }
    };
    constexpr int RB = 8, CB = 8;
    const size_t g_lo = 0, g_hi = groups.size();
    auto raster = [&]() {
        for (int rb = tm_lo / RB * RB; rb < tm_hi; rb += RB)
            for (int cb = (upper_only ? rb / CB * CB : 0); cb < nt; cb += CB)
                for (int tm = std::max(rb, tm_lo); tm < std::min(rb + RB, tm_hi); tm++)
                    for (int tn = std::max(cb, upper_only ? tm : 0); tn < std::min(cb + CB, nt); tn++) emit(tm, tn);
    };
    if (seg_major) {
        size_t nseg = 0;
        for (auto &cu : cuts) nseg = std::max(nseg, cu.size() - 1);
        for (int rb = tm_lo / RB * RB; rb < tm_hi; rb += RB)
            for (int cb = (upper_only ? rb / CB * CB : 0); cb < nt; cb += CB)
                for (size_t sg = 0; sg < nseg; sg++) {
                    only_seg = (int)sg;
                    for (int tm = std::max(rb, tm_lo); tm < std::min(rb + RB, tm_hi); tm++)
                        for (int tn = std::max(cb, upper_only ? tm : 0); tn < std::min(cb + CB, nt); tn++) emit(tm, tn);
                }
        only_seg = -1;
    } else if (c->debug_flags & 128u) {      // experiment: group-major (all tiles of one pass group, then the next)
        for (size_t g = g_lo; g < g_hi; g++) {
            only_group = (int)g;
            raster();
        }
        only_group = -1;
    } else {
        raster();
    }
    if (items.empty()) return;
    // wave tail: the items that start last are cut in four, so the chip drains in quarter steps
    if (!(c->debug_flags & 2u) && parts == 1 && !seg_major && (int64_t)items.size() > 4 * slots) {
        const size_t ntail = (size_t)(slots + slots / 2);
        std::vector<Item> tail(items.end() - ntail, items.end());
        items.resize(items.size() - ntail);
        for (const Item &it : tail) {
            const int len = it.st_end - it.st_begin;
            if (len < 64) {
                items.push_back(it);
                continue;
            }
            for (int q = 0; q < 4; q++) {
                Item sub = it;
                sub.st_begin = it.st_begin + (int)((long long)len * q / 4);
                sub.st_end = it.st_begin + (int)((long long)len * (q + 1) / 4);
                items.push_back(sub);
            }
        }
    }
    if (items.size() > 0x3fffffffull) fail("too many work items");

    // Work list and pass descriptors go to the device WITHOUT the host-to-device copy engine: a copy queued
    // on the compute stream would wait behind every genotype chunk still crossing PCIe (one H2D engine
    // queue), and the first tensor pass with it.  The host writes them into mapped pinned memory and a tiny
    // kernel of this stream pulls them into device memory (each un-synchronised launch gets its own slice).
    const size_t ibytes = round_up((int64_t)(items.size() * sizeof(Item)), 256), pbytes = round_up((int64_t)(pd.size() * sizeof(PassDesc)), 256);
    if (c->stage_used + ibytes + pbytes > c->stage_bytes) {
        CUDA_CHECK(cudaStreamSynchronize(c->stream));            // nothing may still read the old buffers
        const size_t want = std::max<size_t>(2 * (c->stage_used + ibytes + pbytes), 4u << 20);
        if (c->stage_host) cudaFreeHost(c->stage_host);
        c->stage_host = nullptr;
        CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void **>(&c->stage_host), want, cudaHostAllocMapped));
        c->stage_bytes = want;
        c->stage_used = 0;
        c->scr_items.alloc(want);
    }
    uint8_t *hslice = c->stage_host + c->stage_used, *dslice = c->scr_items.p + c->stage_used;
    memcpy(hslice, items.data(), items.size() * sizeof(Item));
    memcpy(hslice + ibytes, pd.data(), pd.size() * sizeof(PassDesc));
    {
        const int64_t words = (int64_t)((ibytes + pbytes) / 16);
        pull_staged_kernel<<<(unsigned)((words + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<const uint4 *>(hslice),
                                                                                reinterpret_cast<uint4 *>(dslice), words);
        KERNEL_CHECK(c);
    }
    c->stage_used += ibytes + pbytes;
    c->scr_flags.alloc(2);
    int *derr = c->scr_flags.p + 1;
    if (sync || snp_lo == 0) CUDA_CHECK(cudaMemsetAsync(derr, 0, sizeof(int), c->stream));

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
    const bool box32 = (c->debug_flags & 16u) == 0;   // flag 16: the round-1 16-byte boxes (A/B experiments)
    {
        const int64_t rows = round_up(std::max<int64_t>(c->n_snp, 1), SK);
        cuuint64_t gdim[2] = {(cuuint64_t)c->row_bytes, (cuuint64_t)rows};
        cuuint64_t gstride[1] = {(cuuint64_t)c->row_bytes};
        cuuint32_t box[2] = {box32 ? 32u : 16u, (cuuint32_t)SK};
        cuuint32_t estr[2] = {1, 1};
        const CUtensorMapL2promotion promo = (c->debug_flags & 32u)   ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                             : (c->debug_flags & 64u) ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                                      : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, c->geno2b.p, gdim, gstride, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, box32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE,
                            promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) fail("cuTensorMapEncodeTiled failed (%d)", (int)r);
    }

    if (!c->gram_attr_done) {   // per device: a process may hold contexts on several GPUs
        CUDA_CHECK(cudaFuncSetAttribute(table_gram_kernel3, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        CUDA_CHECK(cudaFuncSetAttribute(table_gram_kernel3_traced, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        c->gram_attr_done = true;
    }

    Params P{};
    P.items = reinterpret_cast<const Item *>(dslice);
    P.passes = reinterpret_cast<const PassDesc *>(dslice + ibytes);
    P.out = out_planes;
    P.ld = npad;
    P.plane_stride = win.rows * npad;
    P.row0 = win.r0;
    P.n_samp = n;
    P.upper_only = upper_only ? 1 : 0;
    P.sh32 = 32;
    P.box32 = box32 ? 1 : 0;
    P.error_flag = derr;
    P.trace = nullptr;
    if (c->debug_flags & 1u) {          // per-item clock stamps for tools/k1_trace.py
        c->scr_trace.alloc(items.size() * 8);
        c->scr_trace.zero(c->stream);
        P.trace = c->scr_trace.p;
        c->trace_items = (int64_t)items.size();
    }
    if (P.trace)
        table_gram_kernel3_traced<<<dim3((unsigned)(2 * items.size())), THREADS, SMEM_BYTES, c->stream>>>(P, tmap);
    else
        table_gram_kernel3<<<dim3((unsigned)(2 * items.size())), THREADS, SMEM_BYTES, c->stream>>>(P, tmap);
    KERNEL_CHECK(c);
    c->hot_launches++;
    c->hot_items = (int64_t)items.size();
    if (sync) gram_tc_check(c);
}

void gram_tc_stage_reset(snprel_ctx *c) { c->stage_used = 0; }

// wait for the launches queued so far and surface a pipeline time-out
void gram_tc_check(snprel_ctx *c) {
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->stage_used = 0;
    if (!c->scr_flags.p) return;
    int herr = 0;
    CUDA_CHECK(cudaMemcpy(&herr, c->scr_flags.p + 1, sizeof(int), cudaMemcpyDeviceToHost));
    if (herr) fail("table_gram_kernel3: pipeline barrier %d timed out", herr);
}

}  // namespace snprel
