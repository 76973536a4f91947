// Measured ceilings for the two kernel families ("measure the same way before
// quoting a fraction").  Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/peaks.cu
// -o build/peaks_r02 ; prints one JSON object.
//
//  1. dense int8 tcgen05 issue peak: CTA pairs (cta_group::2, M=256 N=256 K=32, kind::i8) re-issue the
//     MMAs of K1's stage loop on RESIDENT shared-memory operands (no producers, no TMA, no epilogue),
//     with (a) random int8 operands, (b) all-zero operands.  (a) is the ceiling K1 can reach under the
//     power cap; (b) shows how much of it is data-dependent power.
//  2. LOP3 and POPC issue rates per SM and clock (the pair-count kernels' two pipes), and the
//     IBS-like 7:1 / KING-like 11:3.3 mixes.
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../snprelate_b200/csrc/tc_ptx.cuh"

using namespace snprel::tc;

#define CK(x)                                                                             \
    do {                                                                                  \
        cudaError_t e_ = (x);                                                             \
        if (e_ != cudaSuccess) {                                                          \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));    \
            exit(1);                                                                      \
        }                                                                                 \
    } while (0)

// ---------------------------------------------------------------------------------------------
constexpr int HM = 128, SK = 128, MMA_K = 32, NSTAGE = 4;
constexpr int A_BYTES = HM * SK;
constexpr int STAGE_BYTES = 3 * A_BYTES;
constexpr int LBO = (HM / 16) * 128, SBO = 128;
constexpr int BAR_OFFSET = NSTAGE * STAGE_BYTES;
constexpr int SMEM_BYTES = BAR_OFFSET + 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {
    asm volatile(
        "{\n .reg .b16 m;\n mov.b16 m, 3;\n"
        " tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n}" ::"r"(bar)
        : "memory");
}

// one CTA pair per cluster; `nstage` stages of 4 K-steps x 2 accumulators each (exactly K1's issue pattern)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
int8_peak_kernel(int nstage, uint32_t seed, int zero, int *sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_base + BAR_OFFSET;
    auto empty_bar = [&](int s) { return bar_base + 8u * s; };
    const uint32_t tmem_slot = bar_base + 8u * (NSTAGE + 1);
    volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem + BAR_OFFSET + 8 * (NSTAGE + 1));
    // resident operands
    uint32_t x = seed ^ (blockIdx.x * 0x9E3779B9u) ^ (threadIdx.x * 0x85EBCA6Bu);
    for (int i = threadIdx.x; i < NSTAGE * STAGE_BYTES / 4; i += blockDim.x) {
        x = x * 1664525u + 1013904223u;
        uint32_t v = x ^ (x >> 13);
        reinterpret_cast<uint32_t *>(smem)[i] = zero ? 0u : v;
    }
    fence_proxy_async_smem();
    if (warp == 0) {
        if (lane == 0) {
            for (int s = 0; s < NSTAGE; s++) mbar_init(empty_bar(s), 1);
            mbar_init(empty_bar(NSTAGE), 1);
            fence_mbar_init();
        }
        __syncwarp();
        tmem_alloc2(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    if (warp == 0 && rank == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(256 >> 3) << 17) |
                               ((uint32_t)(256 >> 4) << 24);
        for (int it = 0; it < nstage; it++) {
            const int s = it % NSTAGE;
            if (it >= NSTAGE) mbar_wait(empty_bar(s), (uint32_t)((it / NSTAGE) - 1) & 1u, nullptr, 0);
            tc_fence_after();
            if (lane == 0) {
                const uint32_t stage_addr = smem_base + s * STAGE_BYTES;
#pragma unroll
                for (int j = 0; j < SK / MMA_K; j++) {
                    uint64_t bdesc = make_desc(stage_addr + 2 * A_BYTES + j * 4 * LBO, LBO, SBO);
#pragma unroll
                    for (int p = 0; p < 2; p++) {
                        uint64_t adesc = make_desc(stage_addr + p * A_BYTES + j * 4 * LBO, LBO, SBO);
                        umma2_i8(tmem_base + (uint32_t)(p * 256), adesc, bdesc, idesc, (it > 0 || j > 0) ? 1u : 0u);
                    }
                }
                umma2_commit_mc(empty_bar(s));
            }
            __syncwarp();
        }
        if (lane == 0) umma2_commit_mc(empty_bar(NSTAGE));
        __syncwarp();
    }
    if (warp == 1) {
        // both CTAs wait for the last commit (multicast), then read one value so the work is observable
        mbar_wait(empty_bar(NSTAGE), 0, nullptr, 0);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tmem_base, v);
        if (v[0] == 0x12345678u && sink) sink[0] = 1;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------
// ALU / XU issue microbenchmarks: 8 independent chains per thread, ITER x UNROLL steps
template <int MODE>
__global__ void __launch_bounds__(256) alu_kernel(uint32_t *out, int iters, uint32_t a0) {
    uint32_t r[8];
#pragma unroll
    for (int k = 0; k < 8; k++) r[k] = a0 + threadIdx.x * 977u + k * 131u + blockIdx.x;
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint32_t c1 = a0 * 3u + 1u, c2 = a0 ^ 0x5bd1e995u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (MODE == 0) {            // LOP3 only: r = (r & c1) ^ c2-ish three-input op, dependent per chain
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[k]) : "r"(c1), "r"(acc[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(acc[k]) : "r"(r[k]), "r"(c2));
                } else if (MODE == 1) {     // POPC only: two dependent chains per register pair, nothing else in the loop
                    asm volatile("popc.b32 %0, %0;" : "+r"(r[k]));
                    asm volatile("popc.b32 %0, %0;" : "+r"(acc[k]));
                } else {                    // IBS-like mix: 7 LOP3 per POPC (+ 1 IADD)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[k]) : "r"(c1), "r"(acc[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(r[k]) : "r"(c2), "r"(acc[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[k]) : "r"(c1), "r"(c2));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(r[k]) : "r"(c1), "r"(acc[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[k]) : "r"(c2), "r"(acc[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(r[k]) : "r"(c1), "r"(c2));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[k]) : "r"(c1), "r"(acc[k]));
                    uint32_t p;
                    asm volatile("popc.b32 %0, %1;" : "=r"(p) : "r"(r[k]));
                    acc[k] += p;
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s ^= r[k] ^ acc[k];
    if (s == 0xdeadbeefu) out[0] = s;
}

static float time_kernel(void (*launch)(void *), void *arg, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    launch(arg);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0));
        launch(arg);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

struct I8Arg { int grid, nstage, zero; int *sink; };
static void launch_i8(void *p) {
    I8Arg *a = (I8Arg *)p;
    int8_peak_kernel<<<a->grid, 128, SMEM_BYTES>>>(a->nstage, 12345u, a->zero, a->sink);
}
struct AluArg { int mode, grid, iters; uint32_t *out; };
static void launch_alu(void *p) {
    AluArg *a = (AluArg *)p;
    if (a->mode == 0) alu_kernel<0><<<a->grid, 256>>>(a->out, a->iters, 7u);
    else if (a->mode == 1) alu_kernel<1><<<a->grid, 256>>>(a->out, a->iters, 7u);
    else alu_kernel<2><<<a->grid, 256>>>(a->out, a->iters, 7u);
}

int main(int argc, char **argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    int clock_khz = 0;
    CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev));
    int *sink;
    CK(cudaMalloc(&sink, 64));
    CK(cudaMemset(sink, 0, 64));
    CK(cudaFuncSetAttribute(int8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_max_mhz\": %d", prop.name, sms, clock_khz / 1000);

    // ---- int8 tcgen05: short burst (about 25 ms) and sustained (about 1 s, power cap in effect)
    const double ops_per_stage = 2.0 * 256 * 256 * 128 * 2;   // 4 K-steps x 2 accumulators of 256 x 256 x 32 MACs
    for (int zero = 0; zero < 2; zero++) {
        for (int len = 0; len < 2; len++) {
            I8Arg a{sms / 2 * 2, len == 0 ? 20000 : 800000, zero, sink};
            float ms = time_kernel(launch_i8, &a, len == 0 ? 5 : 2);
            double tops = ops_per_stage * a.nstage * (a.grid / 2) / (ms * 1e-3) / 1e12;
            printf(", \"int8_tcgen05_%s_%s\": {\"tops\": %.1f, \"ms\": %.2f, \"cta_pairs\": %d, \"stages\": %d}",
                   zero ? "zero_operands" : "random_operands", len == 0 ? "burst" : "sustained", tops, ms, a.grid / 2, a.nstage);
            fflush(stdout);
        }
    }
    // ---- ALU / XU issue
    uint32_t *out;
    CK(cudaMalloc(&out, 64));
    const char *names[3] = {"lop3", "popc", "mix_7lop3_1popc_1iadd"};
    const double ops_per_inner[3] = {2.0, 2.0, 9.0};   // instructions per chain step as written
    for (int mode = 0; mode < 3; mode++) {
        AluArg a{mode, sms * 8, 2000, out};
        float ms = time_kernel(launch_alu, &a, 5);
        double inst = (double)a.grid * 256 * (double)a.iters * 64 * ops_per_inner[mode];   // thread-level instructions
        double per_s = inst / (ms * 1e-3);
        printf(", \"alu_%s\": {\"thread_inst_per_s\": %.4g, \"per_sm_per_clk_at_max\": %.1f, \"ms\": %.3f}", names[mode], per_s,
               per_s / sms / (clock_khz * 1e3), ms);
        fflush(stdout);
    }
    printf("}\n");
    return 0;
}
