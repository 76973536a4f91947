// Genotype workspace on the device: the B200 counterpart of the reference's
// CdBaseWorkSpace / CGenoReadBySNP / PackSNPGeno1b layer
// (src/dGenGWAS.h:80-150,295-355; src/dGenGWAS.cpp:361-397,472-552,1218-1475).
//
// HBM layout: 2-bit genotypes (GDS dBit2 code: 0/1/2 = #A alleles, 3 = missing),
// SNP-major, sample fastest, 4 genotypes per byte LSB first, every SNP row padded
// with missing codes to a multiple of 256 samples; SNP rows padded with
// all-missing rows to a multiple of 128.  N*M/4 bytes -- a 10k x 1M data set is
// 2.5 GB, 500k x 800k is 100 GB (fits one 180 GB B200).
#include "common.cuh"
#include <cmath>

namespace snprel {

// ---------------------------------------------------------------------------
// pack: uint8 [cnt][n_samp] -> 2-bit rows
// ---------------------------------------------------------------------------
__global__ void pack_u8_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                               int64_t cnt, int64_t n_samp, int64_t row_bytes) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = cnt * row_bytes;
    if (idx >= total) return;
    int64_t r = idx / row_bytes, b = idx - r * row_bytes;
    const uint8_t *s = src + r * n_samp + b * 4;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int64_t i = b * 4 + k;
        uint32_t g = 3;
        if (i < n_samp) {
            g = s[k];
            if (g > 2) g = 3;   // vec_u8_geno_valid: anything > 2 is missing (src/dVect.cpp:121-144)
        }
        out |= g << (2 * k);
    }
    dst[idx] = (uint8_t)out;
}

// bytes [pad_from, row_bytes) of every row: keep the valid genotypes of a partial byte, set the
// rest (and every byte the host copy did not cover) to the missing code
__global__ void fix_pad_kernel(uint8_t *__restrict__ dst, int64_t cnt, int64_t n_samp, int64_t pad_from,
                               int64_t pad_bytes, int64_t copied, int64_t row_bytes) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cnt * pad_bytes) return;
    int64_t r = idx / pad_bytes, b = pad_from + (idx - r * pad_bytes);
    uint32_t v = 0xFF;
    int64_t rem = n_samp - b * 4;   // valid genotypes in this byte
    if (rem > 0 && b < copied) v = dst[r * row_bytes + b] | ((0xFFu << (2 * rem)) & 0xFF);
    dst[r * row_bytes + b] = (uint8_t)v;
}

__global__ void unpack_u8_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                 int64_t cnt, int64_t n_samp, int64_t row_bytes) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cnt * n_samp) return;
    int64_t r = idx / n_samp, i = idx - r * n_samp;
    dst[idx] = (src[r * row_bytes + (i >> 2)] >> (2 * (i & 3))) & 3;
}

// ---------------------------------------------------------------------------
// synthetic generator (bit-identical to oracle/snprel_oracle.py:synth_geno)
// ---------------------------------------------------------------------------
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void synth_kernel(uint8_t *__restrict__ dst, int64_t n_snp, int64_t n_samp,
                             int64_t row_bytes, uint64_t seed, double maf_lo, double maf_hi,
                             uint32_t thm, int64_t snp_start) {
    // one thread per output byte (4 samples); grid.y = SNP row
    int64_t r = blockIdx.y;
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_snp || b >= row_bytes) return;
    uint64_t l = (uint64_t)(r + snp_start);
    uint64_t hp = splitmix64(seed ^ (l * 0xD1342543DE82EF95ull));
    double u = __dmul_rn((double)(hp >> 11), 1.0 / 9007199254740992.0);
    double p = __dadd_rn(maf_lo, __dmul_rn(__dsub_rn(maf_hi, maf_lo), u));
    double q = __dsub_rn(1.0, p);
    double t0 = __dmul_rn(q, q);
    double t1 = __dadd_rn(t0, __dmul_rn(__dmul_rn(2.0, p), q));
    double f0 = fmin(floor(__dmul_rn(t0, 4294967296.0)), 4294967295.0);
    double f1 = fmin(floor(__dmul_rn(t1, 4294967296.0)), 4294967295.0);
    uint32_t th0 = (uint32_t)f0, th1 = (uint32_t)f1;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int64_t i = b * 4 + k;
        uint32_t g = 3;
        if (i < n_samp) {
            uint64_t key = splitmix64(hp + (uint64_t)i * 0x9E3779B97F4A7C15ull);
            uint32_t rr = (uint32_t)(key >> 32), rm = (uint32_t)key;
            g = (rr >= th0) + (rr >= th1);
            if (rm < thm) g = 3;
        }
        out |= g << (2 * k);
    }
    dst[r * row_bytes + b] = (uint8_t)out;
}

// ---------------------------------------------------------------------------
// per-SNP statistics: vec_u8_geno_count (src/dVect.cpp:30-117) on packed rows.
// One warp per SNP row, uint4 loads (64 samples per lane per step).
// ---------------------------------------------------------------------------
__global__ void snp_stat_kernel(const uint8_t *__restrict__ g, SnpStat *__restrict__ st,
                                int64_t n_rows, int64_t row_bytes, int n_samp_pad) {
    int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    int lane = threadIdx.x & 31;
    const uint4 *p = reinterpret_cast<const uint4 *>(g + row * row_bytes);
    int nq = (int)(row_bytes >> 4);
    int nmiss = 0, n1 = 0, n2 = 0;
    for (int q = lane; q < nq; q += 32) {
        uint4 v = __ldg(p + q);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t lo = w[k] & 0x55555555u, hi = (w[k] >> 1) & 0x55555555u;
            nmiss += __popc(lo & hi);
            n1 += __popc(lo & ~hi);
            n2 += __popc(hi & ~lo);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        nmiss += __shfl_xor_sync(0xffffffffu, nmiss, o);
        n1 += __shfl_xor_sync(0xffffffffu, n1, o);
        n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    }
    if (lane == 0) {
        SnpStat s;
        s.sum = n1 + 2 * n2;
        s.num = n_samp_pad - nmiss;
        s.n1 = n1;
        s.pad = 0;
        st[row] = s;
    }
}

// gather SNP rows (compaction after gnrSelSNP_Base)
__global__ void gather_rows_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst,
                                   const int64_t *__restrict__ idx, int64_t n_out,
                                   int64_t row_bytes) {
    int64_t r = blockIdx.y;
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_out || q >= (row_bytes >> 4)) return;
    const uint4 *s = reinterpret_cast<const uint4 *>(src + idx[r] * row_bytes);
    uint4 *d = reinterpret_cast<uint4 *>(dst + r * row_bytes);
    d[q] = s[q];
}

// ---------------------------------------------------------------------------
// 2-bit -> bit planes: the device PackSNPGeno1b (src/dGenGWAS.cpp:1429-1475).
// Encoding 0->(0,0) 1->(1,0) 2->(1,1) NA->(0,1) as (plane1, plane2); padding SNPs
// are all-missing rows, i.e. (0,1) as the reference requires (:1467-1472).
// Output [word64][sample] of uint4 {p1.lo, p1.hi, p2.lo, p2.hi}: the pair kernel
// reads one coalesced uint4 per (sample, 64-SNP word).
// ---------------------------------------------------------------------------
__global__ void planes_kernel(const uint8_t *__restrict__ g, uint4 *__restrict__ planes,
                              int64_t n_words, int64_t row_bytes, int64_t n_samp_pad) {
    int64_t w = blockIdx.y;
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // byte column = 4 samples
    if (w >= n_words || b >= row_bytes) return;
    unsigned long long p1[4] = {0, 0, 0, 0}, p2[4] = {0, 0, 0, 0};
    const uint8_t *src = g + (w * 64) * row_bytes + b;
#pragma unroll 8
    for (int s = 0; s < 64; s++) {
        uint32_t v = src[(int64_t)s * row_bytes];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t code = (v >> (2 * k)) & 3;
            unsigned long long lo = code & 1, hi = code >> 1;
            p1[k] |= (lo ^ hi) << s;
            p2[k] |= hi << s;
        }
    }
    uint4 *dst = planes + w * n_samp_pad + b * 4;
#pragma unroll
    for (int k = 0; k < 4; k++)
        dst[k] = make_uint4((uint32_t)p1[k], (uint32_t)(p1[k] >> 32), (uint32_t)p2[k],
                            (uint32_t)(p2[k] >> 32));
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
void geno_begin(snprel_ctx *c, int64_t n_samp, int64_t cap) {
    geno_wait(c);
    output_wait(c);
    if (n_samp <= 0) fail("snprel_geno_begin: n_samp must be positive");
    if (cap < 0) fail("snprel_geno_begin: negative SNP capacity");
    if (n_samp >= (1ll << 30)) fail("snprel_geno_begin: too many samples");
    c->n_samp = n_samp;
    c->n_samp_pad = round_up(n_samp, SAMP_PAD);
    c->row_bytes = c->n_samp_pad / 4;
    c->snp_cap = round_up(cap > 0 ? cap : 1, SNP_PAD);
    c->n_snp = 0;
    // same or smaller workspace: keep the allocation (a cudaFree + cudaMalloc of a few GB costs
    // 3-13 ms per call); a much smaller one gives the memory back
    const size_t need = (size_t)c->snp_cap * c->row_bytes;
    if (c->geno2b.n > 2 * need + (64u << 20)) c->geno2b.release();
    if (c->scr_out.n > 2 * (size_t)c->n_samp_pad * c->n_samp_pad) c->scr_out.release();
    c->geno2b.alloc(need);
    c->stat.alloc(c->snp_cap);
    drop_derived(c);
}

static void need_room(snprel_ctx *c, int64_t cnt, const char *who) {
    if (c->n_samp <= 0) fail("%s: no genotype workspace (call snprel_geno_begin first)", who);
    if (cnt < 0) fail("%s: negative SNP count", who);
    if (c->n_snp + cnt > c->snp_cap)
        fail("%s: SNP capacity exceeded (%lld + %lld > %lld)", who, (long long)c->n_snp,
             (long long)cnt, (long long)c->snp_cap);
}

static void invalidate(snprel_ctx *c) { drop_derived(c); }

void geno_push_u8(snprel_ctx *c, const uint8_t *host, int64_t cnt) {
    need_room(c, cnt, "snprel_geno_push_u8");
    if (cnt == 0) return;
    if (!host) fail("snprel_geno_push_u8: NULL block");
    // stream the block through a bounded staging buffer so host pages are touched once
    const int64_t max_rows = std::max<int64_t>(1, (int64_t)(256ll << 20) / c->n_samp);
    c->stage_u8.alloc((size_t)std::min(cnt, max_rows) * c->n_samp);
    for (int64_t done = 0; done < cnt; done += max_rows) {
        int64_t rows = std::min(max_rows, cnt - done);
        CUDA_CHECK(cudaMemcpyAsync(c->stage_u8.p, host + done * c->n_samp, (size_t)rows * c->n_samp,
                                   cudaMemcpyHostToDevice, c->stream));
        int64_t total = rows * c->row_bytes;
        pack_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(
            c->stage_u8.p, c->geno2b.p + (c->n_snp + done) * c->row_bytes, rows, c->n_samp,
            c->row_bytes);
        KERNEL_CHECK(c);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));   // staging buffer is reused
    }
    c->n_snp += cnt;
    invalidate(c);
}

void geno_push_2b(snprel_ctx *c, const uint8_t *host, int64_t cnt, int64_t row_bytes_in) {
    need_room(c, cnt, "snprel_geno_push_2b");
    if (cnt == 0) return;
    if (!host) fail("snprel_geno_push_2b: NULL block");
    if (row_bytes_in < (c->n_samp + 3) / 4)
        fail("snprel_geno_push_2b: row_bytes %lld too small for %lld samples",
             (long long)row_bytes_in, (long long)c->n_samp);
    uint8_t *dst = c->geno2b.p + c->n_snp * c->row_bytes;
    // one strided DMA straight into the padded device rows (no staging copy), then a small
    // kernel rewrites the bytes at / beyond the last sample so that padding reads as missing
    const int64_t w = (c->n_samp + 3) / 4;
    if (row_bytes_in == c->row_bytes)   // host rows already have the device pitch: one linear copy
        CUDA_CHECK(cudaMemcpyAsync(dst, host, (size_t)cnt * c->row_bytes, cudaMemcpyHostToDevice, c->stream));
    else
        CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)c->row_bytes, host, (size_t)row_bytes_in, (size_t)w, (size_t)cnt,
                                     cudaMemcpyHostToDevice, c->stream));
    const int64_t pad_from = c->n_samp / 4;              // first byte that holds any padding sample
    const int64_t pad_bytes = c->row_bytes - pad_from;
    if (pad_bytes > 0) {
        int64_t total = cnt * pad_bytes;
        fix_pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(
            dst, cnt, c->n_samp, pad_from, pad_bytes, row_bytes_in == c->row_bytes ? c->row_bytes : w, c->row_bytes);
        KERNEL_CHECK(c);
    }
    CUDA_CHECK(cudaStreamSynchronize(c->stream));   // the host block may be reused by the caller
    c->n_snp += cnt;
    invalidate(c);
}

// The same push without waiting for the copies: the block goes out in chunks that end at multiples of
// STREAM_CHUNK rows (the K1 segment) on a second stream, each followed by an event.  The host block must
// stay valid (and should be pinned) until the next library call on this context returns.
void geno_push_2b_async(snprel_ctx *c, const uint8_t *host, int64_t cnt, int64_t row_bytes_in) {
    need_room(c, cnt, "snprel_geno_push_2b_async");
    if (cnt == 0) return;
    if (!host) fail("snprel_geno_push_2b_async: NULL block");
    if (row_bytes_in < (c->n_samp + 3) / 4)
        fail("snprel_geno_push_2b_async: row_bytes %lld too small for %lld samples", (long long)row_bytes_in,
             (long long)c->n_samp);
    if (!c->copy_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    // anything queued on the compute stream that still reads these rows must finish first
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    const int64_t w = (c->n_samp + 3) / 4, pad_bytes = c->row_bytes - c->n_samp / 4;
    if (!c->copy_ev0) CUDA_CHECK(cudaEventCreate(&c->copy_ev0));
    if (c->pending.empty()) CUDA_CHECK(cudaEventRecord(c->copy_ev0, c->copy_stream));
    for (int64_t l0 = c->n_snp, end = c->n_snp + cnt; l0 < end;) {
        const int64_t l1 = std::min(end, (l0 / STREAM_CHUNK + 1) * STREAM_CHUNK), rows = l1 - l0;
        uint8_t *dst = c->geno2b.p + l0 * c->row_bytes;
        const uint8_t *src = host + (l0 - c->n_snp) * row_bytes_in;
        if (row_bytes_in == c->row_bytes)
            CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)rows * c->row_bytes, cudaMemcpyHostToDevice, c->copy_stream));
        else
            CUDA_CHECK(cudaMemcpy2DAsync(dst, (size_t)c->row_bytes, src, (size_t)row_bytes_in, (size_t)w, (size_t)rows,
                                         cudaMemcpyHostToDevice, c->copy_stream));
        // (the padding bytes of the rows are rewritten by whoever consumes the chunk, on the compute stream:
        //  a kernel on the copy stream would queue behind the resident tensor-pass CTAs and stall the copies)
        cudaEvent_t ev;
        CUDA_CHECK(cudaEventCreate(&ev));
        CUDA_CHECK(cudaEventRecord(ev, c->copy_stream));
        c->pending.push_back({l0, l1, ev, pad_bytes > 0 ? (row_bytes_in == c->row_bytes ? c->row_bytes : w) : (int64_t)-1});
        l0 = l1;
    }
    c->n_snp += cnt;
    invalidate(c);
}

// padding bytes of an arrived chunk (compute stream; the caller has ordered it after the chunk's event)
void geno_fix_chunk_padding(snprel_ctx *c, const snprel_ctx::PendingCopy &p) {
    if (p.copied < 0) return;
    const int64_t pad_from = c->n_samp / 4, pad_bytes = c->row_bytes - pad_from, rows = p.l1 - p.l0;
    const int64_t total = rows * pad_bytes;
    if (total <= 0) return;
    fix_pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(c->geno2b.p + p.l0 * c->row_bytes, rows, c->n_samp,
                                                                         pad_from, pad_bytes, p.copied, c->row_bytes);
    KERNEL_CHECK(c);
}

void geno_wait(snprel_ctx *c) {
    if (c->pending.empty()) return;
    CUDA_CHECK(cudaStreamSynchronize(c->copy_stream));
    {   // span of the copies: first chunk queued -> last chunk arrived
        float ms = 0;
        if (c->copy_ev0 && cudaEventElapsedTime(&ms, c->copy_ev0, c->pending.back().ev) == cudaSuccess) c->last_copy_ms = ms;
    }
    for (auto &p : c->pending) {
        if (!p.consumed) geno_fix_chunk_padding(c, p);
        cudaEventDestroy(p.ev);
    }
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->pending.clear();
}

// GDS dBit2 payload (uncompressed genotype node of a SNP GDS file with sample.order): ONE continuous
// LSB-first 2-bit stream, sample fastest, with no per-row padding -- rows are not byte aligned when
// n_samp is not a multiple of 4.  Repack to the padded device rows.
__global__ void repack_bits_kernel(const uint8_t *__restrict__ stream, int64_t bit0, uint8_t *__restrict__ dst,
                                   int64_t cnt, int64_t n_samp, int64_t row_bytes) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= cnt * row_bytes) return;
    int64_t r = idx / row_bytes, b = idx - r * row_bytes;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        int64_t i = b * 4 + k;
        uint32_t g = 3;
        if (i < n_samp) {
            int64_t pos = bit0 + (r * n_samp + i) * 2;     // even: a genotype never straddles a byte
            g = (stream[pos >> 3] >> (pos & 7)) & 3u;
        }
        out |= g << (2 * k);
    }
    dst[idx] = (uint8_t)out;
}

void geno_push_bitstream(snprel_ctx *c, const uint8_t *host, int64_t first_genotype, int64_t cnt) {
    need_room(c, cnt, "snprel_geno_push_bitstream");
    if (cnt == 0) return;
    if (!host) fail("snprel_geno_push_bitstream: NULL stream");
    if (first_genotype < 0) fail("snprel_geno_push_bitstream: negative offset");
    // stage the covering byte range of a bounded number of rows at a time
    const int64_t max_rows = std::max<int64_t>(1, (int64_t)(256ll << 20) * 4 / c->n_samp);
    for (int64_t done = 0; done < cnt; done += max_rows) {
        const int64_t rows = std::min(max_rows, cnt - done);
        const int64_t g0 = first_genotype + done * c->n_samp;          // first genotype of this chunk
        const int64_t byte0 = g0 / 4, byte1 = (g0 + rows * c->n_samp + 3) / 4;
        c->stage_u8.alloc((size_t)(byte1 - byte0));
        CUDA_CHECK(cudaMemcpyAsync(c->stage_u8.p, host + byte0, (size_t)(byte1 - byte0), cudaMemcpyHostToDevice,
                                   c->stream));
        const int64_t total = rows * c->row_bytes;
        repack_bits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(
            c->stage_u8.p, (g0 - byte0 * 4) * 2, c->geno2b.p + (c->n_snp + done) * c->row_bytes, rows, c->n_samp,
            c->row_bytes);
        KERNEL_CHECK(c);
        CUDA_CHECK(cudaStreamSynchronize(c->stream));   // staging buffer is reused
    }
    c->n_snp += cnt;
    invalidate(c);
}

void geno_synth(snprel_ctx *c, int64_t n_snp, uint64_t seed, double maf_lo, double maf_hi,
                double miss_rate, int64_t snp_start) {
    need_room(c, n_snp, "snprel_geno_synth");
    if (n_snp == 0) return;
    if (!(maf_lo > 0 && maf_hi <= 1.0 && maf_lo <= maf_hi)) fail("snprel_geno_synth: bad MAF range");
    if (!(miss_rate >= 0 && miss_rate < 1)) fail("snprel_geno_synth: bad missing rate");
    double t = miss_rate * 4294967296.0;
    uint32_t thm = t >= 4294967295.0 ? 4294967295u : (uint32_t)t;
    dim3 grid((unsigned)((c->row_bytes + 255) / 256), 1);
    // grid.y is limited to 65535: loop in slabs
    for (int64_t r0 = 0; r0 < n_snp; r0 += 65535) {
        int64_t rows = std::min<int64_t>(65535, n_snp - r0);
        grid.y = (unsigned)rows;
        synth_kernel<<<grid, 256, 0, c->stream>>>(c->geno2b.p + (c->n_snp + r0) * c->row_bytes,
                                                  rows, c->n_samp, c->row_bytes, seed, maf_lo,
                                                  maf_hi, thm, snp_start + r0);
        KERNEL_CHECK(c);
    }
    // a shard of a larger data set (row 0 of this workspace is SNP snp_start - position > 0 of it): the same
    // offset keys the rounding draws of the covariance path (snprel_set_snp_origin)
    if (snp_start - c->n_snp > 0) c->snp_origin = snp_start - c->n_snp;
    c->n_snp += n_snp;
    invalidate(c);
}

// SNP-sharded loading of a workspace that finally holds ALL SNPs (tiled N x N output): a rank
// reserves the whole range, seeks to its own SNP block, pushes it, receives the other blocks straight
// into the device rows (NCCL all-gather / peer copies) and then commits the row count.
void geno_seek(snprel_ctx *c, int64_t snp_index) {
    if (c->n_samp <= 0) fail("snprel_geno_seek: no genotype workspace");
    if (snp_index < 0 || snp_index > c->snp_cap) fail("snprel_geno_seek: position outside the reserved SNP range");
    c->n_snp = snp_index;
    invalidate(c);
}
void geno_commit(snprel_ctx *c, int64_t n_snp) {
    if (c->n_samp <= 0) fail("snprel_geno_commit: no genotype workspace");
    if (n_snp < 0 || n_snp > c->snp_cap) fail("snprel_geno_commit: more SNPs than the reserved range");
    c->n_snp = n_snp;
    invalidate(c);
}

// rows [n_snp, round_up(n_snp, SNP_PAD)) must read as all-missing
void geno_pad_tail(snprel_ctx *c) {
    int64_t end = round_up(std::max<int64_t>(c->n_snp, 1), SNP_PAD);
    if (end > c->snp_cap) fail("internal: SNP padding exceeds capacity");
    if (end > c->n_snp)
        CUDA_CHECK(cudaMemsetAsync(c->geno2b.p + c->n_snp * c->row_bytes, 0xFF,
                                   (size_t)(end - c->n_snp) * c->row_bytes, c->stream));
}

void geno_copy_u8(snprel_ctx *c, uint8_t *out) {
    if (c->n_samp <= 0) fail("snprel_geno_copy_u8: no genotype workspace");
    if (!out) fail("snprel_geno_copy_u8: NULL output");
    const int64_t max_rows = std::max<int64_t>(1, (int64_t)(256ll << 20) / c->n_samp);
    c->stage_u8.alloc((size_t)std::min(std::max<int64_t>(c->n_snp, 1), max_rows) * c->n_samp);
    for (int64_t done = 0; done < c->n_snp; done += max_rows) {
        int64_t rows = std::min(max_rows, c->n_snp - done);
        int64_t total = rows * c->n_samp;
        unpack_u8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(
            c->geno2b.p + done * c->row_bytes, c->stage_u8.p, rows, c->n_samp, c->row_bytes);
        KERNEL_CHECK(c);
        CUDA_CHECK(cudaMemcpyAsync(out + done * c->n_samp, c->stage_u8.p, (size_t)total,
                                   cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
}

void geno_copy_2b(snprel_ctx *c, uint8_t *out, int64_t row_bytes_out) {
    if (c->n_samp <= 0) fail("snprel_geno_copy_2b: no genotype workspace");
    if (!out) fail("snprel_geno_copy_2b: NULL output");
    int64_t need = (c->n_samp + 3) / 4;
    if (row_bytes_out < need) fail("snprel_geno_copy_2b: row_bytes too small");
    if (c->n_snp == 0) return;
    int64_t w = std::min(row_bytes_out, c->row_bytes);
    CUDA_CHECK(cudaMemcpy2DAsync(out, (size_t)row_bytes_out, c->geno2b.p, (size_t)c->row_bytes, (size_t)w,
                                 (size_t)c->n_snp, cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
}

// per-SNP statistics of rows [l0, l0 + rows) only (streamed accumulate, grm.cu)
void snp_stats_range(snprel_ctx *c, int64_t l0, int64_t rows) {
    if (rows <= 0) return;
    const int warps = 8;
    snp_stat_kernel<<<(unsigned)((rows + warps - 1) / warps), warps * 32, 0, c->stream>>>(
        c->geno2b.p + l0 * c->row_bytes, c->stat.p + l0, rows, c->row_bytes, (int)c->n_samp_pad);
    KERNEL_CHECK(c);
}

void ensure_stats(snprel_ctx *c) {
    if (c->n_samp <= 0) fail("no genotype workspace (call snprel_geno_begin first)");
    if (c->stat_valid) return;
    geno_pad_tail(c);
    int64_t rows = round_up(std::max<int64_t>(c->n_snp, 1), SNP_PAD);
    const int warps = 8;
    snp_stat_kernel<<<(unsigned)((rows + warps - 1) / warps), warps * 32, 0, c->stream>>>(
        c->geno2b.p, c->stat.p, rows, c->row_bytes, (int)c->n_samp_pad);
    KERNEL_CHECK(c);
    c->stat_valid = true;
}

static std::vector<SnpStat> stats_to_host(snprel_ctx *c) {
    ensure_stats(c);
    std::vector<SnpStat> h((size_t)c->n_snp);
    if (c->n_snp > 0)
        CUDA_CHECK(cudaMemcpyAsync(h.data(), c->stat.p, h.size() * sizeof(SnpStat),
                                   cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    // padding samples were counted as missing against n_samp_pad; num is exact already
    return h;
}

// Get_AF_MR_perSNP, SNP-major branch (src/dGenGWAS.cpp:535-551)
void snp_ratefreq(snprel_ctx *c, double *af, double *maf, double *mr) {
    std::vector<SnpStat> h = stats_to_host(c);
    const double nan = __builtin_nan("");
    for (int64_t l = 0; l < c->n_snp; l++) {
        int num = h[l].num, sum = h[l].sum;
        double f = num > 0 ? (double)sum / (2 * num) : nan;
        if (af) af[l] = f;
        if (maf) maf[l] = num > 0 ? (f < 1 - f ? f : 1 - f) : nan;
        if (mr) mr[l] = 1 - ((double)num) / (double)c->n_samp;
    }
}

// IBD::Init_EPrIBD_IBS (src/genIBD.cpp:253-338): per-SNP expected P(IBS i | IBD j) terms summed
// in SNP order (the reference's own summation order; an O(M) host loop over the device-computed
// per-SNP counts).  sums = {a00, a01, a02, a11, a12, nValid}, not yet divided by nValid so that
// SNP-sharded ranks can add them.  afreq_in != NULL: use those frequencies, no finite-sample
// correction (src/genIBS.cpp:585-587; the AA/AB/BB counts are zero in that mode).
void ibd_mom_sums(snprel_ctx *c, const double *afreq_in, double *sums, double *afreq_out) {
    if (!sums) fail("snprel_ibd_mom_sums: NULL output");
    std::vector<SnpStat> h = stats_to_host(c);
    const double nan = __builtin_nan("");
    double e00 = 0, e01 = 0, e02 = 0, e11 = 0, e12 = 0;
    int64_t nvalid = 0;
    for (int64_t l = 0; l < c->n_snp; l++) {
        long AA = 0, AB = 0, BB = 0;
        if (!afreq_in) {
            AB = h[l].n1;
            AA = (h[l].sum - h[l].n1) / 2;
            BB = h[l].num - AA - AB;
        }
        long n = 2 * (AA + AB + BB);
        double p = (n > 0) ? ((double)(2 * AA + AB) / n) : nan;
        if (afreq_in) {
            p = afreq_in[l];
            if (std::isfinite(p) && (p < 0 || p > 1)) p = nan;
        }
        if (afreq_out) afreq_out[l] = p;
        double q = 1 - p, Na = (double)n;
        double x = (double)(2 * AA + AB), y = (double)(2 * BB + AB);
        double a00, a01, a02, a11, a12;
        if (!afreq_in) {
            a00 = 2*p*p*q*q * ((x-1)/x * (y-1)/y * (Na/(Na-1)) * (Na/(Na-2)) * (Na/(Na-3)));
            a01 = 4*p*p*p*q * ((x-1)/x * (x-2)/x * (Na/(Na-1)) * (Na/(Na-2)) * (Na/(Na-3))) +
                  4*p*q*q*q * ((y-1)/y * (y-2)/y * (Na/(Na-1)) * (Na/(Na-2)) * (Na/(Na-3)));
            a02 = q*q*q*q * ((y-1)/y * (y-2)/y * (y-3)/y * (Na/(Na-1)) * (Na/(Na-2)) * (Na/(Na-3))) +
                  p*p*p*p * ((x-1)/x * (x-2)/x * (x-3)/x * (Na/(Na-1)) * (Na/(Na-2)) * (Na/(Na-3))) +
                  4*p*p*q*q * ((x-1)/x * (y-1)/y * (Na/(Na-1)) * (Na/(Na-2)) * (Na/(Na-3)));
            a11 = 2*p*p*q * ((x-1)/x * Na/(Na-1) * Na/(Na-2)) +
                  2*p*q*q * ((y-1)/y * Na/(Na-1) * Na/(Na-2));
            a12 = p*p*p * ((x-1)/x * (x-2)/x * Na/(Na-1) * Na/(Na-2)) +
                  q*q*q * ((y-1)/y * (y-2)/y * Na/(Na-1) * Na/(Na-2)) +
                  p*p*q * ((x-1)/x * Na/(Na-1) * Na/(Na-2)) +
                  p*q*q * ((y-1)/y * Na/(Na-1) * Na/(Na-2));
        } else {
            a00 = 2*p*p*q*q;
            a01 = 4*p*p*p*q + 4*p*q*q*q;
            a02 = q*q*q*q + p*p*p*p + 4*p*p*q*q;
            a11 = 2*p*p*q + 2*p*q*q;
            a12 = p*p*p + q*q*q + p*p*q + p*q*q;
        }
        if (std::isfinite(a00) && std::isfinite(a01) && std::isfinite(a02) && std::isfinite(a11) &&
            std::isfinite(a12)) {
            e00 += a00; e01 += a01; e02 += a02; e11 += a11; e12 += a12;
            nvalid++;
        }
    }
    sums[0] = e00; sums[1] = e01; sums[2] = e02; sums[3] = e11; sums[4] = e12;
    sums[5] = (double)nvalid;
}

// Select_SNP_Base (src/dGenGWAS.cpp:361-397); with afreq != NULL Select_SNP_Base_Ex
// (src/dGenGWAS.cpp:399-470): MAF from the caller's frequencies, non-finite = dropped.
void select_snp_base(snprel_ctx *c, const double *afreq, int remove_mono, double maf, double missrate,
                     uint8_t *out_sel, int64_t *n_removed) {
    std::vector<SnpStat> h = stats_to_host(c);
    std::vector<int64_t> keep;
    keep.reserve((size_t)c->n_snp);
    for (int64_t l = 0; l < c->n_snp; l++) {
        int num = h[l].num, sum = h[l].sum;
        bool flag = false;
        if (afreq ? std::isfinite(afreq[l]) : num > 0) {
            double f = afreq ? afreq[l] : (double)sum / (2 * num);
            double m = f < 1 - f ? f : 1 - f;
            double r = 1 - ((double)num) / (double)c->n_samp;
            flag = true;
            if (remove_mono && m <= 0) flag = false;
            if (flag && m < maf) flag = false;
            if (flag && r > missrate) flag = false;
        }
        if (out_sel) out_sel[l] = flag ? 1 : 0;
        if (flag) keep.push_back(l);
    }
    int64_t removed = c->n_snp - (int64_t)keep.size();
    if (n_removed) *n_removed = removed;
    if (removed == 0) return;
    // compact rows into a fresh buffer
    int64_t n_out = (int64_t)keep.size();
    int64_t cap = round_up(std::max<int64_t>(n_out, 1), SNP_PAD);
    DevBuf<uint8_t> fresh;
    fresh.alloc((size_t)cap * c->row_bytes);
    if (n_out > 0) {
        DevBuf<int64_t> idx;
        idx.alloc(keep.size());
        CUDA_CHECK(cudaMemcpyAsync(idx.p, keep.data(), keep.size() * sizeof(int64_t),
                                   cudaMemcpyHostToDevice, c->stream));
        dim3 grid((unsigned)(((c->row_bytes >> 4) + 127) / 128), 1);
        for (int64_t r0 = 0; r0 < n_out; r0 += 65535) {
            int64_t rows = std::min<int64_t>(65535, n_out - r0);
            grid.y = (unsigned)rows;
            gather_rows_kernel<<<grid, 128, 0, c->stream>>>(
                c->geno2b.p, fresh.p + r0 * c->row_bytes, idx.p + r0, rows, c->row_bytes);
            KERNEL_CHECK(c);
        }
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    std::swap(c->geno2b.p, fresh.p);
    std::swap(c->geno2b.n, fresh.n);
    c->snp_cap = cap;
    c->n_snp = n_out;
    c->stat.alloc(cap);
    invalidate(c);
}

void ensure_planes(snprel_ctx *c) {
    if (c->n_samp <= 0) fail("no genotype workspace (call snprel_geno_begin first)");
    if (c->planes_valid) return;
    geno_pad_tail(c);
    int64_t words = round_up(std::max<int64_t>(c->n_snp, 1), SNP_PAD) / 64;
    c->planes.alloc((size_t)words * c->n_samp_pad);
    c->plane_words = words;
    dim3 grid((unsigned)((c->row_bytes + 127) / 128), 1);
    for (int64_t w0 = 0; w0 < words; w0 += 65535) {
        int64_t nw = std::min<int64_t>(65535, words - w0);
        grid.y = (unsigned)nw;
        planes_kernel<<<grid, 128, 0, c->stream>>>(c->geno2b.p + w0 * 64 * c->row_bytes,
                                                   c->planes.p + w0 * c->n_samp_pad, nw,
                                                   c->row_bytes, c->n_samp_pad);
        KERNEL_CHECK(c);
    }
    c->planes_valid = true;
}

}  // namespace snprel
