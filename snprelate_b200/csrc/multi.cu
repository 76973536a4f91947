// Several GPUs of one box behind ONE handle and ONE host process: what an R
// session needs to use all eight B200s through the .Call entry points (r_shim.cpp).
//
//   * SNP-block sharding: device r owns the r-th contiguous SNP range; pushes are routed by SNP index.
//   * accumulate: every device plans and accumulates its shard on its own host thread (the library's
//     calls block on their stream), with ONE fixed-point format agreed from the merged plan statistics
//     -- the same algebra as snprelate_b200/dist.py (reduce_plan / accumulate_sharded), no process group.
//   * reduction over NVLink / NVSwitch peer memory, hand written (no NCCL in this path): every device
//     sums ITS row slice of every partial plane straight out of its peers' HBM (reduce-scatter by
//     peer loads), then the devices that finish the result pull the reduced slices (all devices for
//     the small per-sample buffers).  Only the live upper-triangle columns of the N x N planes cross
//     the links.  Partials are exact integers, so the sum is bit-identical for any device count.
//   * finishing calls (snprel_grm, snprel_pca, snprel_ibs_num, ...) then run on the root device's
//     context, which holds the global accumulators (snprel_multi_ctx).
// Two entries of `devices` may name the same GPU (tests on a one-GPU box run the whole path that way).
#include <condition_variable>
#include <mutex>
#include <thread>

#include "common.cuh"

#include <cuda.h>

using namespace snprel;

struct snprel_multi {
    std::vector<snprel_ctx *> ctx;
    std::vector<int> dev;
    int64_t n_samp = 0, cap = 0, per_dev = 0, pos = 0;
    int active = 0;          // devices that own a non-empty SNP range
    std::string err;
    bool replicated = false; // every device reserves the whole SNP range (tiled N x N output)
    bool gathered = false;
    double reduce_ms = 0;    // last peer reduction (CUDA events on the root's stream + host barrier)
    int64_t reduce_bytes = 0;   // bytes that crossed the links in it
};

namespace {

constexpr int MAX_DEV = 16;

template <class T>
struct PeerSrc {
    const T *p[MAX_DEV];
    int n;
};

// dst[idx] (+)= sum_k src.p[k][idx] over rows [ra, rb) of every plane; live columns only.
// 16-byte peer loads, four independent vectors per thread in flight (the loads cross NVLink: latency,
// not issue, is what has to be covered); ld and the first live column are multiples of 256 elements, so
// every row segment is 16-byte aligned.
template <class T>
struct Vec16 {
    static constexpr int N = 16 / sizeof(T);
    T v[N];
};
template <class T>
__device__ __forceinline__ Vec16<T> ld16(const T *p) {
    Vec16<T> r;
    const uint4 u = *reinterpret_cast<const uint4 *>(p);
    memcpy(r.v, &u, 16);
    return r;
}
template <class T>
__device__ __forceinline__ void st16(T *p, const Vec16<T> &r) {
    uint4 u;
    memcpy(&u, r.v, 16);
    *reinterpret_cast<uint4 *>(p) = u;
}

template <class T, bool ASSIGN>
__global__ void __launch_bounds__(256)
peer_reduce_kernel(T *__restrict__ dst, const PeerSrc<T> src, int64_t count, int64_t ld, int64_t rows, int64_t row0,
                   int tri, int64_t ra, int64_t rb, int planes) {
    constexpr int V = Vec16<T>::N, U = 4;
    const int64_t r = ra + blockIdx.x;
    if (r >= rb) return;
    const int64_t c0 = tri ? ((row0 + r) & ~255ll) : 0;
    for (int p = blockIdx.y; p < planes; p += gridDim.y) {
        const int64_t base = ((int64_t)p * rows + r) * ld;
        const int64_t end = min(ld, count - base);            // elements of this row that exist
        const int64_t nvec = end > c0 ? (end - c0) / V : 0;
        for (int64_t i0 = (int64_t)threadIdx.x; i0 < nvec; i0 += (int64_t)blockDim.x * U) {
            Vec16<T> acc[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int64_t i = i0 + (int64_t)u * blockDim.x;
                if (i < nvec) {
                    if (ASSIGN) {
#pragma unroll
                        for (int e = 0; e < V; e++) acc[u].v[e] = T(0);
                    } else {
                        acc[u] = ld16(dst + base + c0 + i * V);
                    }
                }
            }
            for (int k = 0; k < src.n; k++) {
                Vec16<T> x[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int64_t i = i0 + (int64_t)u * blockDim.x;
                    if (i < nvec) x[u] = ld16(src.p[k] + base + c0 + i * V);
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const int64_t i = i0 + (int64_t)u * blockDim.x;
                    if (i < nvec) {
#pragma unroll
                        for (int e = 0; e < V; e++) acc[u].v[e] += x[u].v[e];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int64_t i = i0 + (int64_t)u * blockDim.x;
                if (i < nvec) st16(dst + base + c0 + i * V, acc[u]);
            }
        }
        // scalar tail of a dense buffer whose length is not a multiple of the vector
        for (int64_t c = c0 + nvec * V + threadIdx.x; c < end; c += blockDim.x) {
            T v = ASSIGN ? T(0) : dst[base + c];
            for (int k = 0; k < src.n; k++) v += src.p[k][base + c];
            dst[base + c] = v;
        }
    }
}

struct Shape {
    int64_t ld, rows, row0;
    int planes, tri;
};
// a dense buffer is walked as rows of 4096 elements
Shape shape_of(const ReduceBuf &b) {
    if (b.ld > 0 && b.rows > 0) return Shape{b.ld, b.rows, b.row0, (int)(b.count / (b.ld * b.rows)), 1};
    const int64_t ld = 4096;
    return Shape{ld, (b.count + ld - 1) / ld, 0, 1, 0};
}
int64_t live_elems(const Shape &s, int64_t ra, int64_t rb) {
    int64_t t = 0;
    for (int64_t r = ra; r < rb; r++) t += s.ld - (s.tri ? ((s.row0 + r) & ~255ll) : 0);
    return t * s.planes;
}

template <class T>
void launch(bool assign, cudaStream_t st, void *dst, const std::vector<const void *> &srcs, const ReduceBuf &b,
            const Shape &s, int64_t ra, int64_t rb) {
    if (rb <= ra || srcs.empty()) return;
    PeerSrc<T> ps;
    ps.n = (int)srcs.size();
    for (int k = 0; k < ps.n; k++) ps.p[k] = static_cast<const T *>(srcs[k]);
    dim3 grid((unsigned)(rb - ra), (unsigned)std::min(s.planes, 8));
    if (assign)
        peer_reduce_kernel<T, true><<<grid, 256, 0, st>>>(static_cast<T *>(dst), ps, b.count, s.ld, s.rows, s.row0, s.tri, ra, rb, s.planes);
    else
        peer_reduce_kernel<T, false><<<grid, 256, 0, st>>>(static_cast<T *>(dst), ps, b.count, s.ld, s.rows, s.row0, s.tri, ra, rb, s.planes);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) fail("peer_reduce_kernel: %s", cudaGetErrorString(e));
}
void launch_kind(int kind, bool assign, cudaStream_t st, void *dst, const std::vector<const void *> &srcs,
                 const ReduceBuf &b, const Shape &s, int64_t ra, int64_t rb) {
    if (kind == 0) launch<long long>(assign, st, dst, srcs, b, s, ra, rb);
    else if (kind == 1) launch<unsigned int>(assign, st, dst, srcs, b, s, ra, rb);
    else launch<double>(assign, st, dst, srcs, b, s, ra, rb);
}

void set_dev(int d) {
    cudaError_t e = cudaSetDevice(d);
    if (e != cudaSuccess) fail("cudaSetDevice(%d): %s", d, cudaGetErrorString(e));
}
void sync_all(snprel_multi *m) {
    for (int i = 0; i < m->active; i++) {
        set_dev(m->dev[i]);
        CUDA_CHECK(cudaStreamSynchronize(m->ctx[i]->stream));
    }
}

// run fn(i) for every active device on its own host thread; the first error wins
template <class F>
void on_each(snprel_multi *m, F fn) {
    std::vector<std::string> errs((size_t)m->active);
    std::vector<std::thread> th;
    for (int i = 0; i < m->active; i++)
        th.emplace_back([&, i]() {
            try {
                fn(i);
            } catch (const Error &e) {
                errs[i] = e.msg.empty() ? "error" : e.msg;
            } catch (const std::exception &e) {
                errs[i] = e.what();
            }
        });
    for (auto &t : th) t.join();
    for (int i = 0; i < m->active; i++)
        if (!errs[i].empty()) fail("device %d: %s", m->dev[i], errs[i].c_str());
}
void ck(snprel_multi *m, int i, int rc) {
    if (rc != 0) fail("%s", snprel_last_error(m->ctx[i]));
}

// in-place sum of reduce buffer k of every active context; the result lands on device `root`
// (root < 0: on every device).  Small buffers always go everywhere: each device's later epilogues need them.
void peer_reduce(snprel_multi *m, int root) {
    const int nd = m->active;
    if (nd <= 1) return;
    const size_t nb = m->ctx[0]->reduce_list.size();
    for (int i = 1; i < nd; i++)
        if (m->ctx[i]->reduce_list.size() != nb) fail("internal: the devices disagree on the reduce buffers");
    cudaEvent_t e0, e1;
    set_dev(m->dev[0]);
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    CUDA_CHECK(cudaEventRecord(e0, m->ctx[0]->stream));
    m->reduce_bytes = 0;
    // phase 1: device d sums row slice d of every buffer out of its peers' memory
    std::vector<Shape> shapes(nb);
    for (size_t k = 0; k < nb; k++) {
        const ReduceBuf &b0 = m->ctx[0]->reduce_list[k];
        for (int i = 1; i < nd; i++)
            if (m->ctx[i]->reduce_list[k].count != b0.count || m->ctx[i]->reduce_list[k].kind != b0.kind)
                fail("internal: reduce buffer %d differs between devices", (int)k);
        shapes[k] = shape_of(b0);
    }
    const int esz[3] = {8, 4, 8};
    for (int d = 0; d < nd; d++) {
        set_dev(m->dev[d]);
        for (size_t k = 0; k < nb; k++) {
            const ReduceBuf &b = m->ctx[d]->reduce_list[k];
            if (b.count <= 0) continue;
            const Shape &s = shapes[k];
            const int64_t ra = s.rows * d / nd, rb = s.rows * (d + 1) / nd;
            std::vector<const void *> srcs;
            for (int p = 0; p < nd; p++)
                if (p != d) srcs.push_back(m->ctx[p]->reduce_list[k].ptr);
            launch_kind(b.kind, false, m->ctx[d]->stream, b.ptr, srcs, b, s, ra, rb);
            m->reduce_bytes += live_elems(s, ra, rb) * esz[b.kind] * (nd - 1);
        }
    }
    sync_all(m);
    // phase 2: the finishing device(s) pull the reduced slices of the others
    for (int t = 0; t < nd; t++) {
        set_dev(m->dev[t]);
        for (size_t k = 0; k < nb; k++) {
            const ReduceBuf &b = m->ctx[t]->reduce_list[k];
            if (b.count <= 0) continue;
            const bool everywhere = root < 0 || b.count <= (1 << 22);
            if (!everywhere && t != root) continue;
            const Shape &s = shapes[k];
            for (int d = 0; d < nd; d++) {
                if (d == t) continue;
                const int64_t ra = s.rows * d / nd, rb = s.rows * (d + 1) / nd;
                std::vector<const void *> srcs{m->ctx[d]->reduce_list[k].ptr};
                launch_kind(b.kind, true, m->ctx[t]->stream, b.ptr, srcs, b, s, ra, rb);
                m->reduce_bytes += live_elems(s, ra, rb) * esz[b.kind];
            }
        }
    }
    sync_all(m);
    set_dev(m->dev[0]);
    CUDA_CHECK(cudaEventRecord(e1, m->ctx[0]->stream));
    CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    m->reduce_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
}


// ---------------------------------------------------------------------------
// The same reduction for ONE PROCESS PER GPU (torchrun / torch.distributed, bench.py --gpus N): the peers'
// buffers are mapped through CUDA IPC handles instead of being plain pointers of this process.  The
// host side (snprelate_b200/dist.py:peer_reduce_buffers) exchanges the handles once per buffer
// allocation and places a barrier between the phases; the kernels are the ones above.
// ---------------------------------------------------------------------------
struct IpcPeers {
    struct Map { cudaIpcMemHandle_t h; void *base; };
    std::vector<Map> maps;                       // opened allocations (cached: the buffers are persistent)
    std::vector<std::vector<const void *>> ptr;  // [buffer][rank] (own rank: the local pointer)
    int world = 0, rank = 0;
    ~IpcPeers() {
        for (auto &m : maps) cudaIpcCloseMemHandle(m.base);
    }
    void *open(const cudaIpcMemHandle_t &h) {
        for (auto &m : maps)
            if (memcmp(&m.h, &h, sizeof(h)) == 0) return m.base;
        void *base = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) fail("cudaIpcOpenMemHandle: %s (peer-memory reduction needs NVLink / PCIe peer access between the ranks' GPUs)", cudaGetErrorString(e));
        maps.push_back({h, base});
        return base;
    }
};

// route `cnt` SNP rows starting at global position m->pos to their owners
template <class F>
void route(snprel_multi *m, int64_t cnt, const char *who, F push) {
    if (m->n_samp <= 0) fail("%s: call snprel_multi_geno_begin first", who);
    if (cnt < 0 || m->pos + cnt > m->cap) fail("%s: more SNPs than the capacity given to snprel_multi_geno_begin", who);
    int64_t done = 0;
    while (done < cnt) {
        const int64_t g = m->pos + done;
        const int i = (int)(g / m->per_dev);
        const int64_t take = std::min(cnt - done, (int64_t)(i + 1) * m->per_dev - g);
        push(i, done, take);
        done += take;
    }
    m->pos += cnt;
}

}  // namespace

#define MULTI_BEGIN(m)   \
    if (!(m)) return 1;  \
    try {
#define MULTI_END(m)                 \
    return 0;                        \
    }                                \
    catch (const Error &e) {         \
        (m)->err = e.msg;            \
        return 1;                    \
    }                                \
    catch (const std::exception &e) {\
        (m)->err = e.what();         \
        return 2;                    \
    }

static std::string g_multi_create_error;

extern "C" {

int snprel_multi_create(const int *devices, int n_dev, snprel_multi **out) {
    if (!out || !devices || n_dev <= 0 || n_dev > MAX_DEV) {
        g_multi_create_error = "snprel_multi_create: bad arguments (1..16 devices)";
        return 1;
    }
    *out = nullptr;
    snprel_multi *m = new snprel_multi();
    for (int i = 0; i < n_dev; i++) {
        snprel_ctx *c = nullptr;
        if (snprel_create(&c, devices[i]) != 0) {
            g_multi_create_error = snprel_last_error(nullptr);
            for (auto *x : m->ctx) snprel_destroy(x);
            delete m;
            return 1;
        }
        m->ctx.push_back(c);
        m->dev.push_back(devices[i]);
    }
    // peer access between every pair of distinct devices: the reduction reads peers' HBM directly
    for (int i = 0; i < n_dev; i++)
        for (int j = 0; j < n_dev; j++) {
            if (devices[i] == devices[j]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[i], devices[j]);
            cudaSetDevice(devices[i]);
            cudaError_t e = can ? cudaDeviceEnablePeerAccess(devices[j], 0) : cudaErrorPeerAccessUnsupported;
            if (e == cudaErrorPeerAccessAlreadyEnabled) {
                cudaGetLastError();
                e = cudaSuccess;
            }
            if (e != cudaSuccess) {
                g_multi_create_error = "snprel_multi_create: no peer access from device " + std::to_string(devices[i]) +
                                       " to device " + std::to_string(devices[j]) + " (" + cudaGetErrorString(e) + ")";
                for (auto *x : m->ctx) snprel_destroy(x);
                delete m;
                return 1;
            }
        }
    m->active = n_dev;
    *out = m;
    return 0;
}

void snprel_multi_destroy(snprel_multi *m) {
    if (!m) return;
    for (auto *c : m->ctx) snprel_destroy(c);
    delete m;
}

const char *snprel_multi_last_error(snprel_multi *m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }

int snprel_multi_device_count(snprel_multi *m) { return m ? (int)m->ctx.size() : 0; }

snprel_ctx *snprel_multi_ctx(snprel_multi *m, int i) {
    return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr;
}

int snprel_multi_geno_begin(snprel_multi *m, int64_t n_samp, int64_t snp_capacity) {
    MULTI_BEGIN(m)
    if (snp_capacity < 0) fail("snprel_multi_geno_begin: negative SNP capacity");
    const int nd = (int)m->ctx.size();
    // contiguous SNP ranges, boundaries at multiples of the kernels' 128-SNP stage; never more
    // devices than 128-SNP blocks
    const int64_t blocks = std::max<int64_t>(1, (snp_capacity + SNP_PAD - 1) / SNP_PAD);
    m->active = (int)std::min<int64_t>(nd, blocks);
    m->per_dev = (blocks + m->active - 1) / m->active * SNP_PAD;
    m->active = (int)std::min<int64_t>(m->active, (std::max<int64_t>(snp_capacity, 1) + m->per_dev - 1) / m->per_dev);
    m->n_samp = n_samp;
    m->cap = snp_capacity;
    m->pos = 0;
    m->replicated = false;
    m->gathered = false;
    for (int i = 0; i < m->active; i++) {
        const int64_t lo = i * m->per_dev, hi = std::min(snp_capacity, lo + m->per_dev);
        ck(m, i, snprel_geno_begin(m->ctx[i], n_samp, std::max<int64_t>(hi - lo, 0)));
        ck(m, i, snprel_set_snp_origin(m->ctx[i], lo));     // keys the rounding draws by the global SNP index
    }
    MULTI_END(m)
}

int snprel_multi_geno_begin_replicated(snprel_multi *m, int64_t n_samp, int64_t snp_capacity) {
    MULTI_BEGIN(m)
    if (snp_capacity < 0) fail("snprel_multi_geno_begin_replicated: negative SNP capacity");
    const int nd = (int)m->ctx.size();
    const int64_t blocks = std::max<int64_t>(1, (snp_capacity + SNP_PAD - 1) / SNP_PAD);
    m->active = nd;                                   // every device computes windows, even with an empty SNP block
    m->per_dev = (blocks + nd - 1) / nd * SNP_PAD;
    m->n_samp = n_samp;
    m->cap = snp_capacity;
    m->pos = 0;
    m->replicated = true;
    m->gathered = false;
    for (int i = 0; i < nd; i++) {
        ck(m, i, snprel_geno_begin(m->ctx[i], n_samp, snp_capacity));
        ck(m, i, snprel_geno_seek(m->ctx[i], std::min(snp_capacity, i * m->per_dev)));
    }
    MULTI_END(m)
}

// every device pulls the other devices' SNP blocks out of their HBM (cudaMemcpyPeerAsync over NVLink)
int snprel_multi_geno_gather(snprel_multi *m) {
    MULTI_BEGIN(m)
    if (!m->replicated) fail("snprel_multi_geno_gather: call snprel_multi_geno_begin_replicated first");
    if (m->pos != m->cap) fail("snprel_multi_geno_gather: %lld of %lld SNPs pushed", (long long)m->pos, (long long)m->cap);
    const int nd = (int)m->ctx.size();
    sync_all(m);
    for (int t = 0; t < nd; t++) {
        set_dev(m->dev[t]);
        for (int s = 0; s < nd; s++) {
            if (s == t) continue;
            const int64_t lo = std::min(m->cap, s * m->per_dev), hi = std::min(m->cap, lo + m->per_dev);
            if (hi <= lo) continue;
            const size_t off = (size_t)lo * m->ctx[t]->row_bytes, bytes = (size_t)(hi - lo) * m->ctx[t]->row_bytes;
            if (m->dev[s] == m->dev[t])
                CUDA_CHECK(cudaMemcpyAsync(m->ctx[t]->geno2b.p + off, m->ctx[s]->geno2b.p + off, bytes, cudaMemcpyDeviceToDevice,
                                           m->ctx[t]->stream));
            else
                CUDA_CHECK(cudaMemcpyPeerAsync(m->ctx[t]->geno2b.p + off, m->dev[t], m->ctx[s]->geno2b.p + off, m->dev[s], bytes,
                                               m->ctx[t]->stream));
        }
    }
    sync_all(m);
    for (int i = 0; i < nd; i++) ck(m, i, snprel_geno_commit(m->ctx[i], m->cap));
    m->gathered = true;
    MULTI_END(m)
}

int snprel_multi_grm_tiled(snprel_multi *m, int method, int64_t window_rows, snprel_sink_fn sink, void *user) {
    MULTI_BEGIN(m)
    if (!m->replicated || !m->gathered) fail("snprel_multi_grm_tiled: needs snprel_multi_geno_begin_replicated + pushes + snprel_multi_geno_gather");
    if (method != SNPREL_GRM_EIGENSTRAT && method != SNPREL_GRM_GCTA && method != SNPREL_GRM_EIGMIX)
        fail("snprel_multi_grm_tiled: methods Eigenstrat, GCTA and EIGMIX only");
    if (!sink) fail("snprel_multi_grm_tiled: NULL sink");
    const int nd = (int)m->ctx.size();
    const int64_t n = m->n_samp, npad = round_up(n, SAMP_PAD);
    if (window_rows < 0 || window_rows % 256) fail("snprel_multi_grm_tiled: window_rows must be a non-negative multiple of 256");
    if (window_rows == 0) {
        // two int64 planes + the float64 result per window row, inside 60 % of the smallest free memory
        int64_t free_min = INT64_MAX;
        for (int i = 0; i < nd; i++) {
            int64_t f = 0, t = 0;
            ck(m, i, snprel_mem_info(m->ctx[i], &f, &t));
            free_min = std::min(free_min, f);
        }
        const int64_t per_row = 3 * 8 * npad;
        window_rows = std::max<int64_t>(256, std::min<int64_t>(npad, (int64_t)(0.6 * (double)free_min / (double)per_row) / 256 * 256));
        // at least two windows per device keep all devices busy
        const int64_t balanced = round_up((npad + 2 * nd - 1) / (2 * nd), 256);
        if (nd > 1) window_rows = std::max<int64_t>(256, std::min(window_rows, balanced));
    }
    const int64_t nw = (n + window_rows - 1) / window_rows;
    std::mutex mu;
    std::condition_variable cv;
    int64_t next = 0;        // next window the sink expects
    bool abort = false;
    std::string sink_err;
    const int saved_active = m->active;
    m->active = nd;
    try {
        on_each(m, [&](int d) {
            snprel_ctx *c = m->ctx[d];
            std::vector<double> host;
            for (int64_t w = d; w < nw; w += nd) {
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (abort) return;
                }
                const int64_t r0 = w * window_rows;
                ck(m, d, snprel_set_row_window(c, r0, std::min(window_rows, npad - r0)));
                int64_t cnt = 0;
                ck(m, d, snprel_window_count(c, &cnt));
                host.resize((size_t)cnt);
                int rc = snprel_grm(c, method, host.data(), 1, nullptr);
                std::unique_lock<std::mutex> lk(mu);
                if (rc != 0) {
                    abort = true;
                    cv.notify_all();
                    lk.unlock();
                    ck(m, d, rc);
                }
                cv.wait(lk, [&] { return next == w || abort; });
                if (abort) return;
                const int64_t first = r0 * n - r0 * (r0 - 1) / 2;      // packed index of (r0, r0)
                if (sink(user, first, host.data(), cnt) != 0) {
                    abort = true;
                    sink_err = "snprel_multi_grm_tiled: the sink asked to stop";
                }
                next = w + 1;
                cv.notify_all();
            }
        });
    } catch (...) {
        m->active = saved_active;
        for (int i = 0; i < nd; i++) snprel_set_row_window(m->ctx[i], 0, 0);
        throw;
    }
    m->active = saved_active;
    for (int i = 0; i < nd; i++) ck(m, i, snprel_set_row_window(m->ctx[i], 0, 0));
    if (!sink_err.empty()) fail("%s", sink_err.c_str());
    MULTI_END(m)
}

int snprel_multi_geno_push_u8(snprel_multi *m, const uint8_t *geno, int64_t cnt) {
    MULTI_BEGIN(m)
    m->gathered = false;
    route(m, cnt, "snprel_multi_geno_push_u8", [&](int i, int64_t off, int64_t take) {
        ck(m, i, snprel_geno_push_u8(m->ctx[i], geno + off * m->n_samp, take));
    });
    MULTI_END(m)
}

int snprel_multi_geno_push_2b(snprel_multi *m, const uint8_t *packed, int64_t cnt, int64_t row_bytes) {
    MULTI_BEGIN(m)
    m->gathered = false;
    route(m, cnt, "snprel_multi_geno_push_2b", [&](int i, int64_t off, int64_t take) {
        ck(m, i, snprel_geno_push_2b(m->ctx[i], packed + off * row_bytes, take, row_bytes));
    });
    MULTI_END(m)
}

int snprel_multi_geno_synth(snprel_multi *m, int64_t n_snp, uint64_t seed, double maf_lo, double maf_hi, double miss_rate,
                            int64_t snp_start) {
    MULTI_BEGIN(m)
    m->gathered = false;
    route(m, n_snp, "snprel_multi_geno_synth", [&](int i, int64_t off, int64_t take) {
        ck(m, i, snprel_geno_synth(m->ctx[i], take, seed, maf_lo, maf_hi, miss_rate, snp_start + (m->pos + off)));
    });
    MULTI_END(m)
}

int snprel_multi_set_row_window(snprel_multi *m, int64_t row0, int64_t rows) {
    MULTI_BEGIN(m)
    for (int i = 0; i < m->active; i++) ck(m, i, snprel_set_row_window(m->ctx[i], row0, rows));
    MULTI_END(m)
}

int snprel_multi_set_count_engine(snprel_multi *m, int engine) {
    MULTI_BEGIN(m)
    for (size_t i = 0; i < m->ctx.size(); i++) ck(m, (int)i, snprel_set_count_engine(m->ctx[i], engine));
    MULTI_END(m)
}

int snprel_multi_set_rounding(snprel_multi *m, int mode) {
    MULTI_BEGIN(m)
    for (size_t i = 0; i < m->ctx.size(); i++) ck(m, (int)i, snprel_set_rounding(m->ctx[i], mode));
    MULTI_END(m)
}

int snprel_multi_accumulate(snprel_multi *m, int est, int bayesian, int root) {
    MULTI_BEGIN(m)
    const int nd = m->active;
    if (m->n_samp <= 0) fail("snprel_multi_accumulate: no genotype workspace");
    if (m->replicated) fail("snprel_multi_accumulate: the workspace is replicated (tiled output): use snprel_multi_grm_tiled");
    if (root >= nd) root = 0;      // (fewer active devices than asked for: tiny data sets)
    const bool cov = est >= SNPREL_GRM_EIGENSTRAT && est <= SNPREL_GRM_EIGMIX;
    if (!cov && est != SNPREL_EST_IBS && est != SNPREL_EST_KING_ROBUST && est != SNPREL_EST_BETA)
        fail("snprel_multi_accumulate: unsupported estimator %d", est);
    std::vector<snprel_plan> plans((size_t)nd);
    for (auto &p : plans) {
        p = snprel_plan{};
        p.frac_bits = p.frac_bits_w = p.frac_bits_d = -1;
        p.bayesian = bayesian;
    }
    on_each(m, [&](int i) { ck(m, i, snprel_plan_local(m->ctx[i], est, &plans[i])); });
    // one fixed-point format for every device: max of the table ranges, sums of the rest (sums are
    // conservative for the per-sample maxima); snprelate_b200/dist.py:reduce_plan
    snprel_plan g = plans[0];
    for (int i = 1; i < nd; i++) {
        g.max_abs = std::max(g.max_abs, plans[i].max_abs);
        g.max_abs_w = std::max(g.max_abs_w, plans[i].max_abs_w);
        g.sum_bound += plans[i].sum_bound;
        g.err_weight += plans[i].err_weight;
        g.scale += plans[i].scale;
        g.total_missing += plans[i].total_missing;
        g.max_missing += plans[i].max_missing;
        g.n_snp += plans[i].n_snp;
        // (a sum of per-device maxima bounds the maximum of the sums; 0 = some device did not measure it)
        g.diag_bound = (g.diag_bound > 0 && plans[i].diag_bound > 0) ? g.diag_bound + plans[i].diag_bound : 0.0;
        g.sum_rest += plans[i].sum_rest;
        g.err_weight2 += plans[i].err_weight2;
    }
    on_each(m, [&](int i) { ck(m, i, snprel_accumulate(m->ctx[i], est, &g)); });
    peer_reduce(m, root);
    for (int i = 0; i < nd; i++) ck(m, i, snprel_mark_reduced(m->ctx[i]));
    MULTI_END(m)
}

int snprel_multi_last_reduce(snprel_multi *m, double *ms, int64_t *bytes) {
    MULTI_BEGIN(m)
    if (ms) *ms = m->reduce_ms;
    if (bytes) *bytes = m->reduce_bytes;
    MULTI_END(m)
}

// ---- one process per GPU: peer-memory reduction through CUDA IPC -----------------------------------
int snprel_reduce_ipc_export(snprel_ctx *c, int idx, void *handle64, int64_t *offset) {
    if (!c) return 1;
    try {
        set_dev(c->device);
        if (idx < 0 || idx >= (int)c->reduce_list.size()) fail("snprel_reduce_ipc_export: index out of range");
        if (!handle64 || !offset) fail("snprel_reduce_ipc_export: NULL output");
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
        void *p = c->reduce_list[idx].ptr;
        memset(handle64, 0, 64);
        *offset = 0;
        if (p && c->reduce_list[idx].count > 0) {
            // the handle names the whole allocation; the buffer may start inside it
            CUdeviceptr base = 0;
            size_t size = 0;
            cudaPointerAttributes at;
            CUDA_CHECK(cudaPointerGetAttributes(&at, p));
            typedef CUresult (*RangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);
            static RangeFn range = nullptr;
            if (!range) {
                void *fn = nullptr;
                cudaDriverEntryPointQueryResult q;
                CUDA_CHECK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
                if (!fn || q != cudaDriverEntryPointSuccess) fail("cuMemGetAddressRange is not available");
                range = reinterpret_cast<RangeFn>(fn);
            }
            if (range(&base, &size, (CUdeviceptr)p) != CUDA_SUCCESS) fail("cuMemGetAddressRange failed");
            cudaIpcMemHandle_t h;
            CUDA_CHECK(cudaIpcGetMemHandle(&h, (void *)base));
            memcpy(handle64, &h, 64);
            *offset = (int64_t)((CUdeviceptr)p - base);
        }
        return 0;
    } catch (const Error &e) {
        c->err = e.msg;
        return 1;
    }
}

// handles / offsets: [world][n_buffers] (64 bytes each / int64 each), as gathered from every rank
int snprel_peer_reduce_open(snprel_ctx *c, int rank, int world, const void *handles, const int64_t *offsets) {
    if (!c) return 1;
    try {
        set_dev(c->device);
        if (world < 1 || world > MAX_DEV || rank < 0 || rank >= world) fail("snprel_peer_reduce_open: bad rank / world");
        if (!c->ipc_peers) c->ipc_peers = new IpcPeers();
        IpcPeers *ip = static_cast<IpcPeers *>(c->ipc_peers);
        const size_t nb = c->reduce_list.size();
        ip->world = world;
        ip->rank = rank;
        ip->ptr.assign(nb, std::vector<const void *>((size_t)world, nullptr));
        const unsigned char *hb = static_cast<const unsigned char *>(handles);
        // mappings of allocations a peer has since replaced are closed before anything new is opened
        for (size_t i = 0; i < ip->maps.size();) {
            bool wanted = false;
            for (size_t k = 0; k < nb && !wanted; k++)
                for (int r = 0; r < world && !wanted; r++)
                    wanted = r != rank && c->reduce_list[k].count > 0 &&
                             memcmp(&ip->maps[i].h, hb + ((size_t)r * nb + k) * 64, 64) == 0;
            if (wanted) {
                i++;
            } else {
                cudaIpcCloseMemHandle(ip->maps[i].base);
                ip->maps.erase(ip->maps.begin() + (long)i);
            }
        }
        for (size_t k = 0; k < nb; k++)
            for (int r = 0; r < world; r++) {
                if (r == rank) {
                    ip->ptr[k][r] = c->reduce_list[k].ptr;
                    continue;
                }
                if (c->reduce_list[k].count <= 0) continue;
                cudaIpcMemHandle_t h;
                memcpy(&h, hb + ((size_t)r * nb + k) * 64, 64);
                ip->ptr[k][r] = static_cast<const char *>(ip->open(h)) + offsets[(size_t)r * nb + k];
            }
        return 0;
    } catch (const Error &e) {
        c->err = e.msg;
        return 1;
    }
}

// phase 1: sum this rank's row slice of every buffer out of the peers' memory (in place).
// phase 2: pull the other ranks' reduced slices (every rank when root < 0, else the root only; small
// buffers always).  The caller puts a cross-process barrier before phase 1, between the phases and after.
int snprel_peer_reduce_phase(snprel_ctx *c, int phase, int root, int64_t *link_bytes) {
    if (!c) return 1;
    try {
        set_dev(c->device);
        IpcPeers *ip = static_cast<IpcPeers *>(c->ipc_peers);
        if (!ip || ip->ptr.size() != c->reduce_list.size()) fail("snprel_peer_reduce_phase: call snprel_peer_reduce_open first");
        const int nd = ip->world, me = ip->rank;
        const int esz[3] = {8, 4, 8};
        int64_t bytes = 0;
        for (size_t k = 0; k < c->reduce_list.size(); k++) {
            const ReduceBuf &b = c->reduce_list[k];
            if (b.count <= 0) continue;
            const Shape s = shape_of(b);
            if (phase == 1) {
                const int64_t ra = s.rows * me / nd, rb = s.rows * (me + 1) / nd;
                std::vector<const void *> srcs;
                for (int p = 0; p < nd; p++)
                    if (p != me) srcs.push_back(ip->ptr[k][p]);
                launch_kind(b.kind, false, c->stream, b.ptr, srcs, b, s, ra, rb);
                bytes += live_elems(s, ra, rb) * esz[b.kind] * (nd - 1);
            } else {
                const bool everywhere = root < 0 || b.count <= (1 << 22);
                if (!everywhere && me != root) continue;
                for (int d = 0; d < nd; d++) {
                    if (d == me) continue;
                    const int64_t ra = s.rows * d / nd, rb = s.rows * (d + 1) / nd;
                    std::vector<const void *> srcs{ip->ptr[k][d]};
                    launch_kind(b.kind, true, c->stream, b.ptr, srcs, b, s, ra, rb);
                    bytes += live_elems(s, ra, rb) * esz[b.kind];
                }
            }
        }
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (link_bytes) *link_bytes = bytes;
        return 0;
    } catch (const Error &e) {
        c->err = e.msg;
        return 1;
    }
}

void snprel_peer_reduce_close(snprel_ctx *c) {
    if (!c || !c->ipc_peers) return;
    cudaSetDevice(c->device);
    delete static_cast<IpcPeers *>(c->ipc_peers);
    c->ipc_peers = nullptr;
}

}  // extern "C"
