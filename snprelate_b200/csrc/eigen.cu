// Top-k eigenpairs of the N x N relatedness matrix -- the step after the Gram accumulation in
// snpgdsPCA / snpgdsEIGMIX.  The reference calls LAPACK dspevx on the packed matrix
// (CalcEigen, src/genPCA.cpp:1262-1346); an O(N^3) library step, not the hot path, but at
// N = 10 000 a full cuSOLVER decomposition (1.7 s) costs six times the whole Gram accumulation.
//
// Two solvers, both built from library calls (cuBLAS / cuSOLVER):
//   * dense: cusolverDnXsyevd on -C, keep the first k columns (any n, any k);
//   * Chebyshev-filtered subspace iteration (Zhou, Saad, Tiago, Chelikowsky 2006) for k << n:
//     a block of b = max(2k, k+32) vectors is repeatedly passed through a degree-m Chebyshev
//     polynomial of C that is bounded by 1 on [lambda_min, theta_b] (the unwanted part of the
//     spectrum, bounds from a short Lanczos run and the previous Rayleigh-Ritz step) and grows
//     like cosh(m acosh(.)) above it, re-orthonormalised (Householder QR) and rotated to Ritz
//     vectors; pairs whose residual ||C x - theta x|| is below 1e-11 |theta_1| are locked and
//     deflated until k are locked.
//     Work = (#iterations x m) GEMMs C (n x n) x Y (n x b): ~100-300 block products, 0.1-0.2 s at
//     n = 10 000 even when the wanted eigenvalues sit in a Marchenko-Pastur bulk edge (the
//     structure-free synthetic benchmark; real population structure converges in 2-3 rounds).
//     If the residuals stall the dense solver is used: never a wrong answer, at worst a slow one.
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <algorithm>
#include <chrono>
#include <cmath>

#include "common.cuh"

namespace snprel {

#define CUBLAS_CHECK(expr)                                                                   \
    do {                                                                                     \
        cublasStatus_t _s = (expr);                                                          \
        if (_s != CUBLAS_STATUS_SUCCESS) ::snprel::fail("cuBLAS error %d at %s:%d", (int)_s, __FILE__, __LINE__); \
    } while (0)
#define CUSOLVER_CHECK(expr)                                                                 \
    do {                                                                                     \
        cusolverStatus_t _s = (expr);                                                        \
        if (_s != CUSOLVER_STATUS_SUCCESS) ::snprel::fail("cuSOLVER error %d at %s:%d", (int)_s, __FILE__, __LINE__); \
    } while (0)

__global__ void sym_fill_kernel(const double *__restrict__ src, double *__restrict__ dst, double sign, int64_t n) {
    int64_t i = blockIdx.x, j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= n || j < i) return;
    double v = sign * src[i * n + j];
    dst[i * n + j] = v;
    dst[j * n + i] = v;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__global__ void random_fill_kernel(double *__restrict__ x, int64_t count, uint64_t seed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint64_t h = mix64(seed ^ (uint64_t)i * 0xD1342543DE82EF95ull);
    x[i] = (double)(h >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

// res[j] = || CX[:, j] - theta[j] X[:, j] ||_2
__global__ void residual_kernel(const double *__restrict__ CX, const double *__restrict__ X,
                                const double *__restrict__ theta, int64_t n, double *__restrict__ res) {
    const int j = blockIdx.x;
    const double th = theta[j];
    double s = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        double r = CX[(int64_t)j * n + i] - th * X[(int64_t)j * n + i];
        s += r * r;
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) res[j] = sqrt(sh[0]);
}

// dst = -(h + h^T)/2 (b x b): Rayleigh quotient matrix, negated so that syevd's ascending order is
// the descending order of C's Ritz values
__global__ void neg_sym_kernel(const double *__restrict__ h, double *__restrict__ dst, int b) {
    int i = blockIdx.x, j = threadIdx.x;
    if (i < b && j < b) dst[i * b + j] = -0.5 * (h[i * b + j] + h[j * b + i]);
}

static dim3 tri_grid(int64_t n) { return dim3((unsigned)n, (unsigned)((n + 127) / 128)); }

// eigenvalues / eigenvectors of a small symmetric matrix on the host (cyclic Jacobi)
static void jacobi_eig(std::vector<double> &a, int m, std::vector<double> &w, std::vector<double> &v) {
    v.assign((size_t)m * m, 0.0);
    for (int i = 0; i < m; i++) v[i * m + i] = 1.0;
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < m; p++)
            for (int q = p + 1; q < m; q++) off += a[p * m + q] * a[p * m + q];
        if (off < 1e-300) break;
        for (int p = 0; p < m; p++)
            for (int q = p + 1; q < m; q++) {
                double apq = a[p * m + q];
                if (std::fabs(apq) < 1e-300) continue;
                double th = (a[q * m + q] - a[p * m + p]) / (2 * apq);
                double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1));
                double cs = 1 / std::sqrt(t * t + 1), sn = t * cs;
                for (int r = 0; r < m; r++) {
                    double arp = a[r * m + p], arq = a[r * m + q];
                    a[r * m + p] = cs * arp - sn * arq;
                    a[r * m + q] = sn * arp + cs * arq;
                }
                for (int r = 0; r < m; r++) {
                    double apr = a[p * m + r], aqr = a[q * m + r];
                    a[p * m + r] = cs * apr - sn * aqr;
                    a[q * m + r] = sn * apr + cs * aqr;
                }
                for (int r = 0; r < m; r++) {
                    double vrp = v[r * m + p], vrq = v[r * m + q];
                    v[r * m + p] = cs * vrp - sn * vrq;
                    v[r * m + q] = sn * vrp + cs * vrq;
                }
            }
    }
    w.resize(m);
    for (int i = 0; i < m; i++) w[i] = a[i * m + i];
}

struct Handles {
    cublasHandle_t blas = nullptr;
    cusolverDnHandle_t sol = nullptr;
    cusolverDnParams_t params = nullptr;
    explicit Handles(cudaStream_t s) {
        if (cublasCreate(&blas) != CUBLAS_STATUS_SUCCESS) fail("cublasCreate failed");
        if (cusolverDnCreate(&sol) != CUSOLVER_STATUS_SUCCESS) {
            cublasDestroy(blas);
            fail("cusolverDnCreate failed");
        }
        cublasSetStream(blas, s);
        cusolverDnSetStream(sol, s);
        cusolverDnCreateParams(&params);
    }
    ~Handles() {
        if (params) cusolverDnDestroyParams(params);
        if (sol) cusolverDnDestroy(sol);
        if (blas) cublasDestroy(blas);
    }
};

// library handles live as long as the context (creating them costs more than a small solve)
static Handles &handles(snprel_ctx *c) {
    if (!c->eig_handles) c->eig_handles = new Handles(c->stream);
    return *static_cast<Handles *>(c->eig_handles);
}
void eigen_release(snprel_ctx *c) {
    delete static_cast<Handles *>(c->eig_handles);
    c->eig_handles = nullptr;
}

// dense solver: full decomposition of -C (64-bit API, divide and conquer), first k columns kept
static void dense_topk(snprel_ctx *c, Handles &h, const double *m_upper, int64_t n, int k, double *eigval,
                       double *eigvec) {
    DevBuf<double> a, w;
    a.alloc((size_t)n * n);
    w.alloc((size_t)n);
    sym_fill_kernel<<<tri_grid(n), 128, 0, c->stream>>>(m_upper, a.p, -1.0, n);
    KERNEL_CHECK(c);
    DevBuf<int> info;
    info.alloc(1);
    size_t wdev = 0, whost = 0;
    CUSOLVER_CHECK(cusolverDnXsyevd_bufferSize(h.sol, h.params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER,
                                               (int64_t)n, CUDA_R_64F, a.p, (int64_t)n, CUDA_R_64F, w.p, CUDA_R_64F,
                                               &wdev, &whost));
    DevBuf<uint8_t> work;
    work.alloc(std::max<size_t>(wdev, 1));
    std::vector<uint8_t> hwork(std::max<size_t>(whost, 1));
    cusolverStatus_t st = cusolverDnXsyevd(h.sol, h.params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int64_t)n,
                                           CUDA_R_64F, a.p, (int64_t)n, CUDA_R_64F, w.p, CUDA_R_64F, work.p, wdev,
                                           hwork.data(), whost, info.p);
    int hinfo = 0;
    CUDA_CHECK(cudaMemcpyAsync(&hinfo, info.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (st != CUSOLVER_STATUS_SUCCESS || hinfo != 0)
        fail("eigen-decomposition error (%d), infinite or missing values in the genetic covariance matrix!",
             hinfo);   // wording follows src/genPCA.cpp:1330-1334
    std::vector<double> hw((size_t)n);
    CUDA_CHECK(cudaMemcpyAsync(hw.data(), w.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (eigvec)
        CUDA_CHECK(cudaMemcpyAsync(eigvec, a.p, (size_t)n * k * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (eigval)
        for (int i = 0; i < k; i++) eigval[i] = -hw[i];
}

// [lower, upper] estimates of the spectrum from an m-step Lanczos run, widened by the last
// residual (the usual safeguarded bounds of filtered subspace iteration)
static void lanczos_bounds(snprel_ctx *c, Handles &h, const double *A, int64_t n, double &lo, double &hi) {
    const int m = (int)std::min<int64_t>(16, n);
    DevBuf<double> V, w;
    V.alloc((size_t)n * 2);
    w.alloc((size_t)n);
    double *v0 = V.p, *v1 = V.p + n;
    random_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(v0, n, 0x5eedULL);
    KERNEL_CHECK(c);
    const int ni = (int)n;
    double nrm = 0;
    CUBLAS_CHECK(cublasDnrm2(h.blas, ni, v0, 1, &nrm));
    double inv = 1.0 / nrm;
    CUBLAS_CHECK(cublasDscal(h.blas, ni, &inv, v0, 1));
    std::vector<double> al, be;
    const double one = 1, zero = 0;
    double beta_prev = 0;
    for (int j = 0; j < m; j++) {
        CUBLAS_CHECK(cublasDgemv(h.blas, CUBLAS_OP_N, ni, ni, &one, A, ni, v0, 1, &zero, w.p, 1));
        if (j > 0) {
            double nb = -beta_prev;
            CUBLAS_CHECK(cublasDaxpy(h.blas, ni, &nb, v1, 1, w.p, 1));
        }
        double a = 0;
        CUBLAS_CHECK(cublasDdot(h.blas, ni, w.p, 1, v0, 1, &a));
        double na = -a;
        CUBLAS_CHECK(cublasDaxpy(h.blas, ni, &na, v0, 1, w.p, 1));
        al.push_back(a);
        double b = 0;
        CUBLAS_CHECK(cublasDnrm2(h.blas, ni, w.p, 1, &b));
        be.push_back(b);
        if (!(b > 0) || j == m - 1) break;
        // v1 <- v0, v0 <- w / b
        CUBLAS_CHECK(cublasDcopy(h.blas, ni, v0, 1, v1, 1));
        CUBLAS_CHECK(cublasDcopy(h.blas, ni, w.p, 1, v0, 1));
        double ib = 1.0 / b;
        CUBLAS_CHECK(cublasDscal(h.blas, ni, &ib, v0, 1));
        beta_prev = b;
    }
    const int mm = (int)al.size();
    std::vector<double> T((size_t)mm * mm, 0.0), wv, S;
    for (int i = 0; i < mm; i++) {
        T[i * mm + i] = al[i];
        if (i + 1 < mm) T[i * mm + i + 1] = T[(i + 1) * mm + i] = be[i];
    }
    jacobi_eig(T, mm, wv, S);
    const double blast = be.back();
    lo = 1e300;
    hi = -1e300;
    for (int i = 0; i < mm; i++) {
        double r = std::fabs(blast * S[(mm - 1) * mm + i]);
        lo = std::min(lo, wv[i] - r);
        hi = std::max(hi, wv[i] + r);
    }
}

// Chebyshev-filtered subspace iteration with locking; returns false when it did not converge.
//   * converged leading Ritz pairs are locked (moved to L) and the active block is kept orthogonal
//     to L after EVERY filter step: rounding noise along a locked eigenvector would otherwise be
//     amplified by p(lambda_locked) / p(lambda_active), which is astronomically large when a few
//     population-structure eigenvalues tower over the noise bulk;
//   * for the same reason the degree of a round is limited so that the amplification ratio inside
//     the active block stays below 1e12 (low degrees while dominant eigenvalues are still active).
static bool chfsi_topk(snprel_ctx *c, Handles &h, const double *m_upper, int64_t n, int k, double *eigval,
                       double *eigvec, int *rounds_out, int *gemms_out) {
    const int b = (int)std::min<int64_t>(n, std::max(2 * k, k + 32));
    const int deg_max = 40, max_rounds = 80;
    const double tol = 1e-11;
    const int ni = (int)n;
    DevBuf<double> A, BX, BY, BC, BZ, L, H, Hs, W, res, theta, tau, P;
    A.alloc((size_t)n * n);
    sym_fill_kernel<<<tri_grid(n), 128, 0, c->stream>>>(m_upper, A.p, 1.0, n);
    KERNEL_CHECK(c);
    const size_t nb = (size_t)n * b;
    BX.alloc(nb);
    BY.alloc(nb);
    BC.alloc(nb);
    BZ.alloc(nb);
    L.alloc((size_t)n * k);
    H.alloc((size_t)b * b);
    Hs.alloc((size_t)b * b);
    P.alloc((size_t)b * b);
    W.alloc((size_t)b);
    res.alloc((size_t)b);
    theta.alloc((size_t)b);
    tau.alloc((size_t)b);
    DevBuf<int> info;
    info.alloc(1);
    int lw_qr = 0, lw_org = 0, lw_ev = 0;
    CUSOLVER_CHECK(cusolverDnDgeqrf_bufferSize(h.sol, ni, b, BX.p, ni, &lw_qr));
    CUSOLVER_CHECK(cusolverDnDorgqr_bufferSize(h.sol, ni, b, b, BX.p, ni, tau.p, &lw_org));
    CUSOLVER_CHECK(cusolverDnDsyevd_bufferSize(h.sol, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, b, Hs.p, b, W.p,
                                               &lw_ev));
    DevBuf<double> work;
    work.alloc((size_t)std::max(std::max(lw_qr, lw_org), std::max(lw_ev, 1)) + 1024);
    const int lwork = (int)work.n;
    const double one = 1, zero = 0, mone = -1;
    double gemm_cols = 0;   // columns pushed through C (a block product = b columns)
    // wall time per phase (stream synchronised at the boundaries): filter, QR, Rayleigh-Ritz
    double t_filter = 0, t_qr = 0, t_rr = 0;
    auto now = [&]() {
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    };
    int nl = 0;             // locked pairs
    std::vector<double> lamL;

    auto orthonormalise = [&](double *M, int cols) {
        CUSOLVER_CHECK(cusolverDnDgeqrf(h.sol, ni, cols, M, ni, tau.p, work.p, lwork, info.p));
        CUSOLVER_CHECK(cusolverDnDorgqr(h.sol, ni, cols, cols, M, ni, tau.p, work.p, lwork, info.p));
    };
    // M (n x cols) -= L (L^T M)
    auto project_out_locked = [&](double *M, int cols) {
        if (nl == 0) return;
        CUBLAS_CHECK(cublasDgemm(h.blas, CUBLAS_OP_T, CUBLAS_OP_N, nl, cols, ni, &one, L.p, ni, M, ni, &zero, P.p, nl));
        CUBLAS_CHECK(cublasDgemm(h.blas, CUBLAS_OP_N, CUBLAS_OP_N, ni, cols, nl, &mone, L.p, ni, P.p, nl, &one, M, ni));
    };
    std::vector<double> hth((size_t)b), hres((size_t)b);
    // Rayleigh-Ritz on span(Q) (n x cols): Xo <- Q S, BC <- (C Q) S, hth descending, hres residual norms
    auto rayleigh_ritz = [&](const double *Q, int cols, double *Xo) {
        CUBLAS_CHECK(cublasDgemm(h.blas, CUBLAS_OP_N, CUBLAS_OP_N, ni, cols, ni, &one, A.p, ni, Q, ni, &zero, BZ.p, ni));
        gemm_cols += cols;
        CUBLAS_CHECK(cublasDgemm(h.blas, CUBLAS_OP_T, CUBLAS_OP_N, cols, cols, ni, &one, Q, ni, BZ.p, ni, &zero, H.p, cols));
        neg_sym_kernel<<<cols, cols, 0, c->stream>>>(H.p, Hs.p, cols);
        KERNEL_CHECK(c);
        CUSOLVER_CHECK(cusolverDnDsyevd(h.sol, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, cols, Hs.p, cols, W.p,
                                        work.p, lwork, info.p));
        int hinfo = 0;
        CUDA_CHECK(cudaMemcpyAsync(&hinfo, info.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaMemcpyAsync(hth.data(), W.p, (size_t)cols * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if (hinfo != 0) fail("eigen-decomposition error (%d) in the Rayleigh-Ritz step", hinfo);
        for (int i = 0; i < cols; i++) hth[i] = -hth[i];
        CUDA_CHECK(cudaMemcpyAsync(theta.p, hth.data(), (size_t)cols * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CUBLAS_CHECK(cublasDgemm(h.blas, CUBLAS_OP_N, CUBLAS_OP_N, ni, cols, cols, &one, Q, ni, Hs.p, cols, &zero, Xo, ni));
        CUBLAS_CHECK(cublasDgemm(h.blas, CUBLAS_OP_N, CUBLAS_OP_N, ni, cols, cols, &one, BZ.p, ni, Hs.p, cols, &zero, BC.p, ni));
        const int nres = std::min(cols, k - nl);
        residual_kernel<<<nres, 256, 0, c->stream>>>(BC.p, Xo, theta.p, n, res.p);
        KERNEL_CHECK(c);
        CUDA_CHECK(cudaMemcpyAsync(hres.data(), res.p, (size_t)nres * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
    };

    double lo = 0, hi = 0;
    lanczos_bounds(c, h, A.p, n, lo, hi);
    random_fill_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, c->stream>>>(BY.p, (int64_t)nb, 0xC0FFEEULL);
    KERNEL_CHECK(c);
    orthonormalise(BY.p, b);
    rayleigh_ritz(BY.p, b, BX.p);
    double *Xa = BX.p, *CXa = BC.p, *Yfree = BY.p;   // active Ritz block, C times it, a free n x b buffer
    int ba = b, off = 0;                              // active width; offset of the active pairs in hth / hres
    const double scale = std::max(std::max(std::fabs(hth[0]), std::fabs(hi)), 1e-300);
    double best = 1e300;
    int stalled = 0, round = 0;
    bool ok = false;
    for (;; round++) {
        // lock the converged leading pairs of the active block
        int q = 0;
        while (q < ba && nl + q < k && hres[q] < tol * scale) q++;
        if (q > 0) {
            CUDA_CHECK(cudaMemcpyAsync(L.p + (size_t)nl * n, Xa, (size_t)q * n * sizeof(double), cudaMemcpyDeviceToDevice,
                                       c->stream));
            for (int i = 0; i < q; i++) lamL.push_back(hth[off + i]);
            nl += q;
            Xa += (size_t)q * n;
            CXa += (size_t)q * n;
            ba -= q;
            off += q;
        }
        if (nl >= k) {
            ok = true;
            break;
        }
        const double top = hres[q] / scale;
        if (q > 0 || top < 0.7 * best) stalled = 0;
        else stalled++;
        best = std::min(best, top);
        if (round >= max_rounds || stalled >= 8) break;

        // damp [a, bcut], amplify above; the polynomial is scaled to 1 at the largest active Ritz value
        const double a = std::min(lo, hth[off + ba - 1]), bcut = hth[off + ba - 1], a0 = hth[off];
        const double e = 0.5 * (bcut - a), cc = 0.5 * (bcut + a);
        if (!(e > 0) || !(a0 - cc > e)) break;
        const double xmax = (a0 - cc) / e;
        const int deg = (int)std::max(2.0, std::min((double)deg_max, std::floor(27.6 / std::acosh(xmax))));
        double sig = e / (a0 - cc);
        const double sig1 = sig;
        const int nba = (int)((size_t)n * ba);
        // y = (C X - cc X) * sig1 / e        (C X is at hand from the Rayleigh-Ritz step)
        double *xp = Xa, *y = Yfree;
        const double t0 = now();
        CUDA_CHECK(cudaMemcpyAsync(y, CXa, (size_t)nba * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
        double s0 = sig1 / e, s1 = -cc * sig1 / e;
        CUBLAS_CHECK(cublasDscal(h.blas, nba, &s0, y, 1));
        CUBLAS_CHECK(cublasDaxpy(h.blas, nba, &s1, xp, 1, y, 1));
        project_out_locked(y, ba);
        for (int i = 2; i <= deg; i++) {   // three-term recurrence: (xp, y) -> (y, ynew), ynew overwrites xp
            const double sig2 = 1.0 / (2.0 / sig1 - sig);
            const double alpha = 2.0 * sig2 / e, beta = -sig * sig2, gam = -2.0 * cc * sig2 / e;
            CUBLAS_CHECK(cublasDgemm(h.blas, CUBLAS_OP_N, CUBLAS_OP_N, ni, ba, ni, &alpha, A.p, ni, y, ni, &beta, xp, ni));
            gemm_cols += ba;
            CUBLAS_CHECK(cublasDaxpy(h.blas, nba, &gam, y, 1, xp, 1));
            project_out_locked(xp, ba);
            std::swap(xp, y);
            sig = sig2;
        }
        project_out_locked(y, ba);
        const double t1 = now();
        orthonormalise(y, ba);
        if (nl > 0) {
            // Householder QR can re-introduce components along L at the level eps * cond(block):
            // remove them and orthonormalise once more (the block is now within ~1e-4 of orthonormal).
            // (A Cholesky QR would do, but the first cusolverDnDpotrf / cublasDtrsm call of a process
            // costs ~90 s of lazy kernel loading on this CUDA 12.9 / sm_100 stack; geqrf + orgqr
            // take 1.5 ms per round at n = 10 000.)
            project_out_locked(y, ba);
            orthonormalise(y, ba);
        }
        const double t2 = now();
        rayleigh_ritz(y, ba, xp);          // new Ritz block into the scratch of the recurrence
        const double t3 = now();
        t_filter += t1 - t0;
        t_qr += t2 - t1;
        t_rr += t3 - t2;
        Xa = xp;
        CXa = BC.p;
        Yfree = (xp >= BX.p && xp < BX.p + nb) ? BY.p : BX.p;
        off = 0;
    }
    if (rounds_out) *rounds_out = round;
    if (gemms_out) *gemms_out = (int)std::lround(gemm_cols / b);
    c->eig_phase_ms[0] = t_filter;
    c->eig_phase_ms[1] = t_qr;
    c->eig_phase_ms[2] = t_rr;
    if (!ok) return false;
    // descending order (pairs are locked roughly, not strictly, in that order)
    std::vector<int> order((size_t)k);
    for (int i = 0; i < k; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y2) { return lamL[x] > lamL[y2]; });
    for (int i = 0; i < k; i++) {
        if (eigval) eigval[i] = lamL[order[i]];
        if (eigvec)
            CUDA_CHECK(cudaMemcpyAsync(eigvec + (size_t)i * n, L.p + (size_t)order[i] * n, (size_t)n * sizeof(double),
                                       cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return true;
}

// top-k eigenpairs of the symmetric matrix `m_upper` (n x n, upper triangle valid), descending;
// eigval[n]: k values then NaN (src/genPCA.cpp:1339-1341); eigvec: n x k column-major
void eigen_topk(snprel_ctx *c, const double *m_upper, int64_t n, int k, double *eigval, double *eigvec) {
    if (k <= 0) return;
    if (k > n) k = (int)n;
    if (n > 2147483647ll / 130) fail("eigen step: too many samples for the dense eigen solvers");
    Handles &h = handles(c);
    c->eig_rounds = 0;
    c->eig_gemms = 0;
    bool done = false;
    // (block width b = max(2k, k+32) <= 512: the small Rayleigh-Ritz helpers use one thread per column)
    const bool try_filter = n >= 2048 && (int64_t)k * 8 <= n && std::max(2 * k, k + 32) <= 512 && !(c->debug_flags & 4u);
    if (try_filter) done = chfsi_topk(c, h, m_upper, n, k, eigval, eigvec, &c->eig_rounds, &c->eig_gemms);
    c->eig_solver = done ? 1 : 0;
    if (!done) dense_topk(c, h, m_upper, n, k, eigval, eigvec);
    if (eigval) {
        const double nan = __builtin_nan("");
        for (int64_t i = k; i < n; i++) eigval[i] = nan;
    }
}

// ---------------------------------------------------------------------------
// dense helpers for the randomized PCA (project.cu): library calls on small / tall-skinny operands
// ---------------------------------------------------------------------------
void la_transpose(snprel_ctx *c, int64_t m, int64_t n, const double *A, int64_t lda, double *B, int64_t ldb) {
    Handles &h = handles(c);
    const double one = 1.0, zero = 0.0;
    CUBLAS_CHECK(cublasDgeam(h.blas, CUBLAS_OP_T, CUBLAS_OP_N, (int)m, (int)n, &one, A, (int)lda, &zero, B, (int)ldb, B, (int)ldb));
}

void la_orthonormalise(snprel_ctx *c, double *A, int64_t m, int n) {
    Handles &h = handles(c);
    if (m > 2147483647ll) fail("la_orthonormalise: too many rows");
    DevBuf<double> tau, work;
    DevBuf<int> info;
    tau.alloc((size_t)n);
    info.alloc(1);
    int lw_qr = 0, lw_org = 0;
    CUSOLVER_CHECK(cusolverDnDgeqrf_bufferSize(h.sol, (int)m, n, A, (int)m, &lw_qr));
    CUSOLVER_CHECK(cusolverDnDorgqr_bufferSize(h.sol, (int)m, n, n, A, (int)m, tau.p, &lw_org));
    const int lwork = std::max(lw_qr, lw_org);
    work.alloc((size_t)lwork);
    CUSOLVER_CHECK(cusolverDnDgeqrf(h.sol, (int)m, n, A, (int)m, tau.p, work.p, lwork, info.p));
    CUSOLVER_CHECK(cusolverDnDorgqr(h.sol, (int)m, n, n, A, (int)m, tau.p, work.p, lwork, info.p));
    int hinfo = 0;
    CUDA_CHECK(cudaMemcpyAsync(&hinfo, info.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (hinfo != 0) fail("cusolverDnDgeqrf / Dorgqr failed (info %d)", hinfo);
}

void la_svd_tall(snprel_ctx *c, double *A, int64_t m, int n, double *S_host, double *U, double *VT) {
    Handles &h = handles(c);
    if (m < n) fail("la_svd_tall: needs m >= n");
    DevBuf<double> S, work;
    DevBuf<int> info;
    S.alloc((size_t)n);
    info.alloc(1);
    int lwork = 0;
    CUSOLVER_CHECK(cusolverDnDgesvd_bufferSize(h.sol, (int)m, n, &lwork));
    work.alloc((size_t)std::max(lwork, 1));
    CUSOLVER_CHECK(cusolverDnDgesvd(h.sol, U ? 'S' : 'N', VT ? 'S' : 'N', (int)m, n, A, (int)m, S.p, U, (int)m, VT, n,
                                    work.p, lwork, nullptr, info.p));
    int hinfo = 0;
    CUDA_CHECK(cudaMemcpyAsync(&hinfo, info.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaMemcpyAsync(S_host, S.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (hinfo != 0) fail("cusolverDnDgesvd failed (info %d)", hinfo);
}

}  // namespace snprel
