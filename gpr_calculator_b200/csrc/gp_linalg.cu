// gp_linalg.cu — the O(N^2) / O(N^3) GP algebra that follows the covariance build, kept on device.
//
// Replaces the scipy/numpy steps of GP.log_marginal_likelihood, GP.fit, GP.set_K_inv and
// GP.predict* (gaussianprocess.py:128-202, 286-317, 338-377, 880-908).  The factorisation and
// triangular solves are cuSOLVER (potrf / potrs / potri) and the K* K^-1 product is cuBLAS DGEMM:
// library calls, reported separately by bench.py and not optimised (north_star).  The
// bandwidth-bound pieces (noise, log-det, trace of (alpha alpha^T - K^-1) dK, fused mean/variance
// row reductions) are hand-written kernels with fixed-order reductions.
//
// All matrices are row-major; a symmetric row-major matrix is handed to the column-major libraries
// unchanged, "lower" here == CUBLAS_FILL_MODE_UPPER there.
#include "common.cuh"
#include <cstdint>
#include <cusolverDn.h>
#include <cublas_v2.h>
#include <cstdlib>

namespace {

cusolverDnHandle_t g_solver = nullptr;
cublasHandle_t g_blas = nullptr;

int handles(cudaStream_t st) {
    { int rc0 = gprb_pool_init(); if (rc0) return rc0; }
    if (!g_solver) {
        if (cusolverDnCreate(&g_solver) != CUSOLVER_STATUS_SUCCESS) { gprb_set_error("cusolverDnCreate failed"); return GPRB_ERR_CUDA; }
    }
    if (!g_blas) {
        if (cublasCreate(&g_blas) != CUBLAS_STATUS_SUCCESS) { gprb_set_error("cublasCreate failed"); return GPRB_ERR_CUDA; }
    }
    if (cusolverDnSetStream(g_solver, st) != CUSOLVER_STATUS_SUCCESS || cublasSetStream(g_blas, st) != CUBLAS_STATUS_SUCCESS) {
        gprb_set_error("setting the library stream failed");
        return GPRB_ERR_CUDA;
    }
    return GPRB_OK;
}

// Which triangle holds the Cholesky factor.  Default: CUBLAS_FILL_MODE_UPPER on the column-major view (the factor L in the
// row-major lower triangle).  GPRB_POTRF_LOWER=1 (experimental, opt-in: cuSOLVER's potrf is 16 % faster in that mode at
// N = 32 980, tools/potrf_compare.py; not yet run on a GPU) stores L^T in the row-major upper triangle instead; every
// consumer of the factor below takes the mode from here, and GP.L_ reads the same switch.
bool factor_lower() {
    static const bool lower = getenv("GPRB_POTRF_LOWER") != nullptr;
    return lower;
}
cublasFillMode_t factor_uplo() { return factor_lower() ? CUBLAS_FILL_MODE_LOWER : CUBLAS_FILL_MODE_UPPER; }
// the two triangular solves of (factor factor^T) X = B, first and second
cublasOperation_t solve_op(int step) {
    return factor_lower() ? (step == 0 ? CUBLAS_OP_N : CUBLAS_OP_T) : (step == 0 ? CUBLAS_OP_T : CUBLAS_OP_N);
}

__global__ void add_noise_kernel(double *K, long long ld, int N, int NE, double ne2, double nf2) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) K[(long long)i * ld + i] += (i < NE) ? ne2 : nf2;
}

// A[i][j] = A[j][i] for i > j (fill the lower triangle from the upper one)
__global__ void mirror_upper_kernel(double *A, long long ld, int N) {
    __shared__ double t[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;       // source block (bi, bj) with bj >= bi
    if (bj < bi) return;
    const int i = bi * 32 + threadIdx.y, j = bj * 32 + threadIdx.x;
    if (i < N && j < N) t[threadIdx.y][threadIdx.x] = A[(long long)i * ld + j];
    __syncthreads();
    const int ti = bj * 32 + threadIdx.y, tj = bi * 32 + threadIdx.x;   // destination (transposed block)
    if (ti < N && tj < N && ti > tj) A[(long long)ti * ld + tj] = t[threadIdx.x][threadIdx.y];
}

// after potri on the (row-major) lower triangle: copy it to the upper triangle
__global__ void mirror_lower_kernel(double *A, long long ld, int N) {
    __shared__ double t[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj > bi) return;
    const int i = bi * 32 + threadIdx.y, j = bj * 32 + threadIdx.x;
    if (i < N && j < N) t[threadIdx.y][threadIdx.x] = A[(long long)i * ld + j];
    __syncthreads();
    const int ti = bj * 32 + threadIdx.y, tj = bi * 32 + threadIdx.x;   // transposed block
    if (ti < N && tj < N && tj > ti) A[(long long)ti * ld + tj] = t[threadIdx.x][threadIdx.y];
}

__device__ __forceinline__ double block_sum(double v, double *sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0) for (int i = 0; i < nw; i++) r += sh[i];
    return r;   // valid in thread 0
}

__global__ void lml_terms_kernel(const double *L, long long ld, int N, const double *y, const double *alpha, double *out) {
    __shared__ double sh[32];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        a += log(L[(long long)i * ld + i]);
        b += y[i] * alpha[i];
    }
    a = block_sum(a, sh);
    b = block_sum(b, sh);
    if (threadIdx.x == 0) { out[0] = a; out[1] = b; }
}

// partial[block] = sum over the block's rows of sum_j (alpha_i alpha_j - Kinv_ij) * dK_ij
// partial2[block] = sum over rows of (alpha_i^2 - Kinv_ii) * w_i
// KinvE != NULL (sharded inverse, gprb_lml_grad_trace_rows): the columns j < NE of a force row come from the
// energy rows of the inverse, Kinv[i, j] = KinvE[j * ldE + i]; Kinv then only has to be valid for j >= i.
__global__ void __launch_bounds__(256) trace_kernel(int N, int r0, int r1, const double *__restrict__ alpha,
                                                    const double *__restrict__ Kinv, long long ldi,
                                                    const double *__restrict__ dK, long long lddk,
                                                    int NE, double we, double wf, int upper_only, double *partial,
                                                    const double *__restrict__ KinvE = nullptr, long long ldE = 0) {
    __shared__ double sh[32];
    double acc = 0.0, acc2 = 0.0;
    for (int i = r0 + blockIdx.x; i < r1; i += gridDim.x) {
        const double ai = alpha[i];
        const double *ki = Kinv + (long long)i * ldi;
        if (dK) {
            const double *di = dK + (long long)(i - r0) * lddk;
            if (upper_only) {
                // upper_only == 2 (row-sharded build): energy rows hold K_ee only, force rows hold K_fe and the
                // J >= I blocks of K_ff:  tr = EE (upper, doubled) + 2 FE + FF (upper, doubled)
                const int jend = (upper_only == 2 && i < NE) ? NE : N;
                for (int j = i + threadIdx.x; j < jend; j += blockDim.x) {
                    const double t = fma(ai, alpha[j], -ki[j]) * di[j];
                    acc += (j == i) ? t : 2.0 * t;
                }
                if (upper_only == 2 && i >= NE) {
                    if (KinvE)
                        for (int j = threadIdx.x; j < NE; j += blockDim.x) acc += 2.0 * fma(ai, alpha[j], -KinvE[(long long)j * ldE + i]) * di[j];
                    else
                        for (int j = threadIdx.x; j < NE; j += blockDim.x) acc += 2.0 * fma(ai, alpha[j], -ki[j]) * di[j];
                }
            } else {
                for (int j = threadIdx.x; j < N; j += blockDim.x) acc = fma(fma(ai, alpha[j], -ki[j]), di[j], acc);
            }
        }
        if (threadIdx.x == 0) acc2 += (ai * ai - ki[i]) * (i < NE ? we : wf);
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = acc; partial[2 * blockIdx.x + 1] = acc2; }
}

__global__ void block_sum_kernel(int N, int r0, int r1, int c0, int c1, const double *__restrict__ alpha,
                                 const double *__restrict__ Kinv, long long ldi, double *partial) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int i = r0 + blockIdx.x; i < r1; i += gridDim.x) {
        const double ai = alpha[i];
        const double *ki = Kinv + (long long)i * ldi;
        for (int j = c0 + threadIdx.x; j < c1; j += blockDim.x) acc += fma(ai, alpha[j], -ki[j]);
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) { partial[2 * blockIdx.x] = acc; partial[2 * blockIdx.x + 1] = 0.0; }
}

__global__ void final_sum_kernel(const double *partial, int n, double *out) {
    // one thread, fixed order: deterministic
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < n; i++) { a += partial[2 * i]; b += partial[2 * i + 1]; }
        out[0] = 0.5 * a; out[1] = 0.5 * b;
    }
}

// one CTA per test row: mean = Ks[i,:].alpha ; var = max(diag - Ks[i,:].W[i,:], 0)
__global__ void __launch_bounds__(256) predict_rows_kernel(int N, const double *__restrict__ Ks, long long ldks,
                                                           const double *__restrict__ alpha, const double *__restrict__ W,
                                                           const double *__restrict__ diag, double *mean, double *var) {
    __shared__ double sh[32];
    const int i = blockIdx.x;
    const double *k = Ks + (long long)i * ldks;
    double m = 0.0, v = 0.0;
    if (W) {
        const double *w = W + (long long)i * N;
        for (int j = threadIdx.x; j < N; j += blockDim.x) { const double kj = k[j]; m = fma(kj, alpha[j], m); v = fma(kj, w[j], v); }
    } else {
        for (int j = threadIdx.x; j < N; j += blockDim.x) m = fma(k[j], alpha[j], m);
    }
    m = block_sum(m, sh);
    v = block_sum(v, sh);
    if (threadIdx.x == 0) {
        mean[i] = m;
        if (var) { const double r = diag[i] - v; var[i] = r < 0.0 ? 0.0 : r; }   // gaussianprocess.py:906-907
    }
}

int copy_scalars(double *host, const double *dev, int n, cudaStream_t st) {
    GPRB_CUDA(cudaMemcpyAsync(host, dev, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    GPRB_CUDA(cudaStreamSynchronize(st));
    return GPRB_OK;
}

}  // namespace

extern "C" int gprb_add_noise(double *K, long long ldk, int N, int NE, double noise_e, double noise_f, void *stream) {
    GPRB_REQUIRE(K && N >= 0, "gprb_add_noise: bad argument");
    if (N == 0) return GPRB_OK;
    add_noise_kernel<<<(N + 255) / 256, 256, 0, (cudaStream_t)stream>>>(K, ldk, N, NE, noise_e * noise_e, noise_f * noise_f);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    return GPRB_OK;
}

extern "C" int gprb_chol_factor(double *K, long long ldk, int N, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GPRB_REQUIRE(K && N > 0, "gprb_chol_factor: bad argument");
    int rc = handles(st);
    if (rc) return rc;
    int hinfo = -1;
    int *info = nullptr;
    GPRB_CUDA(cudaMallocAsync((void **)&info, sizeof(int), st));
    cusolverStatus_t cs;
    if (getenv("GPRB_POTRF_LEGACY") == nullptr) {
        // generic 64-bit interface (same speed as cusolverDnDpotrf at N = 32 980 on B200: 445 ms; the library's LOWER
        // fill mode takes 374 ms, tools/potrf_compare.py -- a switch of the stored triangle left for the next round)
        static cusolverDnParams_t params = nullptr;
        if (!params && cusolverDnCreateParams(&params) != CUSOLVER_STATUS_SUCCESS) {
            gprb_set_error("cusolverDnCreateParams failed"); return GPRB_ERR_CUDA;
        }
        size_t wdev = 0, whost = 0;
        if (cusolverDnXpotrf_bufferSize(g_solver, params, factor_uplo(), (int64_t)N, CUDA_R_64F, K, (int64_t)ldk,
                                        CUDA_R_64F, &wdev, &whost) != CUSOLVER_STATUS_SUCCESS) {
            gprb_set_error("Xpotrf_bufferSize failed"); return GPRB_ERR_CUDA;
        }
        void *dwork = nullptr, *hwork = nullptr;
        GPRB_CUDA(cudaMallocAsync(&dwork, wdev > 0 ? wdev : 8, st));
        if (whost > 0) hwork = malloc(whost);
        cs = cusolverDnXpotrf(g_solver, params, factor_uplo(), (int64_t)N, CUDA_R_64F, K, (int64_t)ldk, CUDA_R_64F,
                              dwork, wdev, hwork, whost, info);
        GPRB_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, st));
        GPRB_CUDA(cudaFreeAsync(dwork, st));
        GPRB_CUDA(cudaFreeAsync(info, st));
        GPRB_CUDA(cudaStreamSynchronize(st));
        free(hwork);
    } else {
        int lwork = 0;
        if (cusolverDnDpotrf_bufferSize(g_solver, factor_uplo(), N, K, (int)ldk, &lwork) != CUSOLVER_STATUS_SUCCESS) {
            gprb_set_error("potrf_bufferSize failed"); return GPRB_ERR_CUDA;
        }
        double *work = nullptr;
        GPRB_CUDA(cudaMallocAsync((void **)&work, (size_t)(lwork > 0 ? lwork : 1) * sizeof(double), st));
        cs = cusolverDnDpotrf(g_solver, factor_uplo(), N, K, (int)ldk, work, lwork, info);
        GPRB_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, st));
        GPRB_CUDA(cudaFreeAsync(work, st));
        GPRB_CUDA(cudaFreeAsync(info, st));
        GPRB_CUDA(cudaStreamSynchronize(st));
    }
    if (cs != CUSOLVER_STATUS_SUCCESS) { gprb_set_error("cusolver potrf status %d", (int)cs); return GPRB_ERR_CUDA; }
    if (hinfo != 0) { gprb_set_error("matrix not positive definite (potrf info = %d)", hinfo); return GPRB_ERR_LINALG; }
    return GPRB_OK;
}

extern "C" int gprb_chol_solve_vec(const double *L, long long ldl, int N, double *b, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GPRB_REQUIRE(L && b && N > 0, "gprb_chol_solve_vec: bad argument");
    int rc = handles(st);
    if (rc) return rc;
    int *info = nullptr;
    GPRB_CUDA(cudaMallocAsync((void **)&info, sizeof(int), st));
    cusolverStatus_t cs = cusolverDnDpotrs(g_solver, factor_uplo(), N, 1, L, (int)ldl, b, N, info);
    GPRB_CUDA(cudaFreeAsync(info, st));
    if (cs != CUSOLVER_STATUS_SUCCESS) { gprb_set_error("cusolverDnDpotrs status %d", (int)cs); return GPRB_ERR_CUDA; }
    return GPRB_OK;
}

__global__ void set_identity_kernel(double *A, long long ld, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) A[(long long)i * ld + i] = 1.0;
}

extern "C" int gprb_chol_inverse(const double *L, long long ldl, int N, double *Kinv, long long ldi, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GPRB_REQUIRE(L && Kinv && N > 0, "gprb_chol_inverse: bad argument");
    int rc = handles(st);
    if (rc) return rc;
    if ((long long)N * N >= (1LL << 31) || getenv("GPRB_FORCE_TRSM") != nullptr) {   // env: test hook for the large-N route
        // cuSOLVER's potri (and the 64-bit trtri) reject N^2 >= 2^31 (N > 46340, e.g. the S4 configuration).
        // Same result the way gaussianprocess.py:195 gets it, cho_solve(L, I): two triangular solves with
        // the 64-bit cuBLAS interface on an identity right-hand side held in the output buffer.
        dim3 grid((N + 31) / 32, (N + 31) / 32), block(32, 32);
        GPRB_CUDA(cudaMemset2DAsync(Kinv, ldi * sizeof(double), 0, (size_t)N * sizeof(double), N, st));
        set_identity_kernel<<<(N + 255) / 256, 256, 0, st>>>(Kinv, ldi, N);
        GPRB_LAUNCHED();
        const double one = 1.0;
        // column-major view: K = U^T U with U in the upper triangle of L's buffer;  U^T Y = I, then U X = Y
        cublasStatus_t b1 = cublasDtrsm_64(g_blas, CUBLAS_SIDE_LEFT, factor_uplo(), solve_op(0), CUBLAS_DIAG_NON_UNIT,
                                           (int64_t)N, (int64_t)N, &one, L, (int64_t)ldl, Kinv, (int64_t)ldi);
        cublasStatus_t b2 = cublasDtrsm_64(g_blas, CUBLAS_SIDE_LEFT, factor_uplo(), solve_op(1), CUBLAS_DIAG_NON_UNIT,
                                           (int64_t)N, (int64_t)N, &one, L, (int64_t)ldl, Kinv, (int64_t)ldi);
        if (b1 != CUBLAS_STATUS_SUCCESS || b2 != CUBLAS_STATUS_SUCCESS) {
            gprb_set_error("cublasDtrsm_64 status %d / %d", (int)b1, (int)b2); return GPRB_ERR_CUDA;
        }
        mirror_lower_kernel<<<grid, block, 0, st>>>(Kinv, ldi, N);     // exactly symmetric, like the potri route
        GPRB_LAUNCHED();
        GPRB_CUDA(cudaGetLastError());
        return GPRB_OK;
    }
    GPRB_CUDA(cudaMemcpy2DAsync(Kinv, ldi * sizeof(double), L, ldl * sizeof(double), (size_t)N * sizeof(double), N,
                                cudaMemcpyDeviceToDevice, st));
    int lwork = 0;
    if (cusolverDnDpotri_bufferSize(g_solver, factor_uplo(), N, Kinv, (int)ldi, &lwork) != CUSOLVER_STATUS_SUCCESS) {
        gprb_set_error("potri_bufferSize failed"); return GPRB_ERR_CUDA;
    }
    double *work = nullptr; int *info = nullptr;
    GPRB_CUDA(cudaMallocAsync((void **)&work, (size_t)(lwork > 0 ? lwork : 1) * sizeof(double), st));
    GPRB_CUDA(cudaMallocAsync((void **)&info, sizeof(int), st));
    cusolverStatus_t cs = cusolverDnDpotri(g_solver, factor_uplo(), N, Kinv, (int)ldi, work, lwork, info);
    GPRB_CUDA(cudaFreeAsync(work, st));
    GPRB_CUDA(cudaFreeAsync(info, st));
    if (cs != CUSOLVER_STATUS_SUCCESS) { gprb_set_error("cusolverDnDpotri status %d", (int)cs); return GPRB_ERR_CUDA; }
    dim3 grid((N + 31) / 32, (N + 31) / 32), block(32, 32);
    if (factor_lower()) mirror_upper_kernel<<<grid, block, 0, st>>>(Kinv, ldi, N);      // potri filled the row-major upper triangle
    else mirror_lower_kernel<<<grid, block, 0, st>>>(Kinv, ldi, N);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    return GPRB_OK;
}

extern "C" int gprb_lml_terms(const double *L, long long ldl, int N, const double *y, const double *alpha,
                              double *out_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GPRB_REQUIRE(L && y && alpha && out_host && N > 0, "gprb_lml_terms: bad argument");
    double *d = nullptr;
    { int rc0 = gprb_pool_init(); if (rc0) return rc0; }
    GPRB_CUDA(cudaMallocAsync((void **)&d, 2 * sizeof(double), st));
    lml_terms_kernel<<<1, 1024, 0, st>>>(L, ldl, N, y, alpha, d);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    int rc = copy_scalars(out_host, d, 2, st);
    cudaFreeAsync(d, st);
    return rc;
}

extern "C" int gprb_lml_grad_trace(int N, int r0, int r1, const double *alpha, const double *Kinv, long long ldi,
                                   const double *dK_rows, long long lddk, int NE, double we, double wf,
                                   int upper_only, double *out_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GPRB_REQUIRE(alpha && Kinv && out_host && 0 <= r0 && r0 <= r1 && r1 <= N, "gprb_lml_grad_trace: bad argument");
    out_host[0] = out_host[1] = 0.0;
    if (r0 == r1) return GPRB_OK;
    const int blocks = (r1 - r0) < 1184 ? (r1 - r0) : 1184;   // 8 x 148
    double *d = nullptr;
    { int rc0 = gprb_pool_init(); if (rc0) return rc0; }
    GPRB_CUDA(cudaMallocAsync((void **)&d, (size_t)(2 * blocks + 2) * sizeof(double), st));
    trace_kernel<<<blocks, 256, 0, st>>>(N, r0, r1, alpha, Kinv, ldi, dK_rows, lddk, NE, we, wf, upper_only, d);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    final_sum_kernel<<<1, 32, 0, st>>>(d, blocks, d + 2 * blocks);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    int rc = copy_scalars(out_host, d + 2 * blocks, 2, st);
    cudaFreeAsync(d, st);
    return rc;
}

extern "C" int gprb_lml_grad_trace_rows(int N, int r0, int r1, const double *alpha, const double *Kinv_rows, long long ldr,
                                        int c0, const double *KinvE, long long ldE, const double *dK_rows, long long lddk,
                                        int NE, double we, double wf, double *out_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GPRB_REQUIRE(alpha && Kinv_rows && out_host && 0 <= r0 && r0 <= r1 && r1 <= N && 0 <= c0 && c0 <= r0,
                 "gprb_lml_grad_trace_rows: bad argument");
    GPRB_REQUIRE(KinvE || r1 <= NE || NE == 0 || c0 == 0, "gprb_lml_grad_trace_rows: force rows need the energy rows of the inverse");
    out_host[0] = out_host[1] = 0.0;
    if (r0 == r1) return GPRB_OK;
    // the kernel indexes Kinv[i * ld + j] with global (i, j): shift the base so that (r0, c0) is element 0 of the slab
    const double *virt = reinterpret_cast<const double *>(reinterpret_cast<uintptr_t>(Kinv_rows) -
                                                          (uintptr_t)(((long long)r0 * ldr + c0) * (long long)sizeof(double)));
    const int blocks = (r1 - r0) < 1184 ? (r1 - r0) : 1184;
    double *d = nullptr;
    { int rc0 = gprb_pool_init(); if (rc0) return rc0; }
    GPRB_CUDA(cudaMallocAsync((void **)&d, (size_t)(2 * blocks + 2) * sizeof(double), st));
    trace_kernel<<<blocks, 256, 0, st>>>(N, r0, r1, alpha, virt, ldr, dK_rows, lddk, NE, we, wf, 2, d, KinvE, ldE);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    final_sum_kernel<<<1, 32, 0, st>>>(d, blocks, d + 2 * blocks);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    int rc = copy_scalars(out_host, d + 2 * blocks, 2, st);
    cudaFreeAsync(d, st);
    return rc;
}

__global__ void unit_columns_kernel(double *B, long long ldb, int n_rows, int col0) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n_rows) B[(long long)k * ldb + col0 + k] = 1.0;
}

extern "C" int gprb_chol_inverse_rows(const double *L, long long ldl, int N, int r0, int r1, int c0,
                                      double *out, long long ldo, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GPRB_REQUIRE(L && out && N > 0 && 0 <= c0 && c0 <= r0 && r0 <= r1 && r1 <= N && ldo >= N - c0,
                 "gprb_chol_inverse_rows: bad argument");
    if (r0 == r1) return GPRB_OK;
    int rc = handles(st);
    if (rc) return rc;
    // K^-1[T, T] = (L_TT L_TT^T)^-1 for the trailing index set T = [c0, N) (L^-1 is triangular), so rows
    // [r0, r1) of the inverse, restricted to the columns >= c0, are the solution of the trailing system with
    // the unit vectors of those rows as right-hand sides (symmetric: row = column).
    const int n = N - c0, nrhs = r1 - r0;
    GPRB_CUDA(cudaMemset2DAsync(out, ldo * sizeof(double), 0, (size_t)n * sizeof(double), nrhs, st));
    unit_columns_kernel<<<(nrhs + 255) / 256, 256, 0, st>>>(out, ldo, nrhs, r0 - c0);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    // what potrs does, through the 64-bit cuBLAS interface (N^2 may exceed 2^31, e.g. the S4 configuration):
    // column-major view, K_T = U^T U with U in the upper triangle of the factor's buffer;  U^T Y = E, then U X = Y
    const double one = 1.0;
    const double *U = L + (long long)c0 * ldl + c0;
    cublasStatus_t b1 = cublasDtrsm_64(g_blas, CUBLAS_SIDE_LEFT, factor_uplo(), solve_op(0), CUBLAS_DIAG_NON_UNIT,
                                       (int64_t)n, (int64_t)nrhs, &one, U, (int64_t)ldl, out, (int64_t)ldo);
    cublasStatus_t b2 = cublasDtrsm_64(g_blas, CUBLAS_SIDE_LEFT, factor_uplo(), solve_op(1), CUBLAS_DIAG_NON_UNIT,
                                       (int64_t)n, (int64_t)nrhs, &one, U, (int64_t)ldl, out, (int64_t)ldo);
    if (b1 != CUBLAS_STATUS_SUCCESS || b2 != CUBLAS_STATUS_SUCCESS) {
        gprb_set_error("cublasDtrsm_64 (inverse rows) status %d / %d", (int)b1, (int)b2); return GPRB_ERR_CUDA;
    }
    return GPRB_OK;
}

extern "C" int gprb_w_block_sum(int N, int r0, int r1, int c0, int c1, const double *alpha, const double *Kinv,
                                long long ldi, double *out_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GPRB_REQUIRE(alpha && Kinv && out_host && 0 <= r0 && r0 <= r1 && r1 <= N && 0 <= c0 && c0 <= c1 && c1 <= N,
                 "gprb_w_block_sum: bad argument");
    out_host[0] = 0.0;
    if (r0 == r1 || c0 == c1) return GPRB_OK;
    const int blocks = (r1 - r0) < 1184 ? (r1 - r0) : 1184;
    double *d = nullptr;
    { int rc0 = gprb_pool_init(); if (rc0) return rc0; }
    GPRB_CUDA(cudaMallocAsync((void **)&d, (size_t)(2 * blocks + 2) * sizeof(double), st));
    block_sum_kernel<<<blocks, 256, 0, st>>>(N, r0, r1, c0, c1, alpha, Kinv, ldi, d);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    final_sum_kernel<<<1, 32, 0, st>>>(d, blocks, d + 2 * blocks);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    double tmp[2];
    int rc = copy_scalars(tmp, d + 2 * blocks, 2, st);
    out_host[0] = tmp[0];
    cudaFreeAsync(d, st);
    return rc;
}

extern "C" int gprb_predict(int m, int N, const double *Ks, long long ldks, const double *alpha,
                            const double *Kinv, long long ldi, const double *diag,
                            double *mean, double *var, double *work, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GPRB_REQUIRE(Ks && alpha && mean && m >= 0 && N > 0, "gprb_predict: bad argument");
    if (m == 0) return GPRB_OK;
    if (var) {
        GPRB_REQUIRE(Kinv && diag && work, "gprb_predict: variance needs Kinv, diag and work");
        int rc = handles(st);
        if (rc) return rc;
        // row-major work[m,N] = Ks[m,N] . Kinv[N,N]  ==  column-major work^T = Kinv^T . Ks^T
        const double one = 1.0, zero = 0.0;
        cublasStatus_t bs = cublasDgemm(g_blas, CUBLAS_OP_N, CUBLAS_OP_N, N, m, N, &one, Kinv, (int)ldi, Ks, (int)ldks,
                                        &zero, work, N);
        if (bs != CUBLAS_STATUS_SUCCESS) { gprb_set_error("cublasDgemm status %d", (int)bs); return GPRB_ERR_CUDA; }
    }
    predict_rows_kernel<<<m, 256, 0, st>>>(N, Ks, ldks, alpha, var ? work : nullptr, diag, mean, var);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    return GPRB_OK;
}

// one CTA per test row: mean = Ks[i,:].alpha ; var = max(diag - |Y[i,:]|^2, 0) with Y = (L^-1 Ks^T)^T
__global__ void __launch_bounds__(256) predict_rows_chol_kernel(int N, const double *__restrict__ Ks, long long ldks,
                                                                const double *__restrict__ alpha, const double *__restrict__ Y,
                                                                const double *__restrict__ diag, double *mean, double *var) {
    __shared__ double sh[32];
    const int i = blockIdx.x;
    const double *k = Ks + (long long)i * ldks;
    const double *y = Y + (long long)i * N;
    double m = 0.0, v = 0.0;
    for (int j = threadIdx.x; j < N; j += blockDim.x) { m = fma(k[j], alpha[j], m); const double t = y[j]; v = fma(t, t, v); }
    m = block_sum(m, sh);
    v = block_sum(v, sh);
    if (threadIdx.x == 0) {
        mean[i] = m;
        const double r = diag[i] - v;
        var[i] = r < 0.0 ? 0.0 : r;
    }
}

extern "C" int gprb_predict_chol(int m, int N, const double *Ks, long long ldks, const double *alpha,
                                 const double *L, long long ldl, const double *diag,
                                 double *mean, double *var, double *work, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GPRB_REQUIRE(Ks && alpha && mean && var && L && diag && work && m >= 0 && N > 0, "gprb_predict_chol: bad argument");
    if (m == 0) return GPRB_OK;
    int rc = handles(st);
    if (rc) return rc;
    // k*^T K^-1 k* = |L^-1 k*|^2: one triangular solve with m right-hand sides (m N^2 flops) instead of the
    // product with the explicit inverse (2 m N^2).  Column-major view: work (N x m) = Ks^T, U^T Y = Ks^T.
    GPRB_CUDA(cudaMemcpy2DAsync(work, (size_t)N * sizeof(double), Ks, (size_t)ldks * sizeof(double), (size_t)N * sizeof(double), m,
                                cudaMemcpyDeviceToDevice, st));
    const double one = 1.0;
    cublasStatus_t bs = cublasDtrsm_64(g_blas, CUBLAS_SIDE_LEFT, factor_uplo(), solve_op(0), CUBLAS_DIAG_NON_UNIT,
                                       (int64_t)N, (int64_t)m, &one, L, (int64_t)ldl, work, (int64_t)N);
    if (bs != CUBLAS_STATUS_SUCCESS) { gprb_set_error("cublasDtrsm_64 (predict) status %d", (int)bs); return GPRB_ERR_CUDA; }
    predict_rows_chol_kernel<<<m, 256, 0, st>>>(N, Ks, ldks, alpha, work, diag, mean, var);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    return GPRB_OK;
}

// dst[j][i] = src[i][j] for a rows x cols block (32 x 32 tiles through shared memory)
__global__ void transpose_copy_kernel(double *dst, long long ldd, const double *src, long long lds, int rows, int cols) {
    __shared__ double t[32][33];
    const int i = blockIdx.y * 32 + threadIdx.y, j = blockIdx.x * 32 + threadIdx.x;
    if (i < rows && j < cols) t[threadIdx.y][threadIdx.x] = src[(long long)i * lds + j];
    __syncthreads();
    const int tj = blockIdx.x * 32 + threadIdx.y, ti = blockIdx.y * 32 + threadIdx.x;   // dst row = src col
    if (tj < cols && ti < rows) dst[(long long)tj * ldd + ti] = t[threadIdx.x][threadIdx.y];
}

extern "C" int gprb_transpose_copy(double *dst, long long ldd, const double *src, long long lds, int rows, int cols, void *stream) {
    GPRB_REQUIRE(dst && src && rows >= 0 && cols >= 0, "gprb_transpose_copy: bad argument");
    if (rows == 0 || cols == 0) return GPRB_OK;
    dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 32);
    transpose_copy_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(dst, ldd, src, lds, rows, cols);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    return GPRB_OK;
}

extern "C" int gprb_symmetrize(double *A, long long ld, int n, void *stream) {
    GPRB_REQUIRE(A && n >= 0, "gprb_symmetrize: bad argument");
    if (n == 0) return GPRB_OK;
    dim3 grid((n + 31) / 32, (n + 31) / 32), block(32, 32);
    mirror_upper_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(A, ld, n);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    return GPRB_OK;
}

namespace {
__global__ void dmma_peak_kernel(double *out, int iters) {
    double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    if (s == 123.456) out[threadIdx.x] = s;
}
}  // namespace

extern "C" int gprb_fp64_dmma_peak(double *tflops_host, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GPRB_REQUIRE(tflops_host, "gprb_fp64_dmma_peak: NULL output");
    int dev = 0, sms = 0;
    GPRB_CUDA(cudaGetDevice(&dev));
    GPRB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double *d = nullptr;
    GPRB_CUDA(cudaMalloc((void **)&d, 1024 * sizeof(double)));
    cudaEvent_t e0, e1;
    GPRB_CUDA(cudaEventCreate(&e0));
    GPRB_CUDA(cudaEventCreate(&e1));
    const int iters = 20000, warps = 8;
    dmma_peak_kernel<<<sms, warps * 32, 0, st>>>(d, iters / 10);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        GPRB_CUDA(cudaEventRecord(e0, st));
        dmma_peak_kernel<<<sms, warps * 32, 0, st>>>(d, iters);
        GPRB_CUDA(cudaEventRecord(e1, st));
        GPRB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        GPRB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    *tflops_host = (double)sms * warps * iters * 16.0 * 512.0 / (best * 1e-3) * 1e-12;
    return GPRB_OK;
}
