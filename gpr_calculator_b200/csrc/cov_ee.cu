// cov_ee.cu — energy-energy covariance K_ee and the eps-regularised energy diagonal.
//
// Replaces rbf_kee_many / rbf_kee_many_with_grad (rbf_kernel.cpp:5-98), dot_kee_many
// (dot_kernel.cpp:5-56) with their wrapper normalisation 1/(n_I n_J) (rbf_kernel.py:56-70,
// dot_kernel.py:46), and the numpy K_ee_RBF / K_ee used only by diag()
// (kernels/base.py:107-130, Dot_mb.py:177-202).
//
// Scalar kernels: the eps-regularised diagonal, and K_ee for descriptors longer than 32 (the DMMA route of
// gprb_kee lives in cov_mma.cu).  2*d flops per pair, a plain
// DFMA kernel: one CTA per (row group I, column-group slice), the rows of I staged in shared
// memory, one warp per column group J, lanes over the pairs; fixed-order reductions.
#include "common.cuh"

namespace {

struct EEParams {
    const double *PA; const int *eleA; const int *row_ptrA; const int *rowsA;
    const double *PB; const int *eleB; const int *row_ptrB; const int *rowsB;
    int ks, n_groupsB, grp_begin;
    double c_sigma2, c_i2l2, c_il3, c_sigma02, zeta;
    int zi, kernel;
    double *K; long long ldk; double *dK; long long lddk;
};

__device__ __forceinline__ double pow_z(double s, double zeta, int zi) {
    if (zi == 2) return s * s;
    if (zi == 3) return s * s * s;
    if (zi == 1) return s;
    if (zi == 4) { double t = s * s; return t * t; }
    return pow(s, zeta);
}

// element k of flat row `prow` of an energy pack (ncomp = 1)
__device__ __forceinline__ const double *row_base(const double *P, int ks, int prow) {
    return P + (size_t)(prow >> 3) * ks * 32 + (prow & 7) * 4;
}

constexpr int EE_MAX_SROWS = 96;   // rows of group I staged in smem per pass (96 * 32 * 8 B = 24 KB at ks = 8)

__global__ void __launch_bounds__(256) kee_kernel(const EEParams P) {
    extern __shared__ double sA[];   // [EE_MAX_SROWS][4*ks]
    __shared__ int sEleA[EE_MAX_SROWS];
    const int I = P.grp_begin + blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int kp = 4 * P.ks;
    const int rowA0 = P.row_ptrA[I];
    const int nA = P.rowsA[I];
    for (int a0 = 0; a0 < max(nA, 1); a0 += EE_MAX_SROWS) {
        const int na = max(0, min(EE_MAX_SROWS, nA - a0));
        __syncthreads();
        for (int i = threadIdx.x; i < na * kp; i += blockDim.x) {
            const int r = i / kp, k = i - r * kp;
            sA[i] = row_base(P.PA, P.ks, rowA0 + a0 + r)[(k >> 2) * 32 + (k & 3)];
        }
        for (int i = threadIdx.x; i < na; i += blockDim.x) sEleA[i] = P.eleA[rowA0 + a0 + i];
        __syncthreads();
        // column groups handled by this CTA: J = blockIdx.y*nwarps + warp, stride gridDim.y*nwarps
        for (int J = blockIdx.y * nwarps + warp; J < P.n_groupsB; J += gridDim.y * nwarps) {
            const int rowB0 = P.row_ptrB[J];
            const int nB = P.rowsB[J];
            double accK = 0.0, accD = 0.0;
            const int npairs = na * nB;
            for (int pidx = lane; pidx < npairs; pidx += 32) {
                const int b = pidx / na, a = pidx - b * na;
                const int eb = P.eleB[rowB0 + b];
                if (eb < 0 || eb != sEleA[a]) continue;
                const double *xb = row_base(P.PB, P.ks, rowB0 + b);
                const double *xa = sA + a * kp;
                double s = 0.0;
                for (int kk = 0; kk < P.ks; kk++) {
                    const double2 b01 = *reinterpret_cast<const double2 *>(xb + kk * 32);
                    const double2 b23 = *reinterpret_cast<const double2 *>(xb + kk * 32 + 2);
                    s = fma(xa[4 * kk + 0], b01.x, s); s = fma(xa[4 * kk + 1], b01.y, s);
                    s = fma(xa[4 * kk + 2], b23.x, s); s = fma(xa[4 * kk + 3], b23.y, s);
                }
                const double D = pow_z(s, P.zeta, P.zi);
                if (P.kernel == GPRB_KERNEL_RBF) {
                    const double Kv = P.c_sigma2 * exp((D - 1.0) * P.c_i2l2);
                    accK += Kv;
                    accD += Kv * (1.0 - D);
                } else {
                    accK += P.c_sigma2 * (D + P.c_sigma02);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                accK += __shfl_xor_sync(0xffffffffu, accK, o);
                accD += __shfl_xor_sync(0xffffffffu, accD, o);
            }
            if (lane == 0) {   // the same lane owns (I, J) in every pass: plain read-modify-write is race free
                const double nn = (double)nA * (double)nB;
                const long long r = blockIdx.x;
                const double vK = nn > 0 ? accK / nn : 0.0, vD = nn > 0 ? accD / nn * P.c_il3 : 0.0;
                if (a0 == 0) { P.K[r * P.ldk + J] = vK; if (P.dK) P.dK[r * P.lddk + J] = vD; }
                else { P.K[r * P.ldk + J] += vK; if (P.dK) P.dK[r * P.lddk + J] += vD; }
            }
        }
    }
}

// diag(): energy rows, k(I,I) with norms (|x|+eps) and d = x1.x2 / (eps + n1 n2); no zero-norm drop.
__global__ void __launch_bounds__(128) kee_diag_kernel(const EEParams P, const double *normA, double eps) {
    const int I = blockIdx.x;
    const int row0 = P.row_ptrA[I], n = P.rowsA[I];
    double acc = 0.0;
    for (int pidx = threadIdx.x; pidx < n * n; pidx += blockDim.x) {
        const int a = pidx / n, b = pidx - a * n;
        int ea = P.eleA[row0 + a], eb = P.eleA[row0 + b];
        ea = ea >= 0 ? ea : -(ea + 2);
        eb = eb >= 0 ? eb : -(eb + 2);
        if (ea != eb) continue;
        const double *xa = row_base(P.PA, P.ks, row0 + a), *xb = row_base(P.PA, P.ks, row0 + b);
        double s = 0.0;
        for (int kk = 0; kk < P.ks; kk++)
            for (int q = 0; q < 4; q++) s = fma(xa[kk * 32 + q], xb[kk * 32 + q], s);
        const double na = normA[row0 + a], nb = normA[row0 + b];
        const double dd = s * na * nb / (eps + (na + eps) * (nb + eps));
        const double D = pow_z(dd, P.zeta, P.zi);
        if (P.kernel == GPRB_KERNEL_RBF) acc += P.c_sigma2 * exp((D - 1.0) * P.c_i2l2);
        else acc += P.c_sigma2 * (D + P.c_sigma02);
    }
    __shared__ double red[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        const double nn = (double)n * (double)n;
        P.K[I] = nn > 0 ? (red[0] + red[1] + red[2] + red[3]) / nn : 0.0;
    }
}

int fill(EEParams &P, int kernel, double p0, double p1, double zeta) {
    P.kernel = kernel;
    P.zeta = zeta;
    int zi = (int)zeta;
    P.zi = ((double)zi == zeta && zi >= 1 && zi <= 4) ? zi : -1;
    P.c_sigma2 = p0 * p0;
    if (kernel == GPRB_KERNEL_RBF) {
        GPRB_REQUIRE(p1 > 0.0, "length scale l must be positive, got %g", p1);
        P.c_i2l2 = 1.0 / (2.0 * p1 * p1);
        P.c_il3 = 1.0 / (p1 * p1 * p1);
        P.c_sigma02 = 0.0;
    } else {
        P.c_i2l2 = P.c_il3 = 0.0;
        P.c_sigma02 = p1 * p1;
    }
    return GPRB_OK;
}

}  // namespace

// scalar K_ee (any descriptor length that fits the staging buffer): the route of gprb_kee (cov_mma.cu) for d > 32
int gprb_kee_scalar(int kernel, const gprb_pack *e1, const gprb_pack *e2, double p0, double p1, double zeta,
                    int grp_begin, int grp_end, double *K, long long ldk, double *dK, long long lddk, cudaStream_t st) {
    GPRB_REQUIRE(e1 && e2 && K, "gprb_kee: NULL argument");
    { int rcd = gprb_check_device(e1, "gprb_kee"); if (rcd || (rcd = gprb_check_device(e2, "gprb_kee"))) return rcd; }
    GPRB_REQUIRE(e1->ncols == 0 && e2->ncols == 0, "gprb_kee: both sides must be energy packs");
    GPRB_REQUIRE(e1->d == e2->d, "gprb_kee: descriptor length mismatch %d vs %d", e1->d, e2->d);
    GPRB_REQUIRE(kernel == GPRB_KERNEL_RBF || kernel == GPRB_KERNEL_DOT, "gprb_kee: unknown kernel %d", kernel);
    GPRB_REQUIRE(0 <= grp_begin && grp_begin <= grp_end && grp_end <= e1->n_groups, "gprb_kee: bad window [%d,%d)", grp_begin, grp_end);
    GPRB_REQUIRE(!(dK && kernel == GPRB_KERNEL_DOT), "gprb_kee: Dot has no dK output (closed form, see header)");
    if (grp_begin == grp_end || e2->n_groups == 0) return GPRB_OK;
    EEParams P = {};
    int rc = fill(P, kernel, p0, p1, zeta);
    if (rc) return rc;
    P.PA = e1->P; P.eleA = e1->elep; P.row_ptrA = e1->d_row_ptr; P.rowsA = e1->d_group_rows;
    P.PB = e2->P; P.eleB = e2->elep; P.row_ptrB = e2->d_row_ptr; P.rowsB = e2->d_group_rows;
    P.ks = e1->ks; P.n_groupsB = e2->n_groups; P.grp_begin = grp_begin;
    P.K = K; P.ldk = ldk; P.dK = dK; P.lddk = lddk;
    const int nI = grp_end - grp_begin;
    int gy = (e2->n_groups + 7) / 8;
    int want = (8 * 148 + nI - 1) / nI;
    if (gy > want) gy = want;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    const size_t smem = (size_t)EE_MAX_SROWS * 4 * P.ks * sizeof(double);
    GPRB_REQUIRE(smem <= 48 * 1024, "gprb_kee: descriptor length %d too large for the staging buffer", e1->d);
    kee_kernel<<<dim3(nI, gy), 256, smem, st>>>(P);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    return GPRB_OK;
}

extern "C" int gprb_kee_diag(int kernel, const gprb_pack *e, double p0, double p1, double zeta, double *out, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    GPRB_REQUIRE(e && out, "gprb_kee_diag: NULL argument");
    { int rcd = gprb_check_device(e, "gprb_kee_diag"); if (rcd) return rcd; }
    GPRB_REQUIRE(e->ncols == 0, "gprb_kee_diag: need an energy pack");
    if (e->n_groups == 0) return GPRB_OK;
    EEParams P = {};
    int rc = fill(P, kernel, p0, p1, zeta);
    if (rc) return rc;
    P.PA = e->P; P.eleA = e->elep; P.row_ptrA = e->d_row_ptr; P.rowsA = e->d_group_rows; P.ks = e->ks;
    P.K = out;
    kee_diag_kernel<<<e->n_groups, 128, 0, st>>>(P, e->norm, GPRB_EPS_NORM);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    return GPRB_OK;
}
