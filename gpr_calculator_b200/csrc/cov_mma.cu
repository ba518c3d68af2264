// cov_mma.cu — covariance blocks K_ff, K_ef/K_fe on the FP64 tensor pipe (DMMA.8x8x4) of sm_100a.
//
// Replaces rbf_kff_many / rbf_kff_many_with_grad / dot_kff_many and rbf_kef_many(_with_grad) /
// dot_kef_many of the reference (rbf_kernel.cpp:101-253, 341-640; dot_kernel.cpp:58-130, 217-336).
//
// Algebra.  With x^ = x/|x|, A~ = (I - x^x^T) dx/dr / |x| precomputed per row
// (pack.cu), every quantity the reference derives per pair from a d x d matrix is one entry of
//        [x^_a ; A~_a^T] (4 x d)  .  [x^_b ; B~_b^T]^T (d x 4)
//   s = x^_a.x^_b     p_c = A~_a[:,c].x^_b     q_e = x^_a.B~_b[:,e]     G_ce = A~_a[:,c].B~_b[:,e]
//   K_ff[3I+c,3J+e] = sum_{a in I, b in J}  g (beta G_ce + gamma p_c q_e)
// so a block of 8 x 8 atom pairs is 16 m8n8k4 accumulator tiles.  Rows are laid out
// component-major (tile = 8 atoms of one component), which puts the complete 4x4 result of the
// pairs (a = lane/4, b = 2*(lane%4)+{0,1}) into the registers of ONE thread: the scalar epilogue
// (exp, powers, rank-1 correction) needs no shuffles.
//
// Execution: one CTA = up to 8 warps = up to 8 row tiles covering whole groups (force centres);
// each warp keeps its A fragments in registers for the whole kernel.  Column tiles are streamed
// through a 2-stage shared-memory ring by 1-D TMA bulk copies (cp.async.bulk + mbarrier), one
// chunk = the tiles of one column group (<= 8 tiles).  Per (I, J) the per-thread partial 3x3
// sums are reduced by warp shuffles, then across the warps of the group through shared memory in
// a fixed order (deterministic), and written once.  Groups larger than 64 rows are split over
// CTAs and combined with fp64 atomics.
#include "common.cuh"
#include <cstdint>

namespace {

constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;

struct CovParams {
    const double *PA; const int *eleA; const int *tile_groupA;
    const int4 *sched;                 // row blocks {tile0, ntiles, group0, flags(bit0 = split group -> atomics)}
    const double *PB; const int *eleB; const int4 *chunks; const int *gcpB; const int *group_rowsB;
    int n_groupsB;
    int n_splits;
    double c_sigma2, c_i2l2, c_il, c_il3, zeta, tol, c_dot;   // c_dot = sigma^2 * zeta
    int zi, use_tol, mode, grp_begin;
    double *K; long long ldk; double *dK; long long lddk;     // kff: K / dK/dl ; kfe: Kfe / dKfe
    double *K2; long long ldk2; double *dK2; long long lddk2; // kfe only: Kef / dKef (transposed copies)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// s^(zeta-2) with the integer cases the reference uses in practice (zeta = 2 default,
// gaussianprocess.py:1027) kept off the pow() path
__device__ __forceinline__ double pow_zm2(double s, double zeta, int zi) {
    if (zi == 2) return 1.0;
    if (zi == 3) return s;
    if (zi == 4) return s * s;
    if (zi == 1) return 1.0 / s;
    return pow(s, zeta - 2.0);
}

// Per-pair scalar weights.  out = w1*G + w2*p q^T ; grad: dout = u1*G + u2*p q^T
// (for NB == 1: out_c = w1 * p_c, dout_c = u1 * p_c)
template <int KERNEL, bool GRAD, bool FF>
__device__ __forceinline__ void pair_weights(const CovParams &P, double s, bool valid,
                                             double &w1, double &w2, double &u1, double &u2) {
    const double sm2 = pow_zm2(s, P.zeta, P.zi);
    const double sm1 = s * sm2;
    if (KERNEL == GPRB_KERNEL_RBF) {
        const double D = s * sm1;
        const double Kv = P.c_sigma2 * exp((D - 1.0) * P.c_i2l2);     // rbf_kernel.cpp:393
        const double g = Kv * P.c_i2l2;                                // dK_dD (:394)
        if (FF && !GRAD && P.use_tol) valid = valid && (g > P.tol);    // (:395) pair cut, non-grad only
        const double gz = valid ? g * P.zeta : 0.0;
        w1 = gz * sm1;                                                 // g * beta
        const double z2 = gz * P.zeta * sm1 * sm1;                     // g zeta^2 s^(2zeta-2)
        if (FF) w2 = gz * (P.zeta - 1.0) * sm2 + z2 * P.c_i2l2;        // g * gamma
        if (GRAD) {
            const double h = (1.0 - D) * P.c_il3 - 2.0 * P.c_il;       // (:622-630), (:245-247)
            u1 = w1 * h;
            if (FF) u2 = w2 * h - z2 * P.c_il3;
        }
    } else {   // Dot: sigma^2 zeta (s^(z-1) G + (z-1) s^(z-2) p q^T)   (dot_kernel.cpp:285-289, dot_kernel.py:256)
        const double cz = valid ? P.c_dot : 0.0;
        w1 = cz * sm1;
        if (FF) w2 = cz * (P.zeta - 1.0) * sm2;
    }
}

template <int NB, int KS, int KERNEL, bool GRAD>
__global__ void __launch_bounds__(THREADS, 1) cov_mma_kernel(const CovParams P) {
    constexpr bool FF = (NB == 4);
    constexpr int NOUT = FF ? 9 : 3;
    constexpr int NTOT = GRAD ? 2 * NOUT : NOUT;
    constexpr int TILE_DOUBLES = NB * KS * 32;
    constexpr uint32_t TILE_BYTES = TILE_DOUBLES * 8;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sB = reinterpret_cast<double *>(smem_raw);                         // [2][CHUNK][TILE_DOUBLES]
    int *sEle = reinterpret_cast<int *>(sB + 2 * GPRB_CHUNK_TILES * TILE_DOUBLES);   // [2][CHUNK*8]
    double *sRed = reinterpret_cast<double *>(sEle + 2 * GPRB_CHUNK_TILES * 8);      // [2][WARPS][NTOT]
    uint64_t *sBar = reinterpret_cast<uint64_t *>(sRed + 2 * WARPS * NTOT);          // [2]
    int *sWg = reinterpret_cast<int *>(sBar + 2);                                     // [WARPS] local group of warp

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int4 blk = P.sched[blockIdx.x];
    const int tile0 = blk.x, ntiles = blk.y, g0 = blk.z;
    const bool split = blk.w & 1;
    const bool active = warp < ntiles;
    const int atile = tile0 + (active ? warp : 0);
    const int g1 = P.tile_groupA[tile0 + ntiles - 1] + 1;      // one past the last group of this block

    if (tid < WARPS) sWg[tid] = tid < ntiles ? P.tile_groupA[tile0 + tid] - g0 : -1;
    if (tid == 0) {
        mbar_init(&sBar[0], 1);
        mbar_init(&sBar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // column chunk range of this CTA
    int c_begin, c_end;
    {
        const int G = P.n_groupsB;
        int ga = (int)((long long)G * blockIdx.y / P.n_splits);
        int gb = (int)((long long)G * (blockIdx.y + 1) / P.n_splits);
        if (P.mode == GPRB_FF_SYMMETRIC || P.mode == GPRB_FF_UPPER) ga = max(ga, g0);
        if (P.mode == GPRB_FF_DIAG) { ga = max(ga, g0); gb = min(gb, g1); }
        if (gb < ga) gb = ga;
        c_begin = P.gcpB[ga];
        c_end = P.gcpB[gb];
    }
    if (c_begin >= c_end) return;

    // A fragments: registers for the whole kernel
    double af[4][KS];
    {
        const double *pa = P.PA + (size_t)atile * 4 * KS * 32 + lane;
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
            for (int k = 0; k < KS; k++) af[c][k] = active ? pa[(c * KS + k) * 32] : 0.0;
    }
    const int ele_a = active ? P.eleA[atile * 8 + (lane >> 2)] : -1;

    auto issue = [&](int ci, int buf) {
        const int4 ch = P.chunks[ci];
        const uint32_t bytes = (uint32_t)ch.y * TILE_BYTES, ebytes = (uint32_t)ch.y * 32u;
        mbar_expect_tx(&sBar[buf], bytes + ebytes);
        bulk_g2s(sB + (size_t)buf * GPRB_CHUNK_TILES * TILE_DOUBLES, P.PB + (size_t)ch.x * TILE_DOUBLES, bytes, &sBar[buf]);
        bulk_g2s(sEle + buf * GPRB_CHUNK_TILES * 8, P.eleB + (size_t)ch.x * 8, ebytes, &sBar[buf]);
    };
    if (tid == 0) {
        issue(c_begin, 0);
        if (c_begin + 1 < c_end) issue(c_begin + 1, 1);
    }

    double out[NTOT];
#pragma unroll
    for (int i = 0; i < NTOT; i++) out[i] = 0.0;
    int par = 0;

    for (int ci = c_begin; ci < c_end; ci++) {
        const int it = ci - c_begin, buf = it & 1;
        const int4 ch = P.chunks[ci];
        mbar_wait(&sBar[buf], (it >> 1) & 1);
        if (active) {
            const double *tb = sB + (size_t)buf * GPRB_CHUNK_TILES * TILE_DOUBLES + lane;
            const int *te = sEle + buf * GPRB_CHUNK_TILES * 8 + 2 * (lane & 3);
            for (int t = 0; t < ch.y; t++, tb += TILE_DOUBLES, te += 8) {
                double acc[4][NB][2];
#pragma unroll
                for (int c = 0; c < 4; c++)
#pragma unroll
                    for (int e = 0; e < NB; e++) { acc[c][e][0] = 0.0; acc[c][e][1] = 0.0; }
#pragma unroll
                for (int k = 0; k < KS; k++) {
                    double bf[NB];
#pragma unroll
                    for (int e = 0; e < NB; e++) bf[e] = tb[(e * KS + k) * 32];
#pragma unroll
                    for (int c = 0; c < 4; c++)
#pragma unroll
                        for (int e = 0; e < NB; e++) dmma(acc[c][e][0], acc[c][e][1], af[c][k], bf[e]);
                }
                const int2 eb = *reinterpret_cast<const int2 *>(te);
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int ele_b = j ? eb.y : eb.x;
                    const bool valid = (ele_a == ele_b) && (ele_a >= 0);
                    double w1, w2 = 0.0, u1 = 0.0, u2 = 0.0;
                    pair_weights<KERNEL, GRAD, FF>(P, acc[0][0][j], valid, w1, w2, u1, u2);
                    if (FF) {
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const double pc = acc[c + 1][0][j];
                            const double t2 = w2 * pc;
                            const double t3 = GRAD ? u2 * pc : 0.0;
#pragma unroll
                            for (int e = 0; e < 3; e++) {
                                const double G = acc[c + 1][(NB == 4) ? e + 1 : 0][j];
                                const double q = acc[0][(NB == 4) ? e + 1 : 0][j];
                                out[c * 3 + e] = fma(w1, G, fma(t2, q, out[c * 3 + e]));
                                if (GRAD) out[NOUT + c * 3 + e] = fma(u1, G, fma(t3, q, out[NOUT + c * 3 + e]));
                            }
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const double pc = acc[c + 1][0][j];
                            out[c] = fma(w1, pc, out[c]);
                            if (GRAD) out[NOUT + c] = fma(u1, pc, out[NOUT + c]);
                        }
                    }
                }
            }
        }
        if (ch.w) {   // last chunk of column group J: reduce over the warp
#pragma unroll
            for (int i = 0; i < NTOT; i++) {
                double v = out[i];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) sRed[(par * WARPS + warp) * NTOT + i] = v;
                out[i] = 0.0;
            }
        }
        __syncthreads();   // everyone is done with stage `buf`; sRed[par] is visible
        if (tid == 0 && ci + 2 < c_end) issue(ci + 2, buf);
        if (ch.w) {
            const int J = ch.z;
            const int ngl = g1 - g0;
            if (tid < ngl * NTOT) {
                const int lg = tid / NTOT, o = tid - lg * NTOT;
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < WARPS; w++)
                    if (sWg[w] == lg) v += sRed[(par * WARPS + w) * NTOT + o];
                const int I = g0 + lg;
                const bool isgrad = GRAD && o >= NOUT;
                const int oo = isgrad ? o - NOUT : o;
                if (FF) {
                    const int c = oo / 3, e = oo - 3 * c;
                    double *dst = isgrad ? P.dK : P.K;
                    const long long ld = isgrad ? P.lddk : P.ldk;
                    if (P.mode == GPRB_FF_DIAG) {
                        if (I == J && c == e) {
                            double *q = dst + 3 * (I - P.grp_begin) + c;
                            if (split) atomicAdd(q, v); else *q = v;
                        }
                    } else {
                        double *q = dst + (long long)(3 * (I - P.grp_begin) + c) * ld + 3 * J + e;
                        if (split) atomicAdd(q, v); else *q = v;
                        if (P.mode == GPRB_FF_SYMMETRIC && J >= g1) {
                            double *qt = dst + (long long)(3 * J + e) * ld + 3 * I + c;
                            if (split) atomicAdd(qt, v); else *qt = v;
                        }
                    }
                } else {
                    // a side = force group I (window), b side = energy group J; K_ef = -(1/n_J) sum
                    const int nJ = P.group_rowsB[J];
                    const double val = nJ > 0 ? -v / (double)nJ : 0.0;
                    const long long row = 3 * (I - P.grp_begin) + oo;
                    double *fe = isgrad ? P.dK : P.K;
                    const long long ldfe = isgrad ? P.lddk : P.ldk;
                    double *ef = isgrad ? P.dK2 : P.K2;
                    const long long ldef = isgrad ? P.lddk2 : P.ldk2;
                    if (fe) { double *q = fe + row * ldfe + J; if (split) atomicAdd(q, val); else *q = val; }
                    if (ef) { double *q = ef + (long long)J * ldef + row; if (split) atomicAdd(q, val); else *q = val; }
                }
            }
            par ^= 1;
        }
    }
}

template <int NB, int KS>
constexpr size_t cov_smem_bytes(bool grad) {
    return (size_t)2 * GPRB_CHUNK_TILES * NB * KS * 32 * 8 + 2 * GPRB_CHUNK_TILES * 8 * 4 +
           (size_t)2 * WARPS * ((NB == 4 ? 9 : 3) * (grad ? 2 : 1)) * 8 + 16 + WARPS * 4 + 128;
}

// Row-side schedule: blocks of <= WARPS tiles made of whole groups; larger groups are split and flagged.
int build_sched(gprb_pack *a, int g0, int g1, cudaStream_t st) {
    if (a->sched_g0 == g0 && a->sched_g1 == g1 && a->sched) return GPRB_OK;
    std::vector<int4> s;
    int g = g0;
    while (g < g1) {
        const int t0 = a->tile_ptr[g];
        const int nt = a->tile_ptr[g + 1] - t0;
        if (nt > WARPS) {
            for (int o = 0; o < nt; o += WARPS) s.push_back(make_int4(t0 + o, nt - o < WARPS ? nt - o : WARPS, g, 1));
            g++;
            continue;
        }
        int ge = g + 1;
        while (ge < g1 && a->tile_ptr[ge + 1] - t0 <= WARPS) ge++;
        s.push_back(make_int4(t0, a->tile_ptr[ge] - t0, g, 0));
        g = ge;
    }
    if (a->sched) { GPRB_CUDA(cudaFree(a->sched)); a->sched = nullptr; }
    a->sched_host = s;
    a->sched_n = (int)s.size();
    a->sched_g0 = g0; a->sched_g1 = g1;
    if (!s.empty()) {
        GPRB_CUDA(cudaMalloc((void **)&a->sched, s.size() * sizeof(int4)));
        GPRB_CUDA(cudaMemcpyAsync(a->sched, a->sched_host.data(), s.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
    }
    return GPRB_OK;
}

bool sched_has_split(const gprb_pack *a) {
    for (auto &b : a->sched_host) if (b.w & 1) return true;
    return false;
}

int integer_zeta(double zeta) {
    int zi = (int)zeta;
    return ((double)zi == zeta && zi >= 1 && zi <= 4) ? zi : -1;
}

template <int NB, int KS, int KERNEL, bool GRAD>
int launch_cov(const CovParams &P, int n_blocks, cudaStream_t st) {
    auto kern = cov_mma_kernel<NB, KS, KERNEL, GRAD>;
    const size_t smem = cov_smem_bytes<NB, KS>(GRAD);
    static bool configured = false;
    if (!configured) {
        GPRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid(n_blocks, P.n_splits);
    kern<<<grid, THREADS, smem, st>>>(P);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    return GPRB_OK;
}

template <int NB, int KS>
int dispatch_cov(int kernel, bool grad, const CovParams &P, int n_blocks, cudaStream_t st) {
    if (kernel == GPRB_KERNEL_RBF) return grad ? launch_cov<NB, KS, GPRB_KERNEL_RBF, true>(P, n_blocks, st)
                                               : launch_cov<NB, KS, GPRB_KERNEL_RBF, false>(P, n_blocks, st);
    return launch_cov<NB, KS, GPRB_KERNEL_DOT, false>(P, n_blocks, st);
}

int fill_kernel_params(CovParams &P, int kernel, double p0, double p1, double zeta) {
    P.zeta = zeta;
    P.zi = integer_zeta(zeta);
    P.c_sigma2 = p0 * p0;
    if (kernel == GPRB_KERNEL_RBF) {
        GPRB_REQUIRE(p1 > 0.0, "length scale l must be positive, got %g", p1);
        P.c_i2l2 = 1.0 / (2.0 * p1 * p1);
        P.c_il = 1.0 / p1;
        P.c_il3 = 1.0 / (p1 * p1 * p1);
    } else {
        P.c_i2l2 = P.c_il = P.c_il3 = 0.0;
    }
    P.c_dot = p0 * p0 * zeta;
    return GPRB_OK;
}

int choose_splits(int n_blocks, int n_groupsB) {
    int sms = 148;
    int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int want = (4 * sms + n_blocks - 1) / (n_blocks > 0 ? n_blocks : 1);
    if (want < 1) want = 1;
    if (want > n_groupsB) want = n_groupsB;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    return want;
}

}  // namespace

extern "C" int gprb_kff(int kernel, const gprb_pack *f1_, const gprb_pack *f2, double p0, double p1, double zeta,
                        int use_tol, double tol, int mode, int grp_begin, int grp_end,
                        double *K, long long ldk, double *dK, long long lddk, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    gprb_pack *f1 = const_cast<gprb_pack *>(f1_);
    GPRB_REQUIRE(f1 && f2 && K, "gprb_kff: NULL argument");
    GPRB_REQUIRE(f1->ncols == 3 && f2->ncols == 3, "gprb_kff: both sides must be force packs");
    GPRB_REQUIRE(f1->d == f2->d, "gprb_kff: descriptor length mismatch %d vs %d", f1->d, f2->d);
    GPRB_REQUIRE(kernel == GPRB_KERNEL_RBF || kernel == GPRB_KERNEL_DOT, "gprb_kff: unknown kernel %d", kernel);
    GPRB_REQUIRE(mode >= 0 && mode <= 3, "gprb_kff: unknown mode %d", mode);
    GPRB_REQUIRE(0 <= grp_begin && grp_begin <= grp_end && grp_end <= f1->n_groups, "gprb_kff: bad window [%d,%d)", grp_begin, grp_end);
    GPRB_REQUIRE(!(dK && kernel == GPRB_KERNEL_DOT), "gprb_kff: Dot has no dK output (closed form, see header)");
    if (mode == GPRB_FF_SYMMETRIC)
        GPRB_REQUIRE(f1 == f2 && grp_begin == 0 && grp_end == f1->n_groups, "gprb_kff: symmetric mode needs f1 == f2 and the full window");
    if (mode == GPRB_FF_DIAG || mode == GPRB_FF_UPPER) GPRB_REQUIRE(f1 == f2, "gprb_kff: diag / upper mode needs f1 == f2");
    if (f1->ks > GPRB_MAX_KS) {
        gprb_set_error("gprb_kff: descriptor length %d > 32 is not supported by the DMMA kernels yet", f1->d);
        return GPRB_ERR_UNSUPPORTED;
    }
    if (grp_begin == grp_end || f2->n_groups == 0) return GPRB_OK;
    int rc = build_sched(f1, grp_begin, grp_end, st);
    if (rc) return rc;
    const int rows = 3 * (grp_end - grp_begin);
    if (sched_has_split(f1)) {
        if (mode == GPRB_FF_DIAG) {
            GPRB_CUDA(cudaMemsetAsync(K, 0, (size_t)rows * sizeof(double), st));
        } else {
            GPRB_CUDA(cudaMemset2DAsync(K, ldk * sizeof(double), 0, (size_t)3 * f2->n_groups * sizeof(double), rows, st));
            if (dK) GPRB_CUDA(cudaMemset2DAsync(dK, lddk * sizeof(double), 0, (size_t)3 * f2->n_groups * sizeof(double), rows, st));
        }
    }
    CovParams P = {};
    rc = fill_kernel_params(P, kernel, p0, p1, zeta);
    if (rc) return rc;
    P.PA = f1->P; P.eleA = f1->elep; P.tile_groupA = f1->tile_group; P.sched = f1->sched;
    P.PB = f2->P; P.eleB = f2->elep; P.chunks = f2->chunks; P.gcpB = f2->d_group_chunk_ptr;
    P.group_rowsB = f2->d_group_rows; P.n_groupsB = f2->n_groups;
    P.tol = tol; P.use_tol = use_tol; P.mode = mode; P.grp_begin = grp_begin;
    P.K = K; P.ldk = ldk; P.dK = dK; P.lddk = lddk;
    P.n_splits = mode == GPRB_FF_DIAG ? 1 : choose_splits(f1->sched_n, f2->n_groups);
    switch (f1->ks) {
        case 8: return dispatch_cov<4, 8>(kernel, dK != nullptr, P, f1->sched_n, st);
        case 7: return dispatch_cov<4, 7>(kernel, dK != nullptr, P, f1->sched_n, st);
        case 6: return dispatch_cov<4, 6>(kernel, dK != nullptr, P, f1->sched_n, st);
        case 5: return dispatch_cov<4, 5>(kernel, dK != nullptr, P, f1->sched_n, st);
        case 4: return dispatch_cov<4, 4>(kernel, dK != nullptr, P, f1->sched_n, st);
        case 3: return dispatch_cov<4, 3>(kernel, dK != nullptr, P, f1->sched_n, st);
        case 2: return dispatch_cov<4, 2>(kernel, dK != nullptr, P, f1->sched_n, st);
        default: return dispatch_cov<4, 1>(kernel, dK != nullptr, P, f1->sched_n, st);
    }
}

extern "C" int gprb_kef(int kernel, const gprb_pack *e, const gprb_pack *f_, double p0, double p1, double zeta,
                        int grp_begin, int grp_end,
                        double *Kef, long long ld_ef, double *Kfe, long long ld_fe,
                        double *dKef, long long ld_def, double *dKfe, long long ld_dfe, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    gprb_pack *f = const_cast<gprb_pack *>(f_);
    GPRB_REQUIRE(e && f && (Kef || Kfe), "gprb_kef: NULL argument");
    GPRB_REQUIRE(e->ncols == 0 && f->ncols == 3, "gprb_kef: need (energy pack, force pack)");
    GPRB_REQUIRE(e->d == f->d, "gprb_kef: descriptor length mismatch %d vs %d", e->d, f->d);
    GPRB_REQUIRE(kernel == GPRB_KERNEL_RBF || kernel == GPRB_KERNEL_DOT, "gprb_kef: unknown kernel %d", kernel);
    GPRB_REQUIRE(0 <= grp_begin && grp_begin <= grp_end && grp_end <= f->n_groups, "gprb_kef: bad window [%d,%d)", grp_begin, grp_end);
    const bool grad = dKef || dKfe;
    GPRB_REQUIRE(!(grad && kernel == GPRB_KERNEL_DOT), "gprb_kef: Dot has no dK output (closed form, see header)");
    if (f->ks > GPRB_MAX_KS) {
        gprb_set_error("gprb_kef: descriptor length %d > 32 is not supported by the DMMA kernels yet", f->d);
        return GPRB_ERR_UNSUPPORTED;
    }
    if (grp_begin == grp_end || e->n_groups == 0) return GPRB_OK;
    int rc = build_sched(f, grp_begin, grp_end, st);
    if (rc) return rc;
    const int rows = 3 * (grp_end - grp_begin);
    if (sched_has_split(f)) {
        const size_t w = (size_t)e->n_groups * sizeof(double);
        if (Kfe) GPRB_CUDA(cudaMemset2DAsync(Kfe, ld_fe * sizeof(double), 0, w, rows, st));
        if (dKfe) GPRB_CUDA(cudaMemset2DAsync(dKfe, ld_dfe * sizeof(double), 0, w, rows, st));
        if (Kef) GPRB_CUDA(cudaMemset2DAsync(Kef, ld_ef * sizeof(double), 0, (size_t)rows * sizeof(double), e->n_groups, st));
        if (dKef) GPRB_CUDA(cudaMemset2DAsync(dKef, ld_def * sizeof(double), 0, (size_t)rows * sizeof(double), e->n_groups, st));
    }
    CovParams P = {};
    rc = fill_kernel_params(P, kernel, p0, p1, zeta);
    if (rc) return rc;
    P.PA = f->P; P.eleA = f->elep; P.tile_groupA = f->tile_group; P.sched = f->sched;
    P.PB = e->P; P.eleB = e->elep; P.chunks = e->chunks; P.gcpB = e->d_group_chunk_ptr;
    P.group_rowsB = e->d_group_rows; P.n_groupsB = e->n_groups;
    P.mode = GPRB_FF_FULL; P.grp_begin = grp_begin;
    P.K = Kfe; P.ldk = ld_fe; P.dK = dKfe; P.lddk = ld_dfe;
    P.K2 = Kef; P.ldk2 = ld_ef; P.dK2 = dKef; P.lddk2 = ld_def;
    P.n_splits = choose_splits(f->sched_n, e->n_groups);
    switch (f->ks) {
        case 8: return dispatch_cov<1, 8>(kernel, grad, P, f->sched_n, st);
        case 7: return dispatch_cov<1, 7>(kernel, grad, P, f->sched_n, st);
        case 6: return dispatch_cov<1, 6>(kernel, grad, P, f->sched_n, st);
        case 5: return dispatch_cov<1, 5>(kernel, grad, P, f->sched_n, st);
        case 4: return dispatch_cov<1, 4>(kernel, grad, P, f->sched_n, st);
        case 3: return dispatch_cov<1, 3>(kernel, grad, P, f->sched_n, st);
        case 2: return dispatch_cov<1, 2>(kernel, grad, P, f->sched_n, st);
        default: return dispatch_cov<1, 1>(kernel, grad, P, f->sched_n, st);
    }
}
