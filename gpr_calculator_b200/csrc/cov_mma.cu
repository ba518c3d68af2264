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
// component-major (tile = 8 rows of one component), which puts the complete 4x4 result of the
// pairs (a = lane/4, b = 2*(lane%4)+{0,1}) into the registers of ONE thread: the scalar epilogue
// (exp, powers, rank-1 correction) needs no shuffles.
//
// Generation 2 execution model:
//  * FLAT tiles: rows of neighbouring groups share 8-row tiles on both sides, so no DMMA work is
//    spent on per-group padding.  Column tiles carry a record of their group segments; a tile that
//    straddles a group boundary is multiplied once and accumulated once per segment with masked
//    weights.  Row tiles are handled by a segmented reduction (static per-thread predicates).
//  * one CTA = 12 warps = 96 consecutive rows; the row tiles live in shared memory (TMA bulk copy
//    at start) so that the kernel fits 168 registers and every scheduler has 3 warps to overlap
//    one warp's scalar epilogue with the other warps' DMMAs (DMMA and DFMA share the FP64 pipe).
//  * column tiles stream through a ring of TMA stages (cp.async.bulk + mbarrier); the warp that
//    releases a stage last refills it — no producer warp and no CTA-wide barrier in the main loop.
//  * flush at the end of a column group: transposing butterfly over the 4 lanes of a row, segmented
//    suffix sum over the 8 rows of the warp, head rows to one of four shared-memory buffers, one
//    mbarrier arrival per warp.  Two flushes later every warp adds the per-warp partials of its share
//    of the outputs in a fixed order and writes them (deterministic; no atomics on the critical path,
//    no designated finisher).
//    Row groups that continue in a neighbouring CTA are combined with fp64 atomics into the
//    pre-zeroed output (two addends: still order independent).
#include "common.cuh"
#pragma nv_diag_suppress 128   // the 4x4-block path below the two-stage branch is unreachable in the two-stage instantiations
#include <cstdint>
#include <cmath>
#include <cstdlib>

namespace {

constexpr int WARPS = 12;
constexpr int THREADS = WARPS * 32;
constexpr int STAGES = 2;
constexpr int CH = GPRB_CHUNK_TILES;
constexpr int REC = GPRB_REC_INTS;
constexpr int MAXROWS = WARPS * 8;
constexpr int BIG_WARPS = 6, BIG_CH = 2;   // CTA shape of the 9..16 k-step kernels (d = 33..64)
constexpr int NFB = 4;                 // flush buffers: partials of flush k are summed two flushes later

struct CovParams {
    const double *PA; const int *eleA; const int *row_groupA;
    const int4 *sched; const int4 *sched_ent;
    const double *PB; const int *recB; const int *row_ptrB; const int *group_rowsB;
    int n_groupsB, n_splits, ks;
    int win_r0, win_r1;                                   // row window on side A (flat rows)
    // k1 = sigma^2/(2 l^2), kz = k1*zeta, c = 1/(2 l^2), h0 = 1/l^3 - 2/l ; Dot: c_dot = sigma^2 zeta
    double k1, kz, c, c_il3, h0, zeta, tol, c_dot;
    double s2, s02;                                          // EE blocks: sigma^2 and (Dot) sigma0^2
    const int *group_rowsA;                                  // EE blocks: rows per group of side A (1 / (n_I n_J))
    int zi, use_tol, mode, grp_begin;
    double *K; long long ldk; double *dK; long long lddk;     // kff: K / dK/dl ; kfe: Kfe / dKfe
    double *K2; long long ldk2; double *dK2; long long lddk2; // kfe only: Kef / dKef (transposed copies)
    // fused all-gather (gprb_kff_multi / gprb_kfe_multi): every finished value of K is also stored into the
    // same slab of n_extra peer matrices (NVLink peer stores, pointers from gprb_peer_open); dK stays local
    double *Kx[GPRB_MAX_DST - 1]; int n_extra;
    int two_stage;                                            // two-stage contraction (no-gradient K_ff, 8 k-steps)
};


__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded waits: a protocol error must surface as a launch failure (trap), never as a hung GPU.
constexpr long long SPIN_LIMIT_CYCLES = 8000000000LL;     // ~4 s at 1.9 GHz
__device__ __noinline__ void spin_timeout(int what) {
    printf("libgpr_b200: cov_mma_kernel wait %d timed out (block %d,%d thread %d)\n", what, blockIdx.x, blockIdx.y, threadIdx.x);
    __trap();
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try(bar, parity))
        if (clock64() - t0 > SPIN_LIMIT_CYCLES) spin_timeout(1);
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// exp(x) for x <= 0 (the RBF exponent -(1-D)/(2 l^2)): 2^(k/32) table + degree-5 polynomial on
// |r| <= ln2/64 (truncation 2e-15 relative), evaluated in Estrin form.  (Round 2 measured a variant with a 1024-entry
// table, degree 3, one-step reduction and an integer clamp -- 5 instead of 8 dependent FP64 steps: same kernel time,
// 1 812 vs 1 811 ms at S5, so the dependent chain is not what bounds the epilogue; profiles/experiments/README.md.)
constexpr int EXP_TAB = 32;
__device__ double d_exp2_tab[EXP_TAB];    // 2^(j / 32)

__device__ __forceinline__ double exp_neg(double x, const double *tab) {
    x = fmax(x, -700.0);                                     // exp(-700) ~ 1e-304: below anything that matters, no branch
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: low word = round-to-nearest integer
    const double t = fma(x, 46.16624130844682903, MAGIC);    // 32 / ln 2
    const int ki = __double2loint(t);
    const double kf = t - MAGIC;
    double r = fma(kf, -0x1.62e42fe000000p-6, x);            // ln2/32, high 29 bits (kf * hi is exact)
    r = fma(kf, -0x1.f473de6af278fp-35, r);                   // ln2/32 - high part
    const double r2 = r * r;
    const double a = fma(r, 1.66666666666666657e-01, 0.5);
    const double b = fma(r, 8.33333333333333322e-03, 4.16666666666666644e-02);
    const double c1 = 1.0 + r;
    const double d = fma(r2, a, c1);
    const double r4 = r2 * r2;
    const double p = fma(r4, b, d);                           // 1 + r + r^2/2 + r^3/6 + r^4/24 + r^5/120
    const double v = tab[ki & 31] * p;
    return __hiloint2double(__double2hiint(v) + ((ki >> 5) << 20), __double2loint(v));
}

// s^(zeta-2): ZI = 2 is compile-time (the reference's default zeta, gaussianprocess.py:1027);
// ZI = 0 picks the small integer cases at run time and falls back to pow()
template <int ZI>
__device__ __forceinline__ double pow_zm2(double s, double zeta, int zi) {
    if (ZI == 2) return 1.0;
    if (zi == 2) return 1.0;
    if (zi == 3) return s;
    if (zi == 4) return s * s;
    if (zi == 1) return 1.0 / s;
    return pow(s, zeta - 2.0);
}

// Per-pair scalar weights.  out = w1*G + w2*p q^T ; grad: dout = u1*G + u2*p q^T
// (for NB == 1: out_c = w1 * p_c, dout_c = u1 * p_c).  `valid` may be cleared by the pair cut.
template <int KERNEL, bool GRAD, bool FF, int ZI, bool EE = false>
__device__ __forceinline__ void pair_weights(const CovParams &P, const double *tab, double s, bool &valid,
                                             double &w1, double &w2, double &u1, double &u2) {
    const double sm2 = pow_zm2<ZI>(s, P.zeta, P.zi);
    const double sm1 = s * sm2;
    if (EE) {
        // energy-energy pair: w1 = k(a, b), u1 = dk/dl = k (1 - D) / l^3   (rbf_kernel.cpp:40-47, 86-94; dot_kernel.cpp:38-44)
        const double D = s * sm1;
        if (KERNEL == GPRB_KERNEL_RBF) {
            w1 = P.s2 * exp_neg(fma(D, P.c, -P.c), tab);
            if (GRAD) u1 = w1 * fma(-P.c_il3, D, P.c_il3);
        } else {
            w1 = P.s2 * (D + P.s02);
        }
        return;
    }
    if (KERNEL == GPRB_KERNEL_RBF) {
        // every weight is E times a factor that does not depend on E: the factors are computed next to the exp
        // chain (instruction-level parallelism), one multiply each once E is known
        const double D = s * sm1;
        const double f1 = P.kz * sm1;                                  // w1 = g beta          = E f1
        const double q2 = f1 * (P.zeta * sm1);                         // g zeta^2 s^(2zeta-2) = E q2
        const double f2 = (ZI == 2) ? fma(q2, P.c, P.kz) : fma(q2, P.c, P.kz * (P.zeta - 1.0) * sm2);   // w2 = g gamma = E f2
        const double h = fma(-P.c_il3, D, P.h0);                       // (1-D)/l^3 - 2/l   (:622-630, :245-247)
        const double g1 = f1 * h, g2 = fma(f2, h, -(q2 * P.c_il3));
        const double E = exp_neg(fma(D, P.c, -P.c), tab);             // exp(-(1-D)/(2l^2))  rbf_kernel.cpp:393
        if (FF && !GRAD && P.use_tol) valid = valid && (E * P.k1 > P.tol);   // dK_dD > tol (:394-395), non-grad only
        w1 = E * f1;
        if (FF) w2 = E * f2;
        if (GRAD) {
            u1 = E * g1;
            if (FF) u2 = E * g2;
        }
    } else {   // Dot: sigma^2 zeta (s^(z-1) G + (z-1) s^(z-2) p q^T)   (dot_kernel.cpp:285-289, dot_kernel.py:256)
        w1 = P.c_dot * sm1;
        if (FF) w2 = P.c_dot * (P.zeta - 1.0) * sm2;
    }
}

// Flush of the per-thread sums `out[NT]` when column group J is complete: transposing butterfly over the 4 lanes of a
// row, segmented suffix sum over the 8 rows of the warp, head rows to flush buffer flush_idx & 3, one mbarrier arrival per
// warp; the final sums of flush k are taken two flushes later (finish).  A macro so that the 4x4-block path and the
// two-stage path expand the very same statements.
#define COV_FLUSH_GROUP(J) \
    do { \
                if (flush_idx >= 2) finish(flush_idx - 2, J2); \
                const int fb = flush_idx & (NFB - 1); \
                const bool b0 = lane & 1, b1 = lane & 2; \
                double wv[N1], zv[N2]; \
_Pragma("unroll") \
                for (int i = 0; i < N1; i++) { \
                    const double lo = out[2 * i], hi = (2 * i + 1 < NT) ? out[(2 * i + 1 < NT) ? 2 * i + 1 : 0] : 0.0; \
                    const double keep = b0 ? hi : lo, send = b0 ? lo : hi; \
                    wv[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1); \
                } \
_Pragma("unroll") \
                for (int i = 0; i < N2; i++) { \
                    const double lo = wv[2 * i], hi = (2 * i + 1 < N1) ? wv[(2 * i + 1 < N1) ? 2 * i + 1 : 0] : 0.0; \
                    const double keep = b1 ? hi : lo, send = b1 ? lo : hi; \
                    zv[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2); \
                } \
_Pragma("unroll") \
                for (int i = 0; i < N2; i++) { \
                    double tv = __shfl_down_sync(0xffffffffu, zv[i], 4); \
                    zv[i] += p1 ? tv : 0.0; \
                    tv = __shfl_down_sync(0xffffffffu, zv[i], 8); \
                    zv[i] += p2 ? tv : 0.0; \
                    tv = __shfl_down_sync(0xffffffffu, zv[i], 16); \
                    zv[i] += p4 ? tv : 0.0; \
                } \
                if (is_head) { \
                    double *dst = sRow + (size_t)(fb * MAXROWS_K + arow_local) * NT + q4; \
_Pragma("unroll") \
                    for (int i = 0; i < N2; i++) \
                        if (4 * i + q4 < NT) dst[4 * i] = zv[i]; \
                } \
_Pragma("unroll") \
                for (int i = 0; i < NT; i++) out[i] = 0.0; \
                __syncwarp(); \
                if (lane == 0) mbar_arrive(&barFlush[fb]); \
                J2 = J1; J1 = J; \
                flush_idx++; \
    } while (0)

// NB = component tiles per column tile: 4 = force columns against force rows (K_ff), 1 = energy columns against force
// rows (K_fe / K_ef), 0 = energy columns against ENERGY rows (K_ee: one component on both sides, two CTAs per SM)
template <int NB, int KS_T, int KERNEL, bool GRAD, int ZI, bool MULTI, bool TWO, bool BIG>
__global__ void __launch_bounds__(BIG ? BIG_WARPS * 32 : THREADS, NB == 0 ? 2 : 1) cov_mma_kernel(const CovParams P) {
    // BIG: descriptors of 33..64 entries (9..16 k-steps): the row tiles of 12 warps no longer fit the shared memory, so 6 warps per CTA and
    // 2-tile column chunks (functional path for e.g. SO3(nmax=4, lmax=4), d = 50; the benchmarked descriptor has d = 30)
    constexpr int WARPS_K = BIG ? BIG_WARPS : WARPS, CH_K = BIG ? BIG_CH : CH, MAXROWS_K = WARPS_K * 8;
    constexpr bool FF = (NB == 4), EE = (NB == 0);
    constexpr int NA = EE ? 1 : 4, NBC = EE ? 1 : NB;        // component tiles of a row tile / of a column tile
    constexpr int NOUT = FF ? 9 : (EE ? 1 : 3);
    constexpr int NT = GRAD ? 2 * NOUT : NOUT;
    constexpr int N1 = (NT + 1) / 2, N2 = (N1 + 1) / 2;      // values left after each transposing butterfly step
    const int KS = KS_T ? KS_T : P.ks;
    const int a_tile_d = NA * KS * 32, b_tile_d = NBC * KS * 32;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sA = reinterpret_cast<double *>(smem_raw);                  // [WARPS_K][a_tile_d]
    double *sB = sA + WARPS_K * a_tile_d;                                  // [STAGES][CH_K][b_tile_d]
    double *sRow = sB + STAGES * CH_K * b_tile_d;                          // [NFB][MAXROWS_K][NT] flush partials
    double *sTab = sRow + NFB * MAXROWS_K * NT;                            // [EXP_TAB]
    int4 *sEnt = reinterpret_cast<int4 *>(sTab + EXP_TAB);               // [MAXROWS_K]
    int *sRec = reinterpret_cast<int *>(sEnt + MAXROWS_K);                 // [STAGES][CH_K][REC]
    uint64_t *sBar = reinterpret_cast<uint64_t *>(sRec + STAGES * CH_K * REC);   // full[STAGES], A, flush[NFB]
    uint64_t *barFlush = sBar + STAGES + 1;
    int *sCnt = reinterpret_cast<int *>(barFlush + NFB);                 // consumed[STAGES]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int4 blk = P.sched[blockIdx.x];
    const int tile0 = blk.x, ntiles = blk.y, ent0 = blk.z, nent = blk.w;

    // column group range of this CTA -> column tiles / chunks
    int ga, gb;
    {
        const int G = P.n_groupsB;
        const int first_group = P.sched_ent[ent0].x, last_group = P.sched_ent[ent0 + nent - 1].x;
        // the column groups this row block sweeps, cut evenly over the n_splits CTAs of the block
        int lo = 0, hi = G;
        if (P.mode == GPRB_FF_SYMMETRIC || P.mode == GPRB_FF_UPPER) lo = first_group;
        if (P.mode == GPRB_FF_DIAG) { lo = first_group; hi = last_group + 1; }
        ga = lo + (int)((long long)(hi - lo) * blockIdx.y / P.n_splits);
        gb = lo + (int)((long long)(hi - lo) * (blockIdx.y + 1) / P.n_splits);
    }
    if (gb <= ga) return;
    const int cr0 = P.row_ptrB[ga], cr1 = P.row_ptrB[gb];
    if (cr1 <= cr0) return;
    const int tb0 = cr0 >> 3, tb1 = (cr1 + 7) >> 3;
    const int c_begin = tb0 / CH_K, c_end = (tb1 + CH_K - 1) / CH_K;

    if (tid < nent) sEnt[tid] = P.sched_ent[ent0 + tid];
    for (int i = tid; i < EXP_TAB; i += blockDim.x) sTab[i] = d_exp2_tab[i];
    if (tid == 0) {
        for (int s = 0; s <= STAGES; s++) mbar_init(&sBar[s], 1);
        for (int s = 0; s < NFB; s++) mbar_init(&barFlush[s], ntiles);     // one arrival per active warp and flush
        for (int s = 0; s < STAGES; s++) sCnt[s] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const uint32_t b_chunk_bytes = (uint32_t)(CH_K * b_tile_d * 8), rec_chunk_bytes = (uint32_t)(CH_K * REC * 4);
    auto issue = [&](int ci, int buf) {
        mbar_expect_tx(&sBar[buf], b_chunk_bytes + rec_chunk_bytes);
        bulk_g2s(sB + (size_t)buf * CH_K * b_tile_d, P.PB + (size_t)ci * CH_K * b_tile_d, b_chunk_bytes, &sBar[buf]);
        bulk_g2s(sRec + buf * CH_K * REC, P.recB + (size_t)ci * CH_K * REC, rec_chunk_bytes, &sBar[buf]);
    };
    if (tid == 0) {
        const uint32_t a_bytes = (uint32_t)(ntiles * a_tile_d * 8);
        mbar_expect_tx(&sBar[STAGES], a_bytes);
        bulk_g2s(sA, P.PA + (size_t)tile0 * a_tile_d, a_bytes, &sBar[STAGES]);
        for (int s = 0; s < STAGES; s++)
            if (c_begin + s < c_end) issue(c_begin + s, s);
    }
    if (warp >= ntiles) return;          // no CTA-wide barrier below this line
    const int n_active = ntiles;

    // static per-thread row data: species, window, segment structure of the warp's 8 rows
    const int ra = lane >> 2, q4 = lane & 3;
    const int arow_local = warp * 8 + ra;
    const int arow = (tile0 + warp) * 8 + ra;
    int ele_a = P.eleA[arow];
    if (arow < P.win_r0 || arow >= P.win_r1) ele_a = -1;
    const int g_a = P.row_groupA[arow];
    // (shuffles first, unconditionally: every lane must take part in a full-mask shuffle)
    const int g_d1 = __shfl_down_sync(0xffffffffu, g_a, 4), g_d2 = __shfl_down_sync(0xffffffffu, g_a, 8);
    const int g_d4 = __shfl_down_sync(0xffffffffu, g_a, 16), g_u1 = __shfl_up_sync(0xffffffffu, g_a, 4);
    const bool p1 = (ra + 1 < 8) && (g_d1 == g_a);
    const bool p2 = (ra + 2 < 8) && (g_d2 == g_a);
    const bool p4 = (ra + 4 < 8) && (g_d4 == g_a);
    const bool is_head = (ra == 0) || (g_u1 != g_a);

    double out[NT];
#pragma unroll
    for (int i = 0; i < NT; i++) out[i] = 0.0;
    double Z[TWO ? 3 : 1][TWO ? 4 : 1][2];      // two-stage path: Z_e[ra][8 t + 2 q4 + {0, 1}] of the current column group
#pragma unroll
    for (int e = 0; e < (TWO ? 3 : 1); e++)
#pragma unroll
        for (int t = 0; t < (TWO ? 4 : 1); t++) { Z[e][t][0] = 0.0; Z[e][t][1] = 0.0; }
    int flush_idx = 0, J1 = 0, J2 = 0;      // J1 / J2: column groups of the previous two flushes

    // Final sums of flush kk (column group J): every warp adds the per-warp partials of ITS share of the
    // (row group, component) outputs in a fixed order and writes them.  Called two flushes later, when the
    // partials of all warps have long been delivered (mbarrier per flush buffer; no atomics, no designated
    // finisher warp, deterministic).  Buffer reuse is safe with NFB = 4: a warp overwrites buffer k & 3 at
    // flush k only after it has seen flush k-2 complete, i.e. after every warp has summed its share of k-4.
    auto finish = [&](int kk, int J) {
        const int fb = kk & (NFB - 1);
        mbar_wait(&barFlush[fb], (kk / NFB) & 1);
        const int total = nent * NT;
        for (int idx = lane * n_active + warp; idx < total; idx += 32 * n_active) {
            const int en = idx / NT, o = idx - en * NT;
            const int4 E4 = sEnt[en];
            const int I = E4.x;
            const double *src = sRow + (size_t)fb * MAXROWS_K * NT + o;
            double v = src[E4.y * NT];
            for (int r = (E4.y & ~7) + 8; r < E4.z; r += 8) v += src[r * NT];
            const bool shared = E4.w != 0;
            const bool isgrad = GRAD && o >= NOUT;
            const int oo = isgrad ? o - NOUT : o;
            if (FF) {
                const int c = oo / 3, e = oo - 3 * c;
                double *dst = isgrad ? P.dK : P.K;
                const long long ld = isgrad ? P.lddk : P.ldk;
                if (P.mode == GPRB_FF_DIAG) {
                    if (I == J && c == e) {
                        double *qd = dst + 3 * (I - P.grp_begin) + c;
                        if (shared) atomicAdd(qd, v); else *qd = v;
                    }
                } else if (!((P.mode == GPRB_FF_SYMMETRIC || P.mode == GPRB_FF_UPPER) && J < I)) {
                    const long long off = (long long)(3 * (I - P.grp_begin) + c) * ld + 3 * J + e;
                    double *qd = dst + off;
                    if (shared) atomicAdd(qd, v); else *qd = v;
                    if (MULTI && !isgrad) {
#pragma unroll
                        for (int p = 0; p < GPRB_MAX_DST - 1; p++)
                            if (p < P.n_extra) { double *qp = P.Kx[p] + off; if (shared) atomicAdd(qp, v); else *qp = v; }
                    }
                    if (P.mode == GPRB_FF_SYMMETRIC && J > I) {
                        double *qt = dst + (long long)(3 * J + e) * ld + 3 * I + c;
                        if (shared) atomicAdd(qt, v); else *qt = v;
                    }
                }
            } else if (EE) {
                // K_ee[I, J] = 1 / (n_I n_J) sum k   (rbf_kernel.py:56-70, dot_kernel.py:46); symmetric mode mirrors
                const double nn = (double)P.group_rowsA[I] * (double)P.group_rowsB[J];
                const double val = nn > 0 ? v / nn : 0.0;
                double *dst = isgrad ? P.dK : P.K;
                const long long ld = isgrad ? P.lddk : P.ldk;
                if (!(P.mode == GPRB_FF_SYMMETRIC && J < I)) {
                    double *qd = dst + (long long)(I - P.grp_begin) * ld + J;
                    if (shared) atomicAdd(qd, val); else *qd = val;
                    if (P.mode == GPRB_FF_SYMMETRIC && J > I) {
                        double *qt = dst + (long long)J * ld + I;
                        if (shared) atomicAdd(qt, val); else *qt = val;
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
                if (fe) {
                    const long long off = row * ldfe + J;
                    double *qd = fe + off;
                    if (shared) atomicAdd(qd, val); else *qd = val;
                    if (MULTI && !isgrad) {
#pragma unroll
                        for (int p = 0; p < GPRB_MAX_DST - 1; p++)
                            if (p < P.n_extra) { double *qp = P.Kx[p] + off; if (shared) atomicAdd(qp, val); else *qp = val; }
                    }
                }
                if (ef) { double *qd = ef + (long long)J * ldef + row; if (shared) atomicAdd(qd, val); else *qd = val; }
            }
        }
    };

    mbar_wait(&sBar[STAGES], 0);         // row tiles have landed
    const double *pa = sA + (size_t)warp * a_tile_d + lane;

    for (int ci = c_begin; ci < c_end; ci++) {
        const int it = ci - c_begin, buf = it % STAGES;
        mbar_wait(&sBar[buf], (it / STAGES) & 1);
        for (int tt = 0; tt < CH_K; tt++) {
            const int t = ci * CH_K + tt;
            if (t < tb0 || t >= tb1) continue;
            const int *rec = sRec + (buf * CH_K + tt) * REC;
            const double *pb = sB + (size_t)(buf * CH_K + tt) * b_tile_d + lane;

            if constexpr (TWO) {
                // Two-stage contraction (no gradient, d in 29..32; design: profiles/experiments/README.md and
                // two_stage_emulation.py; parity: test_two_stage_path_matches_block_path).  Stage 1: x^(a) . [x^; B~_e](b) -> s, q_e
                // (32 DMMAs).  Stage 2, per column segment: Z_e[a, :] += W1 B~_e + (W2 q_e) x^ with the stage-1 accumulator
                // registers as A fragments (k order b = 2 q4 + j) and the B fragments read from the same slabs (48 DMMAs).
                // Per column group: out_ce = A~_c(a) . Z_e(a), then the common flush.
                static_assert(!TWO || (NB == 4 && KS_T == 8 && !GRAD), "two-stage path: K_ff without gradient, 8 k-steps");
                double a1[4][2];
#pragma unroll
                for (int e = 0; e < 4; e++) { a1[e][0] = 0.0; a1[e][1] = 0.0; }
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const double af = pa[k * 32];
#pragma unroll
                    for (int e = 0; e < 4; e++) dmma(a1[e][0], a1[e][1], af, pb[(e * 8 + k) * 32]);
                }
                const int2 eb2 = *reinterpret_cast<const int2 *>(rec + 2 * q4);
                double tw1[2], tw2[2];
                bool tvalid[2];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int ele_b = j ? eb2.y : eb2.x;
                    tvalid[j] = (ele_a == ele_b) && (ele_a >= 0);
                    double du1 = 0.0, du2 = 0.0;
                    tw2[j] = 0.0;
                    pair_weights<KERNEL, false, true, ZI>(P, sTab, a1[0][j], tvalid[j], tw1[j], tw2[j], du1, du2);
                }
                const double *pbt = pb - lane;                      // tile base: stage 2 addresses the slabs by (row b, column)
                const int nseg2 = rec[8];
                for (int sg = 0; sg < nseg2; sg++) {
                    const int J = rec[10 + 2 * sg], mf = rec[11 + 2 * sg];
                    if (J < ga || J >= gb) continue;
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const bool on = tvalid[j] && ((mf >> (2 * q4 + j)) & 1);
                        const double W1 = on ? tw1[j] : 0.0;
                        double V[3];
#pragma unroll
                        for (int e = 0; e < 3; e++) V[e] = on ? tw2[j] * a1[1 + e][j] : 0.0;
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            // B fragment of n-tile t: element (k = q4 -> row b = 2 q4 + j, n = ra -> column 8 t + ra)
                            const int off = (2 * t + (ra >> 2)) * 32 + (2 * q4 + j) * 4 + (ra & 3);
                            const double bx = pbt[off];
#pragma unroll
                            for (int e = 0; e < 3; e++) {
                                dmma(Z[e][t][0], Z[e][t][1], W1, pbt[(1 + e) * 8 * 32 + off]);
                                dmma(Z[e][t][0], Z[e][t][1], V[e], bx);
                            }
                        }
                    }
                    if (!(mf & 0x100)) continue;
                    // column group J is complete: out_ce = sum over this thread's 8 columns of A~_c[ra, col] Z_e[ra, col]
                    const double *par = pa - lane;                  // row tile base (component c + 1 holds A~_c)
#pragma unroll
                    for (int c = 0; c < 3; c++) {
#pragma unroll
                        for (int t = 0; t < 4; t++) {
                            // columns 8 t + 2 q4 + {0, 1}: k-step 2 t + (q4 >> 1), kk = 2 (q4 & 1) + {0, 1} (adjacent: one 16-byte read)
                            const double2 av = *reinterpret_cast<const double2 *>(
                                par + ((c + 1) * 8 + 2 * t + (q4 >> 1)) * 32 + ra * 4 + 2 * (q4 & 1));
#pragma unroll
                            for (int e = 0; e < 3; e++) out[c * 3 + e] = fma(av.x, Z[e][t][0], fma(av.y, Z[e][t][1], out[c * 3 + e]));
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 3; e++)
#pragma unroll
                        for (int t = 0; t < 4; t++) { Z[e][t][0] = 0.0; Z[e][t][1] = 0.0; }
                    COV_FLUSH_GROUP(J);
                }
                continue;
            }

            // species of this thread's two pairs (a = ra, b = 2*q4 + j): read after the DMMAs, except for K_ee
            bool valid[2];
            if (EE) {
                const int2 eb = *reinterpret_cast<const int2 *>(rec + 2 * q4);
                valid[0] = (ele_a == eb.x) && (ele_a >= 0);
                valid[1] = (ele_a == eb.y) && (ele_a >= 0);
                // energy rows of one structure are usually ordered by species (all Mg, all O, ...): a pair of tiles without any
                // same-species pair contributes nothing (rbf_kernel.cpp:26-37) -- skip its DMMAs and exp; a group that ends in
                // the tile still has to be flushed
                if (!__any_sync(0xffffffffu, valid[0] || valid[1])) {
                    const int nseg0 = rec[8];
                    for (int sg = 0; sg < nseg0; sg++) {
                        const int J = rec[10 + 2 * sg], mf = rec[11 + 2 * sg];
                        if (J < ga || J >= gb || !(mf & 0x100)) continue;
                        COV_FLUSH_GROUP(J);
                    }
                    continue;
                }
            }
            double acc[NA][NBC][2];
#pragma unroll
            for (int c = 0; c < NA; c++)
#pragma unroll
                for (int e = 0; e < NBC; e++) { acc[c][e][0] = 0.0; acc[c][e][1] = 0.0; }
            if (KS_T) {
#pragma unroll
                for (int k = 0; k < (KS_T ? KS_T : 1); k++) {
                    double af[NA], bf[NBC];
#pragma unroll
                    for (int c = 0; c < NA; c++) af[c] = pa[(c * KS_T + k) * 32];
#pragma unroll
                    for (int e = 0; e < NBC; e++) bf[e] = pb[(e * KS_T + k) * 32];
#pragma unroll
                    for (int c = 0; c < NA; c++)
#pragma unroll
                        for (int e = 0; e < NBC; e++) dmma(acc[c][e][0], acc[c][e][1], af[c], bf[e]);
                }
            } else {
                for (int k = 0; k < KS; k++) {
                    double af[NA], bf[NBC];
#pragma unroll
                    for (int c = 0; c < NA; c++) af[c] = pa[(c * KS + k) * 32];
#pragma unroll
                    for (int e = 0; e < NBC; e++) bf[e] = pb[(e * KS + k) * 32];
#pragma unroll
                    for (int c = 0; c < NA; c++)
#pragma unroll
                        for (int e = 0; e < NBC; e++) dmma(acc[c][e][0], acc[c][e][1], af[c], bf[e]);
                }
            }

            // weights of this thread's two pairs, once per tile
            if (!EE) {
                const int2 eb = *reinterpret_cast<const int2 *>(rec + 2 * q4);
                valid[0] = (ele_a == eb.x) && (ele_a >= 0);
                valid[1] = (ele_a == eb.y) && (ele_a >= 0);
            }
            double w1[2], w2[2], u1[2], u2[2];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                w2[j] = 0.0; u1[j] = 0.0; u2[j] = 0.0;
                pair_weights<KERNEL, GRAD, FF, ZI, EE>(P, sTab, acc[0][0][j], valid[j], w1[j], w2[j], u1[j], u2[j]);
            }

            const int nseg = rec[8];
            for (int sg = 0; sg < nseg; sg++) {
                const int J = rec[10 + 2 * sg], mf = rec[11 + 2 * sg];
                if (J < ga || J >= gb) continue;
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const bool on = valid[j] && ((mf >> (2 * q4 + j)) & 1);
                    const double W1 = on ? w1[j] : 0.0;
                    if (EE) {
                        out[0] += W1;
                        if (GRAD) out[NOUT] += on ? u1[j] : 0.0;
                    } else if (FF) {
                        const double W2 = on ? w2[j] : 0.0, U1 = on ? u1[j] : 0.0, U2 = on ? u2[j] : 0.0;
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const double pc = acc[(NA == 4) ? c + 1 : 0][0][j];
                            const double t2 = W2 * pc;
                            const double t3 = GRAD ? U2 * pc : 0.0;
#pragma unroll
                            for (int e = 0; e < 3; e++) {
                                const double G = acc[(NA == 4) ? c + 1 : 0][(NB == 4) ? e + 1 : 0][j];
                                const double qe = acc[0][(NB == 4) ? e + 1 : 0][j];
                                out[c * 3 + e] = fma(W1, G, fma(t2, qe, out[c * 3 + e]));
                                if (GRAD) out[NOUT + c * 3 + e] = fma(U1, G, fma(t3, qe, out[NOUT + c * 3 + e]));
                            }
                        }
                    } else {
                        const double U1 = on ? u1[j] : 0.0;
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            const double pc = acc[(NA == 4) ? c + 1 : 0][0][j];
                            out[c] = fma(W1, pc, out[c]);
                            if (GRAD) out[NOUT + c] = fma(U1, pc, out[NOUT + c]);
                        }
                    }
                }
                if (!(mf & 0x100)) continue;

                COV_FLUSH_GROUP(J);      // column group J is complete
            }
        }
        // release the stage; the warp that arrives last refills it
        __syncwarp();
        if (lane == 0) {
            const int old = atomicAdd(&sCnt[buf], 1);
            if ((old + 1) % n_active == 0 && ci + STAGES < c_end) issue(ci + STAGES, buf);
        }
    }
    // the sums of the last two flushes
    if (flush_idx >= 2) finish(flush_idx - 2, J2);
    if (flush_idx >= 1) finish(flush_idx - 1, J1);
}

size_t cov_smem_bytes(int nb, int ks, bool grad, bool big) {
    const int nt = (nb == 4 ? 9 : (nb == 0 ? 1 : 3)) * (grad ? 2 : 1);
    const int na = nb == 0 ? 1 : 4, nbc = nb == 0 ? 1 : nb;
    const int warps = big ? BIG_WARPS : WARPS, ch = big ? BIG_CH : CH, maxrows = warps * 8;
    return (size_t)(warps * na * ks * 32 + STAGES * ch * nbc * ks * 32 + NFB * maxrows * nt + EXP_TAB) * 8 +
           (size_t)maxrows * 16 + (size_t)STAGES * ch * REC * 4 + (STAGES + 1 + NFB) * 8 + STAGES * 4 + 128;
}

// Row-side schedule for the window [g0, g1): blocks of <= WARPS consecutive flat tiles and, per block, the
// groups it holds {group, first row, one past last row (block-local), continues in a neighbouring block}.
int build_sched(gprb_pack *a, int g0, int g1, cudaStream_t st) {
    const int wpc = a->ks > GPRB_MAX_KS ? BIG_WARPS : WARPS;       // row tiles per CTA of the kernels this pack runs (fixed by its ks)
    if (a->sched_g0 == g0 && a->sched_g1 == g1 && a->sched) return GPRB_OK;
    std::vector<int4> blocks, ents;
    const int r0 = a->row_ptr[g0], r1 = a->row_ptr[g1];
    if (r1 > r0) {
        const int t0 = r0 / 8, t1 = (r1 + 7) / 8;
        int g = g0;
        for (int tb = t0; tb < t1; tb += wpc) {
            const int nt = t1 - tb < wpc ? t1 - tb : wpc;
            const int lo = tb * 8 > r0 ? tb * 8 : r0, hi = (tb + nt) * 8 < r1 ? (tb + nt) * 8 : r1;
            const int e0 = (int)ents.size();
            while (g < g1 && a->row_ptr[g + 1] <= lo) g++;      // groups that ended before this block (or empty)
            int gg = g;
            while (gg < g1 && a->row_ptr[gg] < hi) {
                const int s = a->row_ptr[gg] > lo ? a->row_ptr[gg] : lo;
                const int e = a->row_ptr[gg + 1] < hi ? a->row_ptr[gg + 1] : hi;
                if (e > s) {
                    const int shared = (a->row_ptr[gg] < lo || a->row_ptr[gg + 1] > hi) ? 1 : 0;
                    ents.push_back(make_int4(gg, s - tb * 8, e - tb * 8, shared));
                }
                gg++;
            }
            if ((int)ents.size() > e0) blocks.push_back(make_int4(tb, nt, e0, (int)ents.size() - e0));
        }
    }
    if (a->sched) { GPRB_CUDA(cudaFreeAsync(a->sched, st)); a->sched = nullptr; }
    if (a->sched_ent) { GPRB_CUDA(cudaFreeAsync(a->sched_ent, st)); a->sched_ent = nullptr; }
    a->sched_host = blocks; a->sched_ent_host = ents;
    a->sched_n = (int)blocks.size();
    a->sched_g0 = g0; a->sched_g1 = g1;
    if (!blocks.empty()) {
        GPRB_CUDA(cudaMallocAsync((void **)&a->sched, blocks.size() * sizeof(int4), st));
        GPRB_CUDA(cudaMallocAsync((void **)&a->sched_ent, ents.size() * sizeof(int4), st));
        GPRB_CUDA(cudaMemcpyAsync(a->sched, a->sched_host.data(), blocks.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
        GPRB_CUDA(cudaMemcpyAsync(a->sched_ent, a->sched_ent_host.data(), ents.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
    }
    return GPRB_OK;
}

int integer_zeta(double zeta) {
    int zi = (int)zeta;
    return ((double)zi == zeta && zi >= 1 && zi <= 4) ? zi : -1;
}

// per-device one-time state (a process may drive several GPUs): __constant__ tables and function attributes exist per device
constexpr int MAX_DEV = 64;
int current_device(int *dev) {
    GPRB_CUDA(cudaGetDevice(dev));
    GPRB_REQUIRE(*dev >= 0 && *dev < MAX_DEV, "device index %d out of range", *dev);
    return GPRB_OK;
}

int upload_tables() {
    static bool done[MAX_DEV] = {};
    int dev = 0;
    { int rc = current_device(&dev); if (rc) return rc; }
    if (done[dev]) return GPRB_OK;
    static double tab[EXP_TAB];
    for (int j = 0; j < EXP_TAB; j++) tab[j] = std::exp2((double)j / (double)EXP_TAB);
    GPRB_CUDA(cudaMemcpyToSymbol(d_exp2_tab, tab, sizeof tab));
    done[dev] = true;
    return GPRB_OK;
}

template <int NB, int KS_T, int KERNEL, bool GRAD, int ZI, bool MULTI, bool TWO = false, bool BIG = false>
int launch_cov_m(const CovParams &P, int n_blocks, cudaStream_t st) {
    auto kern = cov_mma_kernel<NB, KS_T, KERNEL, GRAD, ZI, MULTI, TWO, BIG>;
    static bool configured[MAX_DEV] = {};
    int dev = 0;
    { int rc = current_device(&dev); if (rc) return rc; }
    if (!configured[dev]) {
        GPRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)cov_smem_bytes(NB, BIG ? 2 * GPRB_MAX_KS : GPRB_MAX_KS, GRAD, BIG)));
        configured[dev] = true;
    }
    dim3 grid(n_blocks, P.n_splits);
    kern<<<grid, BIG ? BIG_WARPS * 32 : THREADS, cov_smem_bytes(NB, P.ks, GRAD, BIG), st>>>(P);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    return GPRB_OK;
}

// the single-destination kernels carry no peer-store code at all (MULTI = false)
template <int NB, int KS_T, int KERNEL, bool GRAD, int ZI, bool BIG = false>
int launch_cov(const CovParams &P, int n_blocks, cudaStream_t st) {
    if constexpr (BIG)
        return P.n_extra > 0 ? launch_cov_m<NB, KS_T, KERNEL, GRAD, ZI, true, false, true>(P, n_blocks, st)
                             : launch_cov_m<NB, KS_T, KERNEL, GRAD, ZI, false, false, true>(P, n_blocks, st);
    if constexpr (NB == 4 && KS_T == 8 && !GRAD) {
        if (P.two_stage)
            return P.n_extra > 0 ? launch_cov_m<NB, KS_T, KERNEL, GRAD, ZI, true, true>(P, n_blocks, st)
                                 : launch_cov_m<NB, KS_T, KERNEL, GRAD, ZI, false, true>(P, n_blocks, st);
    }
    return P.n_extra > 0 ? launch_cov_m<NB, KS_T, KERNEL, GRAD, ZI, true>(P, n_blocks, st)
                         : launch_cov_m<NB, KS_T, KERNEL, GRAD, ZI, false>(P, n_blocks, st);
}

template <int NB, int KS_T, int ZI, bool BIG = false>
int dispatch_kernel(int kernel, bool grad, const CovParams &P, int n_blocks, cudaStream_t st) {
    if (kernel == GPRB_KERNEL_RBF) return grad ? launch_cov<NB, KS_T, GPRB_KERNEL_RBF, true, ZI, BIG>(P, n_blocks, st)
                                               : launch_cov<NB, KS_T, GPRB_KERNEL_RBF, false, ZI, BIG>(P, n_blocks, st);
    return launch_cov<NB, KS_T, GPRB_KERNEL_DOT, false, ZI, BIG>(P, n_blocks, st);
}

template <int NB>
int dispatch_cov(int kernel, bool grad, const CovParams &P, int n_blocks, cudaStream_t st) {
    const bool z2 = P.zi == 2;
    if (P.ks > GPRB_MAX_KS) return dispatch_kernel<NB, 0, 0, true>(kernel, grad, P, n_blocks, st);     // d = 33..64
    if (P.ks == 8) return z2 ? dispatch_kernel<NB, 8, 2>(kernel, grad, P, n_blocks, st) : dispatch_kernel<NB, 8, 0>(kernel, grad, P, n_blocks, st);
    return z2 ? dispatch_kernel<NB, 0, 2>(kernel, grad, P, n_blocks, st) : dispatch_kernel<NB, 0, 0>(kernel, grad, P, n_blocks, st);
}

int fill_kernel_params(CovParams &P, int kernel, double p0, double p1, double zeta) {
    P.zeta = zeta;
    P.zi = integer_zeta(zeta);
    if (kernel == GPRB_KERNEL_RBF) {
        GPRB_REQUIRE(p1 > 0.0, "length scale l must be positive, got %g", p1);
        P.c = 1.0 / (2.0 * p1 * p1);
        P.k1 = p0 * p0 * P.c;
        P.kz = P.k1 * zeta;
        P.c_il3 = 1.0 / (p1 * p1 * p1);
        P.h0 = P.c_il3 - 2.0 / p1;
    } else {
        P.c = P.k1 = P.kz = P.c_il3 = P.h0 = 0.0;
    }
    P.c_dot = p0 * p0 * zeta;
    P.s2 = p0 * p0;
    P.s02 = kernel == GPRB_KERNEL_DOT ? p1 * p1 : 0.0;
    return upload_tables();
}

int choose_splits(int n_blocks, int n_groupsB) {
    int sms = 148;
    int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // enough CTAs (24 per SM) that the longest one is a small fraction of the launch: a row block's sweep
    // can take tens of milliseconds, and whole-sweep CTAs would leave a long partially filled last wave
    int want = (24 * sms + n_blocks - 1) / (n_blocks > 0 ? n_blocks : 1);
    if (want < 1) want = 1;
    if (want > n_groupsB) want = n_groupsB;
    if (want < 1) want = 1;
    if (want > 65535) want = 65535;
    return want;
}

}  // namespace

// n_dst > 1: K_dst[1..] are the same row slab in peer matrices (fused all-gather); the outputs are then NOT
// zeroed here (every GPU zeroes its own matrix before the ranks synchronise, see dist.PeerMatrix)
static int kff_impl(int kernel, const gprb_pack *f1_, const gprb_pack *f2, double p0, double p1, double zeta,
                    int use_tol, double tol, int mode, int grp_begin, int grp_end,
                    int n_dst, double *const *K_dst, long long ldk, double *dK, long long lddk, cudaStream_t st) {
    gprb_pack *f1 = const_cast<gprb_pack *>(f1_);
    GPRB_REQUIRE(n_dst >= 1 && n_dst <= GPRB_MAX_DST && K_dst, "gprb_kff: need 1..%d destination matrices", GPRB_MAX_DST);
    for (int p = 0; p < n_dst; p++) GPRB_REQUIRE(K_dst[p], "gprb_kff: NULL destination %d", p);
    double *K = K_dst[0];
    GPRB_REQUIRE(f1 && f2 && K, "gprb_kff: NULL argument");
    { int rcd = gprb_check_device(f1, "gprb_kff"); if (rcd || (rcd = gprb_check_device(f2, "gprb_kff"))) return rcd; }
    GPRB_REQUIRE(f1->ncols == 3 && f2->ncols == 3, "gprb_kff: both sides must be force packs");
    GPRB_REQUIRE(f1->d == f2->d, "gprb_kff: descriptor length mismatch %d vs %d", f1->d, f2->d);
    GPRB_REQUIRE(kernel == GPRB_KERNEL_RBF || kernel == GPRB_KERNEL_DOT, "gprb_kff: unknown kernel %d", kernel);
    GPRB_REQUIRE(mode >= 0 && mode <= 3, "gprb_kff: unknown mode %d", mode);
    GPRB_REQUIRE(0 <= grp_begin && grp_begin <= grp_end && grp_end <= f1->n_groups, "gprb_kff: bad window [%d,%d)", grp_begin, grp_end);
    GPRB_REQUIRE(!(dK && kernel == GPRB_KERNEL_DOT), "gprb_kff: Dot has no dK output (closed form, see header)");
    if (mode == GPRB_FF_SYMMETRIC)
        GPRB_REQUIRE(f1 == f2 && grp_begin == 0 && grp_end == f1->n_groups, "gprb_kff: symmetric mode needs f1 == f2 and the full window");
    if (mode == GPRB_FF_DIAG || mode == GPRB_FF_UPPER) GPRB_REQUIRE(f1 == f2, "gprb_kff: diag / upper mode needs f1 == f2");
    if (f1->ks > 2 * GPRB_MAX_KS) {
        gprb_set_error("gprb_kff: descriptor length %d > %d is not supported by the DMMA kernels", f1->d, 8 * GPRB_MAX_KS);
        return GPRB_ERR_UNSUPPORTED;
    }
    if (grp_begin == grp_end || f2->n_groups == 0) return GPRB_OK;
    int rc = build_sched(f1, grp_begin, grp_end, st);
    if (rc) return rc;
    // the output is pre-zeroed: groups without rows stay zero and row groups shared by two CTAs are
    // combined with atomics (UPPER leaves the blocks left of the diagonal zero)
    const int rows = 3 * (grp_end - grp_begin);
    if (mode == GPRB_FF_DIAG) {
        GPRB_REQUIRE(n_dst == 1, "gprb_kff_multi: diag mode has a single destination");
        GPRB_CUDA(cudaMemsetAsync(K, 0, (size_t)rows * sizeof(double), st));
    } else {
        if (n_dst == 1) GPRB_CUDA(cudaMemset2DAsync(K, ldk * sizeof(double), 0, (size_t)3 * f2->n_groups * sizeof(double), rows, st));
        if (dK) GPRB_CUDA(cudaMemset2DAsync(dK, lddk * sizeof(double), 0, (size_t)3 * f2->n_groups * sizeof(double), rows, st));
    }
    if (f1->sched_n == 0 || f2->n_rows == 0) return GPRB_OK;
    CovParams P = {};
    rc = fill_kernel_params(P, kernel, p0, p1, zeta);
    if (rc) return rc;
    P.PA = f1->P; P.eleA = f1->elep; P.row_groupA = f1->row_group; P.sched = f1->sched; P.sched_ent = f1->sched_ent;
    P.PB = f2->P; P.recB = f2->tile_rec; P.row_ptrB = f2->d_row_ptr; P.group_rowsB = f2->d_group_rows;
    P.n_groupsB = f2->n_groups; P.ks = f1->ks;
    P.win_r0 = f1->row_ptr[grp_begin]; P.win_r1 = f1->row_ptr[grp_end];
    P.tol = tol; P.use_tol = use_tol; P.mode = mode; P.grp_begin = grp_begin;
    P.K = K; P.ldk = ldk; P.dK = dK; P.lddk = lddk;
    P.n_extra = n_dst - 1;
    for (int p = 1; p < n_dst; p++) P.Kx[p - 1] = K_dst[p];
    // no-gradient K_ff with 8 k-steps (d = 29..32, the default descriptor): two-stage contraction, 1.44x the 4x4-block path
    // (profiles/r02_two_stage.txt); GPRB_KFF_TWO_STAGE=0 selects the block path for the A/B parity test
    const char *two_env = getenv("GPRB_KFF_TWO_STAGE");
    P.two_stage = (!dK && mode != GPRB_FF_DIAG && f1->ks == GPRB_MAX_KS && !(two_env && two_env[0] == '0')) ? 1 : 0;
    P.n_splits = mode == GPRB_FF_DIAG ? 1 : choose_splits(f1->sched_n, f2->n_groups);
    return dispatch_cov<4>(kernel, dK != nullptr, P, f1->sched_n, st);
}

extern "C" int gprb_kff(int kernel, const gprb_pack *f1, const gprb_pack *f2, double p0, double p1, double zeta,
                        int use_tol, double tol, int mode, int grp_begin, int grp_end,
                        double *K, long long ldk, double *dK, long long lddk, void *stream) {
    GPRB_REQUIRE(K, "gprb_kff: NULL argument");
    double *dst[1] = {K};
    return kff_impl(kernel, f1, f2, p0, p1, zeta, use_tol, tol, mode, grp_begin, grp_end, 1, dst, ldk, dK, lddk, (cudaStream_t)stream);
}

extern "C" int gprb_kff_multi(int kernel, const gprb_pack *f1, const gprb_pack *f2, double p0, double p1, double zeta,
                              int use_tol, double tol, int mode, int grp_begin, int grp_end,
                              int n_dst, double *const *K_dst_host, long long ldk, double *dK, long long lddk, void *stream) {
    GPRB_REQUIRE(mode == GPRB_FF_FULL || mode == GPRB_FF_UPPER, "gprb_kff_multi: mode must be FULL or UPPER");
    return kff_impl(kernel, f1, f2, p0, p1, zeta, use_tol, tol, mode, grp_begin, grp_end, n_dst, K_dst_host, ldk, dK, lddk,
                    (cudaStream_t)stream);
}

static int kef_impl(int kernel, const gprb_pack *e, const gprb_pack *f_, double p0, double p1, double zeta,
                    int grp_begin, int grp_end,
                    double *Kef, long long ld_ef, int n_dst, double *const *Kfe_dst, long long ld_fe,
                    double *dKef, long long ld_def, double *dKfe, long long ld_dfe, cudaStream_t st) {
    gprb_pack *f = const_cast<gprb_pack *>(f_);
    GPRB_REQUIRE(n_dst >= 1 && n_dst <= GPRB_MAX_DST && Kfe_dst, "gprb_kef: need 1..%d destination matrices", GPRB_MAX_DST);
    double *Kfe = Kfe_dst[0];
    for (int p = 1; p < n_dst; p++) GPRB_REQUIRE(Kfe_dst[p], "gprb_kfe_multi: NULL destination %d", p);
    GPRB_REQUIRE(e && f && (Kef || Kfe), "gprb_kef: NULL argument");
    { int rcd = gprb_check_device(e, "gprb_kef"); if (rcd || (rcd = gprb_check_device(f, "gprb_kef"))) return rcd; }
    GPRB_REQUIRE(e->ncols == 0 && f->ncols == 3, "gprb_kef: need (energy pack, force pack)");
    GPRB_REQUIRE(e->d == f->d, "gprb_kef: descriptor length mismatch %d vs %d", e->d, f->d);
    GPRB_REQUIRE(kernel == GPRB_KERNEL_RBF || kernel == GPRB_KERNEL_DOT, "gprb_kef: unknown kernel %d", kernel);
    GPRB_REQUIRE(0 <= grp_begin && grp_begin <= grp_end && grp_end <= f->n_groups, "gprb_kef: bad window [%d,%d)", grp_begin, grp_end);
    const bool grad = dKef || dKfe;
    GPRB_REQUIRE(!(grad && kernel == GPRB_KERNEL_DOT), "gprb_kef: Dot has no dK output (closed form, see header)");
    if (f->ks > 2 * GPRB_MAX_KS) {
        gprb_set_error("gprb_kef: descriptor length %d > %d is not supported by the DMMA kernels", f->d, 8 * GPRB_MAX_KS);
        return GPRB_ERR_UNSUPPORTED;
    }
    if (grp_begin == grp_end || e->n_groups == 0) return GPRB_OK;
    int rc = build_sched(f, grp_begin, grp_end, st);
    if (rc) return rc;
    const int rows = 3 * (grp_end - grp_begin);
    {
        const size_t w = (size_t)e->n_groups * sizeof(double);
        if (Kfe && n_dst == 1) GPRB_CUDA(cudaMemset2DAsync(Kfe, ld_fe * sizeof(double), 0, w, rows, st));
        if (dKfe) GPRB_CUDA(cudaMemset2DAsync(dKfe, ld_dfe * sizeof(double), 0, w, rows, st));
        if (Kef) GPRB_CUDA(cudaMemset2DAsync(Kef, ld_ef * sizeof(double), 0, (size_t)rows * sizeof(double), e->n_groups, st));
        if (dKef) GPRB_CUDA(cudaMemset2DAsync(dKef, ld_def * sizeof(double), 0, (size_t)rows * sizeof(double), e->n_groups, st));
    }
    if (f->sched_n == 0 || e->n_rows == 0) return GPRB_OK;
    CovParams P = {};
    rc = fill_kernel_params(P, kernel, p0, p1, zeta);
    if (rc) return rc;
    P.PA = f->P; P.eleA = f->elep; P.row_groupA = f->row_group; P.sched = f->sched; P.sched_ent = f->sched_ent;
    P.PB = e->P; P.recB = e->tile_rec; P.row_ptrB = e->d_row_ptr; P.group_rowsB = e->d_group_rows;
    P.n_groupsB = e->n_groups; P.ks = f->ks;
    P.win_r0 = f->row_ptr[grp_begin]; P.win_r1 = f->row_ptr[grp_end];
    P.mode = GPRB_FF_FULL; P.grp_begin = grp_begin;
    P.K = Kfe; P.ldk = ld_fe; P.dK = dKfe; P.lddk = ld_dfe;
    P.K2 = Kef; P.ldk2 = ld_ef; P.dK2 = dKef; P.lddk2 = ld_def;
    P.n_extra = n_dst - 1;
    for (int p = 1; p < n_dst; p++) P.Kx[p - 1] = Kfe_dst[p];
    P.n_splits = choose_splits(f->sched_n, e->n_groups);
    return dispatch_cov<1>(kernel, grad, P, f->sched_n, st);
}

extern "C" int gprb_kef(int kernel, const gprb_pack *e, const gprb_pack *f, double p0, double p1, double zeta,
                        int grp_begin, int grp_end,
                        double *Kef, long long ld_ef, double *Kfe, long long ld_fe,
                        double *dKef, long long ld_def, double *dKfe, long long ld_dfe, void *stream) {
    double *dst[1] = {Kfe};
    return kef_impl(kernel, e, f, p0, p1, zeta, grp_begin, grp_end, Kef, ld_ef, 1, dst, ld_fe, dKef, ld_def, dKfe, ld_dfe,
                    (cudaStream_t)stream);
}

extern "C" int gprb_kfe_multi(int kernel, const gprb_pack *e, const gprb_pack *f, double p0, double p1, double zeta,
                              int grp_begin, int grp_end, int n_dst, double *const *Kfe_dst_host, long long ld_fe,
                              double *dKfe, long long ld_dfe, void *stream) {
    GPRB_REQUIRE(Kfe_dst_host && n_dst >= 1 && Kfe_dst_host[0], "gprb_kfe_multi: NULL destination");
    return kef_impl(kernel, e, f, p0, p1, zeta, grp_begin, grp_end, nullptr, 0, n_dst, Kfe_dst_host, ld_fe, nullptr, 0, dKfe, ld_dfe,
                    (cudaStream_t)stream);
}

// K_ee on the same tile machinery (NB = 0: one component per side, 8 DMMAs per 8 x 8 pairs, table exp, tiles without a
// same-species pair skipped).  Replaces rbf_kee_many / rbf_kee_many_with_grad (rbf_kernel.cpp:5-98) and dot_kee_many
// (dot_kernel.cpp:5-56) with the wrappers' 1 / (n_I n_J) (rbf_kernel.py:56-70, dot_kernel.py:46).  Descriptors longer than
// 64 take the scalar kernel of cov_ee.cu.
int gprb_kee_scalar(int kernel, const gprb_pack *e1, const gprb_pack *e2, double p0, double p1, double zeta,
                    int grp_begin, int grp_end, double *K, long long ldk, double *dK, long long lddk, cudaStream_t st);

extern "C" int gprb_kee(int kernel, const gprb_pack *e1_, const gprb_pack *e2, double p0, double p1, double zeta,
                        int grp_begin, int grp_end, double *K, long long ldk, double *dK, long long lddk, void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    gprb_pack *e1 = const_cast<gprb_pack *>(e1_);
    GPRB_REQUIRE(e1 && e2 && K, "gprb_kee: NULL argument");
    { int rcd = gprb_check_device(e1, "gprb_kee"); if (rcd || (rcd = gprb_check_device(e2, "gprb_kee"))) return rcd; }
    GPRB_REQUIRE(e1->ncols == 0 && e2->ncols == 0, "gprb_kee: both sides must be energy packs");
    GPRB_REQUIRE(e1->d == e2->d, "gprb_kee: descriptor length mismatch %d vs %d", e1->d, e2->d);
    GPRB_REQUIRE(kernel == GPRB_KERNEL_RBF || kernel == GPRB_KERNEL_DOT, "gprb_kee: unknown kernel %d", kernel);
    GPRB_REQUIRE(0 <= grp_begin && grp_begin <= grp_end && grp_end <= e1->n_groups, "gprb_kee: bad window [%d,%d)", grp_begin, grp_end);
    GPRB_REQUIRE(!(dK && kernel == GPRB_KERNEL_DOT), "gprb_kee: Dot has no dK output (closed form, see header)");
    if (grp_begin == grp_end || e2->n_groups == 0) return GPRB_OK;
    if (e1->ks > 2 * GPRB_MAX_KS || getenv("GPRB_KEE_SCALAR") != nullptr)  // d > 64; env: A/B switch of the parity tests
        return gprb_kee_scalar(kernel, e1, e2, p0, p1, zeta, grp_begin, grp_end, K, ldk, dK, lddk, st);
    int rc = build_sched(e1, grp_begin, grp_end, st);
    if (rc) return rc;
    const int rows = grp_end - grp_begin;
    GPRB_CUDA(cudaMemset2DAsync(K, ldk * sizeof(double), 0, (size_t)e2->n_groups * sizeof(double), rows, st));
    if (dK) GPRB_CUDA(cudaMemset2DAsync(dK, lddk * sizeof(double), 0, (size_t)e2->n_groups * sizeof(double), rows, st));
    if (e1->sched_n == 0 || e2->n_rows == 0) return GPRB_OK;
    CovParams P = {};
    rc = fill_kernel_params(P, kernel, p0, p1, zeta);
    if (rc) return rc;
    P.PA = e1->P; P.eleA = e1->elep; P.row_groupA = e1->row_group; P.sched = e1->sched; P.sched_ent = e1->sched_ent;
    P.PB = e2->P; P.recB = e2->tile_rec; P.row_ptrB = e2->d_row_ptr; P.group_rowsB = e2->d_group_rows;
    P.group_rowsA = e1->d_group_rows;
    P.n_groupsB = e2->n_groups; P.ks = e1->ks;
    P.win_r0 = e1->row_ptr[grp_begin]; P.win_r1 = e1->row_ptr[grp_end];
    // the training block is symmetric: evaluate the J >= I group pairs once and write both entries
    P.mode = (e1 == e2 && grp_begin == 0 && grp_end == e1->n_groups) ? GPRB_FF_SYMMETRIC : GPRB_FF_FULL;
    P.grp_begin = grp_begin;
    P.K = K; P.ldk = ldk; P.dK = dK; P.lddk = lddk;
    P.n_splits = choose_splits(e1->sched_n, e2->n_groups);
    return dispatch_cov<0>(kernel, dK != nullptr, P, e1->sched_n, st);
}
