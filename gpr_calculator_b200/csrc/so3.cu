// so3.cu — SO(3) power-spectrum descriptor and its Cartesian derivative on device.
//
// Replaces SO3.build_neighbor_list (SO3.py:348-407, incl. the third-party ase NeighborList),
// compute_dcs (SO3.py:608-727: Gauss-Chebyshev radial quadrature x modified spherical Bessel x
// complex Y_lm x Gaussian x cosine cutoff, with gradients) and the per-centre loops of
// SO3.calculate (SO3.py:243-273).  Batched over many structures (atoms concatenated).
//
// Kernels (none of them is GEMM shaped; they are latency / special-function bound and tiny next
// to the covariance build — SURVEY.md §8a a11):
//   so3_neighbors_kernel   one warp per centre atom; ordered (j, image)-sorted compaction so the
//                          pair list and `seq` are deterministic and identical to the oracle's
//   so3_radial_kernel      one warp per neighbour: I_nl(r) = sum_q G[n,q] i_l(2 a r rho_q), dI/dr
//                          (i_l by Miller downward recurrence for z < 30, upward above)
//   so3_power_kernel       one CTA per centre: C_nlm = sum_w c_nlm(w); x_i = P[tril]; then per
//                          neighbour grad c_nlm -> dP, summed over the images of each j into
//                          dxdr[(i,j)], and dxdr[(i,i)] = - sum_{j != i}.
#include "common.cuh"
#include <cuComplex.h>

namespace {

constexpr double PI = 3.14159265358979323846;

struct SO3Geom {
    const int *atom_ptr;      // [S+1] first atom of each structure
    const int *struct_of;     // [n_atoms]
    const double *pos;        // [n_atoms,3]
    const double *cell;       // [S,9]
    const int *nimg;          // [S,3] images searched along each axis (0 if not periodic)
    double rcut;
};

// r = pos[j] + S.cell - pos[i], evaluated in the oracle's operation order without FMA contraction
// so that the strict `< rcut` test flips for exactly the same pairs (bit-exact neighbour lists).
__device__ __forceinline__ double pair_vec(const SO3Geom &g, int s, int i, int j, int sx, int sy, int sz, double *rv) {
    const double *c = g.cell + 9 * s;
    double d2 = 0.0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        double shift = __dadd_rn(__dadd_rn(__dmul_rn((double)sx, c[k]), __dmul_rn((double)sy, c[3 + k])), __dmul_rn((double)sz, c[6 + k]));
        double v = __dadd_rn(__dadd_rn(g.pos[3 * j + k], shift), -g.pos[3 * i + k]);
        rv[k] = v;
        d2 = __dadd_rn(d2, __dmul_rn(v, v));
    }
    return sqrt(d2);
}

// mode 0: count neighbours and unique neighbour atoms; mode 1: fill (nb_ptr given)
__global__ void so3_neighbors_kernel(SO3Geom g, int n_atoms, int mode, int *nnb, int *nuniq,
                                     const int *nb_ptr, int *nb_j, double *nb_rvec) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_atoms) return;
    const int i = warp, s = g.struct_of[i];
    const int a0 = g.atom_ptr[s], na = g.atom_ptr[s + 1] - a0;
    const int mx = g.nimg[3 * s], my = g.nimg[3 * s + 1], mz = g.nimg[3 * s + 2];
    const int nx = 2 * mx + 1, ny = 2 * my + 1, nz = 2 * mz + 1;
    const int nimg = nx * ny * nz;
    const long long ncand = (long long)na * nimg;
    int count = 0, uniq = 0, last_j = -1;
    bool self_seen = false;
    int out = mode ? nb_ptr[i] : 0;
    for (long long c0 = 0; c0 < ncand; c0 += 32) {
        const long long c = c0 + lane;
        bool hit = false;
        int j = 0;
        double rv[3] = {0, 0, 0};
        if (c < ncand) {
            const int jl = (int)(c / nimg), im = (int)(c - (long long)jl * nimg);
            const int sx = im / (ny * nz) - mx, sy = (im / nz) % ny - my, sz = im % nz - mz;   // images sorted (sx, sy, sz)
            j = a0 + jl;
            const double dist = pair_vec(g, s, i, j, sx, sy, sz, rv);
            hit = dist < g.rcut && !(j == i && sx == 0 && sy == 0 && sz == 0);
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (mode == 0) {
            // unique neighbour atoms: candidates are j-major, so count transitions of j among hits
            unsigned mm = m;
            while (mm) {
                const int b = __ffs(mm) - 1;
                mm &= mm - 1;
                const int jj = __shfl_sync(0xffffffffu, j, b);
                if (jj != last_j) { uniq++; last_j = jj; if (jj == i) self_seen = true; }
            }
            count += __popc(m);
        } else {
            if (hit) {
                const int o = out + __popc(m & ((1u << lane) - 1));
                nb_j[o] = j;
                nb_rvec[3 * o] = rv[0]; nb_rvec[3 * o + 1] = rv[1]; nb_rvec[3 * o + 2] = rv[2];
            }
            out += __popc(m);
        }
    }
    if (mode == 0 && lane == 0) { nnb[i] = count; nuniq[i] = uniq + (self_seen ? 0 : 1); }
}

struct SO3Params {
    int nmax, lmax, nq;
    double alpha, rcut;
    const double *rho;      // [nq]
    const double *G;        // [nmax, nq]
    const double *norm_l;   // [lmax+1]
};

// modified spherical Bessel functions of the first kind i_0..i_L at z > 0
__device__ void sph_in(double z, int L, double *out) {
    const double inv = 1.0 / z;
    if (z >= 30.0) {
        const double sh = sinh(z), ch = cosh(z);
        out[0] = sh * inv;
        if (L >= 1) out[1] = (ch - sh * inv) * inv;
        for (int l = 1; l < L; l++) out[l + 1] = out[l - 1] - (2 * l + 1) * inv * out[l];
        return;
    }
    double fp1 = 0.0, f = 1.0;
    for (int l = L + 40; l >= 1; l--) {
        const double fm1 = fp1 + (2 * l + 1) * inv * f;
        fp1 = f; f = fm1;
        if (fabs(f) > 1e150) {
            f *= 1e-150; fp1 *= 1e-150;
            for (int k = l; k <= L; k++) out[k] *= 1e-150;
        }
        if (l - 1 <= L) out[l - 1] = f;
    }
    const double scale = sinh(z) * inv / out[0];
    for (int l = 0; l <= L; l++) out[l] *= scale;
}

constexpr int SO3_MAXL = 16;   // supported lmax (+1 for the gradient) <= 16

// rad[w] = { I[n][l] , dI/dr[n][l] }
// NMAX_T / LMAX_T > 0: compile-time sizes (the default descriptor nmax = 3, lmax = 4): the 2 nmax (lmax+1) partial sums of a lane
// stay in registers.  The generic instantiation (0, 0) indexes them dynamically, i.e. keeps them in local memory (ncu of round 2:
// 310 MB of DRAM traffic per 64 structures for 42 MB of output, long-scoreboard 42 %).
template <int NMAX_T, int LMAX_T>
__global__ void so3_radial_kernel(int n_nb, const double *__restrict__ nb_rvec, SO3Params p, double *__restrict__ rad) {
    extern __shared__ double sm[];   // per warp: 2*nmax*(lmax+1)
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + wib;
    const int nmax = NMAX_T ? NMAX_T : p.nmax, lmax = NMAX_T ? LMAX_T : p.lmax;
    const int L1 = lmax + 1, nnl = nmax * L1;
    if (w >= n_nb) return;
    double *acc = sm + (size_t)wib * 2 * nnl;
    for (int k = lane; k < 2 * nnl; k += 32) acc[k] = 0.0;
    __syncwarp();
    const double rx = nb_rvec[3 * w], ry = nb_rvec[3 * w + 1], rz = nb_rvec[3 * w + 2];
    const double r = sqrt(rx * rx + ry * ry + rz * rz);
    // each lane owns quadrature points q = lane, lane+32, ...; partial sums are combined lane by lane
    // in a fixed order (deterministic)
    constexpr int NPART = NMAX_T ? 2 * NMAX_T * (LMAX_T + 1) : 2 * 6 * (SO3_MAXL + 1);   // generic: nmax <= 6 in the fast path
    double part[NPART];
    const bool small = NMAX_T ? true : (nmax <= 6);
    if (small) {
#pragma unroll
        for (int k = 0; k < NPART; k++) part[k] = 0.0;
    }
    for (int q = lane; q < p.nq; q += 32) {
        const double rho = p.rho[q];
        const double z = 2.0 * p.alpha * r * rho;
        double il[(NMAX_T ? LMAX_T : SO3_MAXL) + 2];
        const int Lb = lmax > 1 ? lmax : 1;
        sph_in(z, Lb, il);
        if (NMAX_T) {
#pragma unroll
            for (int l = 0; l <= LMAX_T; l++) {
                const double dil = (l == 0) ? il[1] : il[l > 0 ? l - 1 : 0] - (l + 1) / z * il[l];
                const double dz = dil * 2.0 * p.alpha * rho;
#pragma unroll
                for (int n = 0; n < (NMAX_T ? NMAX_T : 1); n++) {
                    const double gq = p.G[n * p.nq + q];
                    part[n * (LMAX_T + 1) + l] += gq * il[l];
                    part[(NMAX_T * (LMAX_T + 1)) + n * (LMAX_T + 1) + l] += gq * dz;
                }
            }
        } else {
            for (int l = 0; l <= lmax; l++) {
                const double dil = (l == 0) ? il[1] : il[l - 1] - (l + 1) / z * il[l];
                const double dz = dil * 2.0 * p.alpha * rho;
                for (int n = 0; n < nmax; n++) {
                    const double gq = p.G[n * p.nq + q];
                    if (small) { part[n * L1 + l] += gq * il[l]; part[nnl + n * L1 + l] += gq * dz; }
                    else { atomicAdd(&acc[n * L1 + l], gq * il[l]); atomicAdd(&acc[nnl + n * L1 + l], gq * dz); }
                }
            }
        }
    }
    if (small) {
#pragma unroll
        for (int k = 0; k < NPART; k++) {
            if (k >= 2 * nnl) break;
            double v = part[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) acc[k] = v;
        }
    }
    __syncwarp();
    for (int k = lane; k < 2 * nnl; k += 32) rad[(size_t)w * 2 * nnl + k] = acc[k];
}

// Normalised Y_lm for 0 <= m <= l <= L into Y[(l*(L+1)+m)] (complex), Condon-Shortley phase,
// the convention of scipy.special.sph_harm_y used by the reference (SO3.py:679).
// One lane per m; ct = cos(theta), st = sin(theta), (cp, sp) = (cos phi, sin phi).  The square-root factors of the
// recurrences come from tables built once per CTA (ylm_tables): tp[k] = -sqrt((2k+1)/(2k)), ta / tb [l*(L+1)+m].
__device__ void ylm_column(int m, int L, double ct, double st, double cp, double sp, const double *tp, const double *ta,
                           const double *tb, cuDoubleComplex *Y) {
    double pmm = 0.28209479177387814347;      // sqrt(1 / 4 pi)
    for (int k = 1; k <= m; k++) pmm *= tp[k] * st;
    double cm = 1.0, sm_ = 0.0;                 // e^{i m phi}
    for (int k = 0; k < m; k++) { const double t = cm * cp - sm_ * sp; sm_ = sm_ * cp + cm * sp; cm = t; }
    double pl2 = 0.0, pl1 = pmm;
    Y[m * (L + 1) + m] = make_cuDoubleComplex(pmm * cm, pmm * sm_);
    for (int l = m + 1; l <= L; l++) {
        const double pl = (l == m + 1) ? ta[l * (L + 1) + m] * ct * pl1 : ta[l * (L + 1) + m] * (ct * pl1 - tb[l * (L + 1) + m] * pl2);
        Y[l * (L + 1) + m] = make_cuDoubleComplex(pl * cm, pl * sm_);
        pl2 = pl1; pl1 = pl;
    }
}

__device__ void ylm_tables(int L, int tid, int nt, double *tp, double *ta, double *tb) {
    for (int k = tid; k <= L; k += nt) tp[k] = k ? -sqrt((2.0 * k + 1.0) / (2.0 * k)) : 0.0;
    for (int e = tid; e < (L + 1) * (L + 1); e += nt) {
        const int l = e / (L + 1), m = e % (L + 1);
        double av = 0.0, bv = 0.0;
        if (l == m + 1) av = sqrt(2.0 * m + 3.0);
        else if (l > m + 1) {
            av = sqrt((4.0 * l * l - 1.0) / ((double)l * l - (double)m * m));
            bv = sqrt((((double)l - 1.0) * (l - 1.0) - (double)m * m) / (4.0 * (l - 1.0) * (l - 1.0) - 1.0));
        }
        ta[e] = av; tb[e] = bv;
    }
}

__device__ __forceinline__ cuDoubleComplex ylm_get(const cuDoubleComplex *Y, int L, int l, int m) {
    if (l < 0 || l > L || m > l || m < -l) return make_cuDoubleComplex(0.0, 0.0);
    if (m >= 0) return Y[l * (L + 1) + m];
    const cuDoubleComplex v = Y[l * (L + 1) - m];
    const double sgn = ((-m) & 1) ? -1.0 : 1.0;        // Y_l^{-m} = (-1)^m conj(Y_l^m)
    return make_cuDoubleComplex(sgn * v.x, -sgn * v.y);
}

struct SO3Power {
    const int *nb_ptr; const int *nb_j; const double *nb_rvec; const double *rad;
    const int *numbers; const int *atom_ptr; const int *struct_of; const int *seq_ptr;
    double *x; double *dxdr; long long *seq;
    int derivative;
    const double *pos; const double *inv_vol; double *rdxdr;   // stress: rdxdr[n_seq][d][3][3] = -pstress / volume
};

// Shared-memory plan of so3_power_kernel (doubles unless noted), W = warps per CTA:
//   CTA:      sC [nent] complex | coefficient tables: gradient of Y_lm [L1*M*6], Y recurrences tp [LY+1], ta, tb [(LY+1)^2]
//   per warp: sdc [3 nent] complex (phase A: the warp's partial C_nlm) | sY [(LY+1)^2] complex | sAcc, sSelf [3 d] |
//             stress: sAccR, sSelfR [9 d], sTot [3 d]
__host__ __device__ inline size_t so3_power_cta_doubles(int nent, int L1, int M, int LY) {
    const size_t n = (size_t)2 * nent + (size_t)6 * L1 * M + (size_t)(LY + 1) + (size_t)2 * (LY + 1) * (LY + 1);
    return (n + 1) & ~(size_t)1;          // the per-warp regions start with complex numbers: keep them 16-byte aligned
}
__host__ __device__ inline size_t so3_power_warp_doubles(int nent, int d, int LY, bool stress) {
    const size_t n = (size_t)6 * nent + (size_t)2 * (LY + 1) * (LY + 1) + (size_t)6 * d + (stress ? (size_t)21 * d : 0);
    return (n + 1) & ~(size_t)1;
}

// One CTA per centre atom, W warps; every WARP owns whole neighbours (phase A) / whole groups of images of one neighbour
// atom j (phase B) and works through them without CTA-wide barriers: the geometry scalars are computed redundantly by all
// lanes, Y_lm by lanes 0..LY into the warp's own table, grad c_nlm by the lanes into the warp's own buffer, and every lane
// owns fixed outputs of dP.  The square roots of the Y_lm / grad Y_lm recurrences are tabulated once per CTA.  CTA-wide
// barriers: three (tables + zeroing, C_nlm complete, self rows complete).  (Round 1: one neighbour at a time for the
// whole CTA with three barriers each and six square roots per c_nlm entry: 48 % of the stall samples were barriers.)
__global__ void __launch_bounds__(128, 4) so3_power_kernel(SO3Power a, SO3Params p) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int L1 = p.lmax + 1, M = 2 * p.lmax + 1, LY = p.lmax + 1;   // Y table up to l = lmax+1
    const int nent = p.nmax * L1 * M;
    const int npair = p.nmax * (p.nmax + 1) / 2, d = npair * L1;
    const bool stress = a.rdxdr != nullptr;
    const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31, W = nt >> 5;
    double *base = reinterpret_cast<double *>(raw);
    cuDoubleComplex *sC = reinterpret_cast<cuDoubleComplex *>(base);               // [nent] C_tot
    double *sCoef = base + 2 * nent;                                               // [L1*M][6]
    double *sTp = sCoef + 6 * L1 * M;                                              // [LY+1]
    double *sTa = sTp + (LY + 1), *sTb = sTa + (LY + 1) * (LY + 1);
    double *wbase = base + so3_power_cta_doubles(nent, L1, M, LY) + (size_t)warp * so3_power_warp_doubles(nent, d, LY, stress);
    cuDoubleComplex *sdc = reinterpret_cast<cuDoubleComplex *>(wbase);             // [nent][3] grad c(w) of the warp's neighbour
    cuDoubleComplex *sY = sdc + 3 * nent;                                          // [(LY+1)^2]
    double *sAcc = reinterpret_cast<double *>(sY + (LY + 1) * (LY + 1));           // [d*3] current j group
    double *sSelf = sAcc + 3 * d;                                                  // [d*3] the warp's sum over its groups j != i
    double *sAccR = sSelf + 3 * d;                                                 // [d*9] R_j (x) dP of the current j group
    double *sSelfR = sAccR + 9 * d;                                                // [d*9] the j == i group (own images)
    double *sTot = sSelfR + 9 * d;                                                 // [d*3] the warp's sum of dP over its neighbours
    const size_t wstride = so3_power_warp_doubles(nent, d, LY, stress);
    const int i = blockIdx.x;
    const int w0 = a.nb_ptr[i], w1 = a.nb_ptr[i + 1];
    const int a0 = a.atom_ptr[a.struct_of[i]];
    const int nnl = p.nmax * L1;

    // ---- per-CTA tables -----------------------------------------------------------------------
    ylm_tables(LY, tid, nt, sTp, sTa, sTb);
    for (int e = tid; e < L1 * M; e += nt) {
        // covariant spherical components of grad Y_lm (SO3.py:686-702): coefficients of Y_{l+1} and Y_{l-1}, without 1/r
        const int l = e / M, m = e % M - p.lmax;
        double *c = sCoef + 6 * e;
        for (int k = 0; k < 6; k++) c[k] = 0.0;
        if (l >= 1 && m >= -l && m <= l) {
            const double dl = (double)l, dm = (double)m;
            c[0] = -sqrt(((dl + 1) * (dl + 1) - dm * dm) / (2 * dl + 1) / (2 * dl + 3)) * dl;
            if (abs(m) <= l - 1) c[1] = sqrt((dl * dl - dm * dm) / (2 * dl - 1) / (2 * dl + 1)) * (dl + 1);
            c[2] = -sqrt((dl + dm + 1) * (dl + dm + 2) / 2 / (2 * dl + 1) / (2 * dl + 3)) * dl;
            if (abs(m + 1) <= l - 1) c[3] = sqrt((dl - dm - 1) * (dl - dm) / 2 / (2 * dl - 1) / (2 * dl + 1)) * (dl + 1);
            c[4] = -sqrt((dl - dm + 1) * (dl - dm + 2) / 2 / (2 * dl + 1) / (2 * dl + 3)) * dl;
            if (abs(m - 1) <= l - 1) c[5] = sqrt((dl + dm - 1) * (dl + dm) / 2 / (2 * dl - 1) / (2 * dl + 1)) * (dl + 1);
        }
    }
    for (int e = lane; e < 3 * nent; e += 32) sdc[e] = make_cuDoubleComplex(0.0, 0.0);      // phase A: partial C in sdc[0..nent)
    for (int o = lane; o < 3 * d; o += 32) { sAcc[o] = 0.0; sSelf[o] = 0.0; }
    if (stress) {
        for (int o = lane; o < 9 * d; o += 32) { sAccR[o] = 0.0; sSelfR[o] = 0.0; }
        for (int o = lane; o < 3 * d; o += 32) sTot[o] = 0.0;
    }
    __syncthreads();

    // geometry of neighbour w: every lane computes the scalars, lanes 0..LY the columns of Y_lm (warp-local)
    double g_r, g_ux, g_uy, g_uz, g_gauss, g_dgauss, g_fc, g_dfc, g_Z;
    auto geometry = [&](int w) {
        const double rx = a.nb_rvec[3 * w], ry = a.nb_rvec[3 * w + 1], rz = a.nb_rvec[3 * w + 2];
        const double r = sqrt(rx * rx + ry * ry + rz * rz);
        const double rxy = sqrt(rx * rx + ry * ry);
        const double ct = rz / r, st = rxy / r;
        const double cp = rxy > 0.0 ? rx / rxy : 1.0, sp = rxy > 0.0 ? ry / rxy : 0.0;
        __syncwarp();                       // the previous neighbour's readers of sY are done
        if (lane <= LY) ylm_column(lane, LY, ct, st, cp, sp, sTp, sTa, sTb, sY);
        g_r = r; g_ux = rx / r; g_uy = ry / r; g_uz = rz / r;
        g_gauss = 4.0 * PI * exp(-p.alpha * r * r);
        g_dgauss = -2.0 * p.alpha * r * g_gauss;
        g_fc = 0.5 * (cos(PI * r / p.rcut) + 1.0);
        g_dfc = -0.5 * PI / p.rcut * sin(PI * r / p.rcut);
        // neighbour weight Z_j; weight_on: a neighbour of another species counts negative (SO3.py:381-385)
        const int zj = a.numbers[a.nb_j[w]];
        g_Z = ((a.derivative & 2) && zj != a.numbers[i]) ? -(double)zj : (double)zj;
        __syncwarp();
    };

    // ---- phase A: C_tot = sum_w Z_j N_l 4pi e^{-a r^2} f_c Y_lm I_nl  (warp w % W, partial sums per warp) ----------
    for (int w = w0 + warp; w < w1; w += W) {
        geometry(w);
        const double pref = g_Z * g_gauss * g_fc;
        const double *I = a.rad + (size_t)w * 2 * nnl;
        for (int e = lane; e < nent; e += 32) {
            const int n = e / (L1 * M), l = (e / M) % L1, m = e % M - p.lmax;
            if (m < -l || m > l) continue;
            const cuDoubleComplex y = ylm_get(sY, LY, l, m);
            const double f = pref * p.norm_l[l] * I[n * L1 + l];
            sdc[e].x += f * y.x; sdc[e].y += f * y.y;
        }
    }
    __syncthreads();
    for (int e = tid; e < nent; e += nt) {           // fixed order over the warps: deterministic
        double cx = 0.0, cy = 0.0;
        for (int k = 0; k < W; k++) {
            const cuDoubleComplex v = reinterpret_cast<const cuDoubleComplex *>(base + so3_power_cta_doubles(nent, L1, M, LY) + k * wstride)[e];
            cx += v.x; cy += v.y;
        }
        sC[e] = make_cuDoubleComplex(cx, cy);
    }
    __syncthreads();
    // x_i = Re sum_m C_nlm conj(C_n'lm), tril(n >= n') x l   (SO3.py:248, 266)
    for (int o = tid; o < d; o += nt) {
        const int pr = o / L1, l = o % L1;
        int n = 0; while ((n + 1) * (n + 2) / 2 <= pr) n++;
        const int n2 = pr - n * (n + 1) / 2;
        double s = 0.0;
        for (int m = -l; m <= l; m++) {
            const cuDoubleComplex c1 = sC[(n * L1 + l) * M + m + p.lmax], c2 = sC[(n2 * L1 + l) * M + m + p.lmax];
            s += c1.x * c2.x + c1.y * c2.y;
        }
        a.x[(size_t)i * d + o] = s;
    }
    if (!(a.derivative & 1)) return;

    // ---- phase B: dP per neighbour, grouped by neighbour atom j; group g belongs to warp g % W -------------------
    // rows of `seq` for centre i: the unique neighbour atoms in increasing order, the (i, i) row at its sorted position
    int n_less = 0;
    bool has_self = false;
    {
        int prev = -1;
        for (int w = w0; w < w1; w++) {
            const int j = a.nb_j[w];
            if (j != prev) { prev = j; if (j < i) n_less++; if (j == i) has_self = true; }
        }
    }
    const int row0 = a.seq_ptr[i];
    const int self_row = row0 + n_less;
    const double isq2 = 0.70710678118654752440;
    const double iv = stress ? a.inv_vol[a.struct_of[i]] : 0.0;
    int g = -1, prev_j = -1;
    for (int w = w0; w < w1; w++) {
        const int j = a.nb_j[w];
        if (j != prev_j) { g++; prev_j = j; }
        if (g % W != warp) continue;
        geometry(w);
        const double r = g_r, ir = 1.0 / g_r, gauss = g_gauss, dgauss = g_dgauss, fc = g_fc, dfc = g_dfc, Z = g_Z;
        const double u[3] = {g_ux, g_uy, g_uz};
        const double *I = a.rad + (size_t)w * 2 * nnl;
        const double *dI = I + nnl;
        (void)r;
        for (int e = lane; e < nent; e += 32) {
            const int n = e / (L1 * M), l = (e / M) % L1, m = e % M - p.lmax;
            cuDoubleComplex g0 = make_cuDoubleComplex(0, 0), g1 = g0, g2 = g0;
            if (m >= -l && m <= l) {
                const cuDoubleComplex y = ylm_get(sY, LY, l, m);
                cuDoubleComplex c0 = make_cuDoubleComplex(0, 0), cpl = c0, cmi = c0;
                if (l >= 1) {
                    const double *cf = sCoef + 6 * (l * M + m + p.lmax);
                    const cuDoubleComplex yu0 = ylm_get(sY, LY, l + 1, m), yd0 = ylm_get(sY, LY, l - 1, m);
                    const cuDoubleComplex yup = ylm_get(sY, LY, l + 1, m + 1), ydp = ylm_get(sY, LY, l - 1, m + 1);
                    const cuDoubleComplex yum = ylm_get(sY, LY, l + 1, m - 1), ydm = ylm_get(sY, LY, l - 1, m - 1);
                    c0 = make_cuDoubleComplex((cf[0] * yu0.x + cf[1] * yd0.x) * ir, (cf[0] * yu0.y + cf[1] * yd0.y) * ir);
                    cpl = make_cuDoubleComplex((cf[2] * yup.x - cf[3] * ydp.x) * ir, (cf[2] * yup.y - cf[3] * ydp.y) * ir);
                    cmi = make_cuDoubleComplex((cf[4] * yum.x - cf[5] * ydm.x) * ir, (cf[4] * yum.y - cf[5] * ydm.y) * ir);
                }
                // Cartesian gradient of Y (SO3.py:705-707): x = (c- - c+)/sqrt2, y = i (c- + c+)/sqrt2, z = c0
                const cuDoubleComplex gy[3] = {make_cuDoubleComplex((cmi.x - cpl.x) * isq2, (cmi.y - cpl.y) * isq2),
                                               make_cuDoubleComplex(-(cmi.y + cpl.y) * isq2, (cmi.x + cpl.x) * isq2), c0};
                const double Inl = I[n * L1 + l], dInl = dI[n * L1 + l];
                const double wl = Z * p.norm_l[l];
                cuDoubleComplex out[3];
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    // d/dx_k [gauss * fc * Y * I]   (SO3.py:718-725)
                    const double sc = (dgauss * u[k] * Inl + gauss * dInl * u[k]) * fc + dfc * u[k] * gauss * Inl;
                    out[k].x = wl * (sc * y.x + gauss * fc * Inl * gy[k].x);
                    out[k].y = wl * (sc * y.y + gauss * fc * Inl * gy[k].y);
                }
                g0 = out[0]; g1 = out[1]; g2 = out[2];
            }
            sdc[3 * e] = g0; sdc[3 * e + 1] = g1; sdc[3 * e + 2] = g2;
        }
        __syncwarp();
        // dP[pair(n,n'), l, k] = Re sum_m [ dc_nlm conj(C_n'lm) + conj(dc_n'lm conj(C_nlm)) ]  (SO3.py:249-251)
        for (int o = lane; o < 3 * d; o += 32) {
            const int k = o % 3, pl = o / 3, pr = pl / L1, l = pl % L1;
            int n = 0; while ((n + 1) * (n + 2) / 2 <= pr) n++;
            const int n2 = pr - n * (n + 1) / 2;
            double s = 0.0;
            for (int m = -l; m <= l; m++) {
                const int e1 = (n * L1 + l) * M + m + p.lmax, e2 = (n2 * L1 + l) * M + m + p.lmax;
                const cuDoubleComplex d1 = sdc[3 * e1 + k], d2 = sdc[3 * e2 + k];
                const cuDoubleComplex c1 = sC[e1], c2 = sC[e2];
                s += d1.x * c2.x + d1.y * c2.y + d2.x * c1.x + d2.y * c1.y;
            }
            sAcc[o] += s;
            if (stress) {      // pstress[(i,j)] -= R_j (x) dP(w), R_j = r_i + r_ij (SO3.py:226, 254, 264); the lane owns (pl, :, k)
                sTot[o] += s;
#pragma unroll
                for (int n3 = 0; n3 < 3; n3++)
                    sAccR[(pl * 3 + n3) * 3 + k] += (a.pos[3 * (size_t)i + n3] + a.nb_rvec[3 * (size_t)w + n3]) * s;
            }
        }
        __syncwarp();
        // flush when the next neighbour belongs to another atom
        const bool last_of_j = (w + 1 == w1) || (a.nb_j[w + 1] != j);
        if (!last_of_j) continue;
        if (j == i) {
            for (int o = lane; o < 3 * d; o += 32) sAcc[o] = 0.0;        // own images cancel (SO3.py:267-273)
            if (stress) for (int o = lane; o < 9 * d; o += 32) { sSelfR[o] = sAccR[o]; sAccR[o] = 0.0; }
        } else {
            const int rj = row0 + g + ((!has_self && j > i) ? 1 : 0);
            for (int o = lane; o < 3 * d; o += 32) {
                const double v = sAcc[o];
                a.dxdr[(size_t)rj * 3 * d + o] = v;
                sSelf[o] += v;
                sAcc[o] = 0.0;
            }
            if (stress) for (int o = lane; o < 9 * d; o += 32) { a.rdxdr[(size_t)rj * 9 * d + o] = sAccR[o] * iv; sAccR[o] = 0.0; }
            if (lane == 0) { a.seq[2 * (size_t)rj] = i - a0; a.seq[2 * (size_t)rj + 1] = j - a0; }
        }
        __syncwarp();
    }
    __syncthreads();
    // the (i, i) row: minus the sum over all j != i, combined over the warps in a fixed order
    double *w0base = base + so3_power_cta_doubles(nent, L1, M, LY);
    const size_t offSelf = (size_t)6 * nent + (size_t)2 * (LY + 1) * (LY + 1) + (size_t)3 * d;
    for (int o = tid; o < 3 * d; o += nt) {
        double s = 0.0;
        for (int k = 0; k < W; k++) s += (w0base + k * wstride + offSelf)[o];
        a.dxdr[(size_t)self_row * 3 * d + o] = -s;
    }
    if (stress) {      // pstress[(i,i)] = -sum_{own images} R_j (x) dP + R_i (x) sum_w dP ; rdxdr = -pstress / vol
        const size_t offSelfR = offSelf + (size_t)3 * d + (size_t)9 * d, offTot = offSelfR + (size_t)9 * d;
        for (int o = tid; o < 9 * d; o += nt) {
            const int k = o % 3, n3 = (o / 3) % 3, pl = o / 9;
            double sr = 0.0, stot = 0.0;
            for (int q = 0; q < W; q++) { sr += (w0base + q * wstride + offSelfR)[o]; stot += (w0base + q * wstride + offTot)[pl * 3 + k]; }
            a.rdxdr[(size_t)self_row * 9 * d + o] = (sr - a.pos[3 * (size_t)i + n3] * stot) * iv;
        }
    }
    if (tid == 0) { a.seq[2 * (size_t)self_row] = i - a0; a.seq[2 * (size_t)self_row + 1] = i - a0; }
}

}  // namespace

extern "C" int gprb_so3_neighbors(int n_struct, int n_atoms, const int *atom_ptr, const int *struct_of,
                                  const double *pos, const double *cell, const int *nimg, double rcut,
                                  int mode, int *nnb, int *nuniq, const int *nb_ptr, int *nb_j, double *nb_rvec,
                                  void *stream) {
    GPRB_REQUIRE(n_struct >= 0 && n_atoms >= 0 && rcut > 0, "gprb_so3_neighbors: bad sizes");
    if (n_atoms == 0) return GPRB_OK;
    GPRB_REQUIRE(atom_ptr && struct_of && pos && cell && nimg, "gprb_so3_neighbors: NULL input");
    if (mode == 0) GPRB_REQUIRE(nnb && nuniq, "gprb_so3_neighbors: NULL count outputs");
    else GPRB_REQUIRE(nb_ptr && nb_j && nb_rvec, "gprb_so3_neighbors: NULL fill outputs");
    SO3Geom g{atom_ptr, struct_of, pos, cell, nimg, rcut};
    const int wpb = 4;
    so3_neighbors_kernel<<<(n_atoms + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(g, n_atoms, mode, nnb, nuniq, nb_ptr, nb_j, nb_rvec);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    return GPRB_OK;
}

extern "C" int gprb_so3_radial(int n_nb, const double *nb_rvec, int nmax, int lmax, int nq, double alpha, double rcut,
                               const double *rho, const double *G, double *rad, void *stream) {
    if (n_nb == 0) return GPRB_OK;
    GPRB_REQUIRE(nb_rvec && rho && G && rad, "gprb_so3_radial: NULL argument");
    GPRB_REQUIRE(nmax >= 1 && lmax >= 0 && lmax <= SO3_MAXL - 1, "gprb_so3_radial: need nmax >= 1 and 0 <= lmax <= %d", SO3_MAXL - 1);
    SO3Params p{nmax, lmax, nq, alpha, rcut, rho, G, nullptr};
    const int wpb = 4;
    const size_t smem = (size_t)wpb * 2 * nmax * (lmax + 1) * sizeof(double);
    if (nmax == 3 && lmax == 4)      // the reference's default descriptor (gaussianprocess.py:1028): register-resident partial sums
        so3_radial_kernel<3, 4><<<(n_nb + wpb - 1) / wpb, wpb * 32, smem, (cudaStream_t)stream>>>(n_nb, nb_rvec, p, rad);
    else
        so3_radial_kernel<0, 0><<<(n_nb + wpb - 1) / wpb, wpb * 32, smem, (cudaStream_t)stream>>>(n_nb, nb_rvec, p, rad);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    return GPRB_OK;
}

extern "C" int gprb_so3_power(int n_atoms, const int *nb_ptr, const int *nb_j, const double *nb_rvec, const double *rad,
                              const int *numbers, const int *atom_ptr, const int *struct_of, const int *seq_ptr,
                              int nmax, int lmax, double alpha, double rcut, const double *norm_l, int derivative,
                              double *x, double *dxdr, long long *seq,
                              const double *pos, const double *inv_vol, double *rdxdr, void *stream) {
    if (n_atoms == 0) return GPRB_OK;
    GPRB_REQUIRE(nb_ptr && numbers && atom_ptr && struct_of && norm_l && x, "gprb_so3_power: NULL argument");
    GPRB_REQUIRE(!(derivative & 1) || (seq_ptr && dxdr && seq), "gprb_so3_power: derivative outputs missing");
    GPRB_REQUIRE(nmax >= 1 && lmax >= 0 && lmax <= SO3_MAXL - 1, "gprb_so3_power: need nmax >= 1 and 0 <= lmax <= %d", SO3_MAXL - 1);
    SO3Params p{nmax, lmax, 0, alpha, rcut, nullptr, nullptr, norm_l};
    GPRB_REQUIRE(!rdxdr || ((derivative & 1) && pos && inv_vol), "gprb_so3_power: stress output needs derivative, pos and inv_vol");
    SO3Power a{nb_ptr, nb_j, nb_rvec, rad, numbers, atom_ptr, struct_of, seq_ptr, x, dxdr, seq, derivative, pos, inv_vol, rdxdr};
    const int L1 = lmax + 1, M = 2 * lmax + 1, LY = lmax + 1;
    const int nent = nmax * L1 * M, d = nmax * (nmax + 1) / 2 * L1;
    const size_t cta_b = so3_power_cta_doubles(nent, L1, M, LY) * sizeof(double);
    const size_t warp_b = so3_power_warp_doubles(nent, d, LY, rdxdr != nullptr) * sizeof(double);
    int warps = 4;                                   // warps per centre atom (four CTAs per SM); fewer if one CTA would not fit 200 KB
    while (warps > 1 && cta_b + warps * warp_b > 200 * 1024) warps--;
    const size_t smem = cta_b + warps * warp_b;
    GPRB_REQUIRE(smem <= 200 * 1024, "gprb_so3_power: nmax=%d lmax=%d needs %zu bytes of shared memory", nmax, lmax, smem);
    static size_t configured[64] = {};
    int dev = 0;
    GPRB_CUDA(cudaGetDevice(&dev));
    if (smem > 48 * 1024 && dev >= 0 && dev < 64 && smem > configured[dev]) {
        GPRB_CUDA(cudaFuncSetAttribute(so3_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[dev] = smem;
    }
    so3_power_kernel<<<n_atoms, warps * 32, smem, (cudaStream_t)stream>>>(a, p);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    return GPRB_OK;
}
