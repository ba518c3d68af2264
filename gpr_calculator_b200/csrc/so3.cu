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
__global__ void so3_radial_kernel(int n_nb, const double *__restrict__ nb_rvec, SO3Params p, double *__restrict__ rad) {
    extern __shared__ double sm[];   // per warp: 2*nmax*(lmax+1)
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + wib;
    const int L1 = p.lmax + 1, nnl = p.nmax * L1;
    if (w >= n_nb) return;
    double *acc = sm + (size_t)wib * 2 * nnl;
    for (int k = lane; k < 2 * nnl; k += 32) acc[k] = 0.0;
    __syncwarp();
    const double rx = nb_rvec[3 * w], ry = nb_rvec[3 * w + 1], rz = nb_rvec[3 * w + 2];
    const double r = sqrt(rx * rx + ry * ry + rz * rz);
    // each lane owns quadrature points q = lane, lane+32, ...; partial sums are combined lane by lane
    // in a fixed order (deterministic)
    double part[2 * 6 * (SO3_MAXL + 1)];   // nmax <= 6 in the fast path; larger nmax handled below
    const bool small = p.nmax <= 6;
    if (small) for (int k = 0; k < 2 * nnl; k++) part[k] = 0.0;
    for (int q = lane; q < p.nq; q += 32) {
        const double rho = p.rho[q];
        const double z = 2.0 * p.alpha * r * rho;
        double il[SO3_MAXL + 2];
        const int Lb = p.lmax > 1 ? p.lmax : 1;
        sph_in(z, Lb, il);
        for (int l = 0; l <= p.lmax; l++) {
            const double dil = (l == 0) ? il[1] : il[l - 1] - (l + 1) / z * il[l];
            const double dz = dil * 2.0 * p.alpha * rho;
            for (int n = 0; n < p.nmax; n++) {
                const double gq = p.G[n * p.nq + q];
                if (small) { part[n * L1 + l] += gq * il[l]; part[nnl + n * L1 + l] += gq * dz; }
                else { atomicAdd(&acc[n * L1 + l], gq * il[l]); atomicAdd(&acc[nnl + n * L1 + l], gq * dz); }
            }
        }
    }
    if (small) {
        for (int k = 0; k < 2 * nnl; k++) {
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
// One thread per m; ct = cos(theta), st = sin(theta), (cp, sp) = (cos phi, sin phi).
__device__ void ylm_column(int m, int L, double ct, double st, double cp, double sp, cuDoubleComplex *Y) {
    // Ybar_m^m
    double pmm = sqrt(1.0 / (4.0 * PI));
    for (int k = 1; k <= m; k++) pmm *= -sqrt((2.0 * k + 1.0) / (2.0 * k)) * st;
    // e^{i m phi}
    double cm = 1.0, sm_ = 0.0;
    for (int k = 0; k < m; k++) { const double t = cm * cp - sm_ * sp; sm_ = sm_ * cp + cm * sp; cm = t; }
    double pl2 = 0.0, pl1 = pmm;
    Y[m * (L + 1) + m] = make_cuDoubleComplex(pmm * cm, pmm * sm_);
    for (int l = m + 1; l <= L; l++) {
        double pl;
        if (l == m + 1) pl = sqrt(2.0 * m + 3.0) * ct * pl1;
        else {
            const double a = sqrt((4.0 * l * l - 1.0) / ((double)l * l - (double)m * m));
            const double b = sqrt((((double)l - 1.0) * (l - 1.0) - (double)m * m) / (4.0 * (l - 1.0) * (l - 1.0) - 1.0));
            pl = a * (ct * pl1 - b * pl2);
        }
        Y[l * (L + 1) + m] = make_cuDoubleComplex(pl * cm, pl * sm_);
        pl2 = pl1; pl1 = pl;
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

// one CTA per centre atom
__global__ void __launch_bounds__(128) so3_power_kernel(SO3Power a, SO3Params p) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int L1 = p.lmax + 1, M = 2 * p.lmax + 1, LY = p.lmax + 1;   // Y table up to l = lmax+1
    const int nent = p.nmax * L1 * M;
    const int npair = p.nmax * (p.nmax + 1) / 2, d = npair * L1;
    cuDoubleComplex *sC = reinterpret_cast<cuDoubleComplex *>(raw);                 // [nent] C_tot
    cuDoubleComplex *sdc = sC + nent;                                               // [nent][3] grad c(w)
    cuDoubleComplex *sY = sdc + 3 * nent;                                           // [(LY+1)*(LY+1)]
    double *sAcc = reinterpret_cast<double *>(sY + (LY + 1) * (LY + 1));            // [d*3] current j group
    double *sSelf = sAcc + 3 * d;                                                   // [d*3] sum over j != i
    double *sGeo = sSelf + 3 * d;                                                   // [16] per-neighbour scalars
    double *sAccR = sGeo + 16;                                                      // [d*9] R_j (x) dP of the current j group
    double *sSelfR = sAccR + 9 * d;                                                 // [d*9] the j == i group (own images)
    double *sTot = sSelfR + 9 * d;                                                  // [d*3] sum of dP over all neighbours
    const bool stress = a.rdxdr != nullptr;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int i = blockIdx.x;
    const int w0 = a.nb_ptr[i], w1 = a.nb_ptr[i + 1];
    const int a0 = a.atom_ptr[a.struct_of[i]];
    const int nnl = p.nmax * L1;

    for (int e = tid; e < nent; e += nt) sC[e] = make_cuDoubleComplex(0.0, 0.0);
    for (int o = tid; o < 3 * d; o += nt) { sAcc[o] = 0.0; sSelf[o] = 0.0; }
    if (stress) {
        for (int o = tid; o < 9 * d; o += nt) { sAccR[o] = 0.0; sSelfR[o] = 0.0; }
        for (int o = tid; o < 3 * d; o += nt) sTot[o] = 0.0;
    }
    __syncthreads();

    auto geometry = [&](int w) {
        // thread 0: scalars of neighbour w; threads 0..LY: Y_lm columns
        const double rx = a.nb_rvec[3 * w], ry = a.nb_rvec[3 * w + 1], rz = a.nb_rvec[3 * w + 2];
        const double r = sqrt(rx * rx + ry * ry + rz * rz);
        const double rxy = sqrt(rx * rx + ry * ry);
        const double ct = rz / r, st = rxy / r;
        const double cp = rxy > 0.0 ? rx / rxy : 1.0, sp = rxy > 0.0 ? ry / rxy : 0.0;
        if (tid <= LY) ylm_column(tid, LY, ct, st, cp, sp, sY);
        if (tid == 0) {
            const double gauss = 4.0 * PI * exp(-p.alpha * r * r);
            const double fc = 0.5 * (cos(PI * r / p.rcut) + 1.0);
            const double dfc = -0.5 * PI / p.rcut * sin(PI * r / p.rcut);
            sGeo[0] = r; sGeo[1] = rx / r; sGeo[2] = ry / r; sGeo[3] = rz / r;
            sGeo[4] = gauss; sGeo[5] = -2.0 * p.alpha * r * gauss; sGeo[6] = fc; sGeo[7] = dfc;
            // neighbour weight Z_j; weight_on: a neighbour of another species counts negative (SO3.py:381-385)
            const int zj = a.numbers[a.nb_j[w]];
            sGeo[8] = ((a.derivative & 2) && zj != a.numbers[i]) ? -(double)zj : (double)zj;
        }
    };

    // ---- phase A: C_tot = sum_w Z_j N_l 4pi e^{-a r^2} f_c Y_lm I_nl -----------------------------
    for (int w = w0; w < w1; w++) {
        geometry(w);
        __syncthreads();
        const double pref = sGeo[8] * sGeo[4] * sGeo[6];
        const double *I = a.rad + (size_t)w * 2 * nnl;
        for (int e = tid; e < nent; e += nt) {
            const int n = e / (L1 * M), l = (e / M) % L1, m = e % M - p.lmax;
            if (m < -l || m > l) continue;
            const cuDoubleComplex y = ylm_get(sY, LY, l, m);
            const double f = pref * p.norm_l[l] * I[n * L1 + l];
            sC[e].x += f * y.x; sC[e].y += f * y.y;
        }
        __syncthreads();
    }
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

    // ---- phase B: dP per neighbour, grouped by neighbour atom j -----------------------------------
    int row = a.seq_ptr[i];
    bool self_done = false;
    int self_row = -1;
    const double isq2 = 0.70710678118654752440;
    for (int w = w0; w < w1; w++) {
        __syncthreads();          // previous iteration's readers of sdc / sY / sGeo are done
        geometry(w);
        __syncthreads();
        const double r = sGeo[0], ux = sGeo[1], uy = sGeo[2], uz = sGeo[3];
        const double gauss = sGeo[4], dgauss = sGeo[5], fc = sGeo[6], dfc = sGeo[7], Z = sGeo[8];
        const double *I = a.rad + (size_t)w * 2 * nnl;
        const double *dI = I + nnl;
        for (int e = tid; e < nent; e += nt) {
            const int n = e / (L1 * M), l = (e / M) % L1, m = e % M - p.lmax;
            cuDoubleComplex g0 = make_cuDoubleComplex(0, 0), g1 = g0, g2 = g0;
            if (m >= -l && m <= l) {
                const cuDoubleComplex y = ylm_get(sY, LY, l, m);
                // covariant spherical components of grad Y_lm (SO3.py:686-702)
                cuDoubleComplex c0 = make_cuDoubleComplex(0, 0), cpl = c0, cmi = c0;
                if (l >= 1) {
                    const double ir = 1.0 / r;
                    const double dl = (double)l, dm = (double)m;
                    {
                        const double k1 = -sqrt(((dl + 1) * (dl + 1) - dm * dm) / (2 * dl + 1) / (2 * dl + 3)) * dl * ir;
                        const cuDoubleComplex yu = ylm_get(sY, LY, l + 1, m);
                        c0 = make_cuDoubleComplex(k1 * yu.x, k1 * yu.y);
                        if (abs(m) <= l - 1) {
                            const double k2 = sqrt((dl * dl - dm * dm) / (2 * dl - 1) / (2 * dl + 1)) * (dl + 1) * ir;
                            const cuDoubleComplex yd = ylm_get(sY, LY, l - 1, m);
                            c0.x += k2 * yd.x; c0.y += k2 * yd.y;
                        }
                    }
                    {
                        const double k1 = -sqrt((dl + dm + 1) * (dl + dm + 2) / 2 / (2 * dl + 1) / (2 * dl + 3)) * dl * ir;
                        const cuDoubleComplex yu = ylm_get(sY, LY, l + 1, m + 1);
                        cpl = make_cuDoubleComplex(k1 * yu.x, k1 * yu.y);
                        if (abs(m + 1) <= l - 1) {
                            const double k2 = sqrt((dl - dm - 1) * (dl - dm) / 2 / (2 * dl - 1) / (2 * dl + 1)) * (dl + 1) * ir;
                            const cuDoubleComplex yd = ylm_get(sY, LY, l - 1, m + 1);
                            cpl.x -= k2 * yd.x; cpl.y -= k2 * yd.y;
                        }
                    }
                    {
                        const double k1 = -sqrt((dl - dm + 1) * (dl - dm + 2) / 2 / (2 * dl + 1) / (2 * dl + 3)) * dl * ir;
                        const cuDoubleComplex yu = ylm_get(sY, LY, l + 1, m - 1);
                        cmi = make_cuDoubleComplex(k1 * yu.x, k1 * yu.y);
                        if (abs(m - 1) <= l - 1) {
                            const double k2 = sqrt((dl + dm - 1) * (dl + dm) / 2 / (2 * dl - 1) / (2 * dl + 1)) * (dl + 1) * ir;
                            const cuDoubleComplex yd = ylm_get(sY, LY, l - 1, m - 1);
                            cmi.x -= k2 * yd.x; cmi.y -= k2 * yd.y;
                        }
                    }
                }
                // Cartesian gradient of Y (SO3.py:705-707): x = (c- - c+)/sqrt2, y = i (c- + c+)/sqrt2, z = c0
                const cuDoubleComplex gyx = make_cuDoubleComplex((cmi.x - cpl.x) * isq2, (cmi.y - cpl.y) * isq2);
                const cuDoubleComplex gyy = make_cuDoubleComplex(-(cmi.y + cpl.y) * isq2, (cmi.x + cpl.x) * isq2);
                const cuDoubleComplex gyz = c0;
                const double Inl = I[n * L1 + l], dInl = dI[n * L1 + l];
                const double u[3] = {ux, uy, uz};
                const cuDoubleComplex gy[3] = {gyx, gyy, gyz};
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
        __syncthreads();
        // dP[pair(n,n'), l, k] = Re sum_m [ dc_nlm conj(C_n'lm) + conj(dc_n'lm conj(C_nlm)) ]  (SO3.py:249-251)
        for (int o = tid; o < 3 * d; o += nt) {
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
            if (stress) {      // pstress[(i,j)] -= R_j (x) dP(w), R_j = r_i + r_ij (SO3.py:226, 254, 264)
                sTot[o] += s;
#pragma unroll
                for (int n3 = 0; n3 < 3; n3++)
                    sAccR[(pl * 3 + n3) * 3 + k] += (a.pos[3 * (size_t)i + n3] + a.nb_rvec[3 * (size_t)w + n3]) * s;
            }
        }
        // flush when the next neighbour belongs to another atom
        if (stress) __syncthreads();      // the 9-column accumulators are re-partitioned over the threads below
        const int j = a.nb_j[w];
        const bool last_of_j = (w + 1 == w1) || (a.nb_j[w + 1] != j);
        if (last_of_j) {
            if (!self_done && i < j) { self_row = row++; self_done = true; }
            if (j == i) {
                self_row = row++; self_done = true;
                for (int o = tid; o < 3 * d; o += nt) sAcc[o] = 0.0;     // own images cancel (SO3.py:267-273)
                if (stress) for (int o = tid; o < 9 * d; o += nt) { sSelfR[o] = sAccR[o]; sAccR[o] = 0.0; }
            } else {
                const int rj = row++;
                for (int o = tid; o < 3 * d; o += nt) {
                    const double v = sAcc[o];
                    a.dxdr[(size_t)rj * 3 * d + o] = v;
                    sSelf[o] += v;
                    sAcc[o] = 0.0;
                }
                if (stress) {
                    const double iv = a.inv_vol[a.struct_of[i]];
                    for (int o = tid; o < 9 * d; o += nt) { a.rdxdr[(size_t)rj * 9 * d + o] = sAccR[o] * iv; sAccR[o] = 0.0; }
                }
                if (tid == 0) { a.seq[2 * (size_t)rj] = i - a0; a.seq[2 * (size_t)rj + 1] = j - a0; }
            }
        }
    }
    if (!self_done) self_row = row++;
    for (int o = tid; o < 3 * d; o += nt) a.dxdr[(size_t)self_row * 3 * d + o] = -sSelf[o];
    if (stress) {      // pstress[(i,i)] = -sum_{own images} R_j (x) dP + R_i (x) sum_w dP ; rdxdr = -pstress / vol
        __syncthreads();
        const double iv = a.inv_vol[a.struct_of[i]];
        for (int o = tid; o < 9 * d; o += nt) {
            const int k = o % 3, n3 = (o / 3) % 3, pl = o / 9;
            a.rdxdr[(size_t)self_row * 9 * d + o] = (sSelfR[o] - a.pos[3 * (size_t)i + n3] * sTot[pl * 3 + k]) * iv;
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
    so3_radial_kernel<<<(n_nb + wpb - 1) / wpb, wpb * 32, smem, (cudaStream_t)stream>>>(n_nb, nb_rvec, p, rad);
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
    const size_t smem = (size_t)(4 * nent + (LY + 1) * (LY + 1)) * sizeof(cuDoubleComplex) + (size_t)(6 * d + 16 + 21 * d) * sizeof(double);
    GPRB_REQUIRE(smem <= 200 * 1024, "gprb_so3_power: nmax=%d lmax=%d needs %zu bytes of shared memory", nmax, lmax, smem);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        GPRB_CUDA(cudaFuncSetAttribute(so3_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    so3_power_kernel<<<n_atoms, 128, smem, (cudaStream_t)stream>>>(a, p);
    GPRB_LAUNCHED();
    GPRB_CUDA(cudaGetLastError());
    return GPRB_OK;
}
