"""CPU ORACLE (test infrastructure only) — numpy/scipy restatement of the SO(3) power-spectrum
descriptor and its Cartesian derivative (gpr_calc/SO3.py:186-323, 348-407, 417-453, 608-727).

Parity status: PINNED against the reference's own SO3.calculate run under the stubs of
oracle/ref_harness.py (tests/test_oracle_vs_reference.py, golden vectors in tests/golden/).
The neighbour search restates ase.neighborlist.NeighborList (ase >= 3.23, not vendored, absent
here): radii rcut/2, skin 0, bothways, no self interaction -> all (i, j, S) with
|r_j + S.cell - r_i| < rcut except (i, i, 0).  That third-party piece is UNPINNED by any
reference test; the pair set is order-independent in everything that is compared.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np
from scipy.special import spherical_in, sph_harm_y


def neighbor_pairs(positions, cell, pbc, rcut):
    """Sorted list of (i, j, sx, sy, sz); see module docstring.  O(n^2 images) brute force."""
    positions = np.asarray(positions, float)
    cell = np.asarray(cell, float).reshape(3, 3)
    vol = abs(np.linalg.det(cell))
    nimg = []
    for k in range(3):
        if pbc[k] and vol > 0:
            a, b = cell[(k + 1) % 3], cell[(k + 2) % 3]
            height = vol / np.linalg.norm(np.cross(a, b))
            nimg.append(int(np.ceil(rcut / height)) + 1)
        else:
            nimg.append(0)
    rows = []
    for sx in range(-nimg[0], nimg[0] + 1):
        for sy in range(-nimg[1], nimg[1] + 1):
            for sz in range(-nimg[2], nimg[2] + 1):
                shift = np.array([sx, sy, sz]) @ cell
                dvec = positions[None, :, :] + shift[None, None, :] - positions[:, None, :]
                dist = np.sqrt((dvec ** 2).sum(-1))
                ii, jj = np.nonzero(dist < rcut)
                keep = ~((ii == jj) & (sx == 0) & (sy == 0) & (sz == 0))
                for i, j in zip(ii[keep], jj[keep]):
                    rows.append((int(i), int(j), sx, sy, sz))
    rows.sort()
    return rows


def radial_overlap_W(nmax):
    """W = S^(-1/2) of the polynomial radial basis overlap (SO3.py:417-430)."""
    a = np.arange(1, nmax + 1)
    t = (2 * a + 5) * (2 * a + 6) * (2 * a + 7)
    ab = a[:, None] + a[None, :]
    S = np.sqrt(t[:, None] * t[None, :]) / ((5 + ab) * (6 + ab) * (7 + ab))
    vals, vecs = np.linalg.eig(np.linalg.inv(S))
    return (vecs @ np.diag(np.sqrt(vals)) @ np.linalg.inv(vecs)).real


def radial_table(nmax, lmax, rcut, alpha):
    """Quadrature nodes rho_q and the weights G[n, q] that multiply i_l(2 alpha r rho_q)
    (SO3.py:446-453, 619-633, 646-647)."""
    NQ = (nmax + lmax + 1) * 10
    q = np.arange(1, NQ + 1)
    t = np.cos((2 * q - 1) * np.pi / 2 / NQ)
    w = np.pi / NQ * rcut / 2
    rho = rcut / 2 * (t + 1)
    W = radial_overlap_W(nmax)
    a = np.arange(1, nmax + 1)
    phi = (rcut - rho[None, :]) ** (a[:, None] + 2) / np.sqrt(
        2 * rcut ** (2 * a[:, None] + 7) / (2 * a[:, None] + 5) / (2 * a[:, None] + 6) / (2 * a[:, None] + 7))
    g = W @ phi
    G = g * w * rho ** 2 * np.exp(-alpha * rho ** 2) * np.sqrt(1 - t ** 2)
    return rho, G


def expansion_coefficients(rvec, nmax, lmax, rcut, alpha):
    """c_nlm and grad c_nlm of every neighbour vector (SO3.py:608-727, cosine cutoff :409-415).

    Returns C [P, nmax, lmax+1, 2lmax+1] and dC [P, nmax, lmax+1, 2lmax+1, 3] (complex)."""
    rvec = np.asarray(rvec, float)
    P = len(rvec)
    r = np.linalg.norm(rvec, axis=1)
    u = rvec / r[:, None]
    rho, G = radial_table(nmax, lmax, rcut, alpha)
    z = 2 * alpha * np.outer(r, rho)
    ls = np.arange(lmax + 1)
    bes = np.stack([spherical_in(l, z) for l in ls], axis=-1)                    # [P, Q, L]
    dbes = np.stack([spherical_in(l, z, derivative=True) for l in ls], axis=-1)
    I = np.einsum('nq,pql->pnl', G, bes)
    dI_dr = np.einsum('nq,pql->pnl', G * (2 * alpha * rho)[None, :], dbes)
    gauss = 4 * np.pi * np.exp(-alpha * r ** 2)
    dgauss_dr = -2 * alpha * r * gauss
    fc = 0.5 * (np.cos(np.pi * r / rcut) + 1.0)
    dfc_dr = -0.5 * np.pi / rcut * np.sin(np.pi * r / rcut)
    theta = np.arccos(rvec[:, 2] / r)
    phi = np.arctan2(rvec[:, 1], rvec[:, 0])
    M = 2 * lmax + 1
    Y = np.zeros((P, lmax + 2, 2 * (lmax + 1) + 1), dtype=complex)
    mid = lmax + 1
    for l in range(lmax + 2):
        for m in range(-l, l + 1):
            Y[:, l, mid + m] = sph_harm_y(l, m, theta, phi)
    gY = np.zeros((P, lmax + 1, M, 3), dtype=complex)
    for l in range(1, lmax + 1):
        for m in range(-l, l + 1):
            # covariant spherical components of grad Y_lm (SO3.py:686-707)
            c0 = -np.sqrt(((l + 1) ** 2 - m ** 2) / (2 * l + 1) / (2 * l + 3)) * l * Y[:, l + 1, mid + m] / r
            if abs(m) <= l - 1:
                c0 = c0 + np.sqrt((l ** 2 - m ** 2) / (2 * l - 1) / (2 * l + 1)) * (l + 1) * Y[:, l - 1, mid + m] / r
            cp = -np.sqrt((l + m + 1) * (l + m + 2) / 2 / (2 * l + 1) / (2 * l + 3)) * l * Y[:, l + 1, mid + m + 1] / r
            if abs(m + 1) <= l - 1:
                cp = cp - np.sqrt((l - m - 1) * (l - m) / 2 / (2 * l - 1) / (2 * l + 1)) * (l + 1) * Y[:, l - 1, mid + m + 1] / r
            cm = -np.sqrt((l - m + 1) * (l - m + 2) / 2 / (2 * l + 1) / (2 * l + 3)) * l * Y[:, l + 1, mid + m - 1] / r
            if abs(m - 1) <= l - 1:
                cm = cm - np.sqrt((l + m - 1) * (l + m) / 2 / (2 * l - 1) / (2 * l + 1)) * (l + 1) * Y[:, l - 1, mid + m - 1] / r
            gY[:, l, lmax + m, 0] = (cm - cp) / np.sqrt(2)
            gY[:, l, lmax + m, 1] = 1j * (cm + cp) / np.sqrt(2)
            gY[:, l, lmax + m, 2] = c0
    Yl = Y[:, :lmax + 1, 1:1 + M]
    YI = np.einsum('plm,pnl->pnlm', Yl, I)
    C0 = gauss[:, None, None, None] * YI
    dC = (dgauss_dr[:, None] * u)[:, None, None, None, :] * YI[..., None]
    dC = dC + gauss[:, None, None, None, None] * (
        np.einsum('plmx,pnl->pnlmx', gY, I) + np.einsum('plm,pnl,px->pnlmx', Yl, dI_dr, u))
    dC = dC * fc[:, None, None, None, None] + (dfc_dr[:, None] * u)[:, None, None, None, :] * C0[..., None]
    return C0 * fc[:, None, None, None], dC


def so3_calculate(positions, cell, pbc, numbers, nmax=3, lmax=4, rcut=5.0, alpha=2.0, stress=False, weight_on=False,
                  atom_ids=None):
    """Power spectrum x [n, d], its derivative dxdr [n_seq, d, 3] and seq [n_seq, 2] (int64),
    with the reference's conventions (SO3.py:186-323): weights Z_j, norm_l, tril(n >= n') x l layout,
    dxdr[(i, j)] = dx_i/dr_j summed over images, dxdr[(i, i)] = - sum_{j != i}.
    stress=True also returns rdxdr [n_seq, d, 3, 3] = -pstress / volume (SO3.py:253-273, 304-306):
    pstress[(i, j)] = -sum_w R_j(w) (x) dP(w), pstress[(i, i)] += R_i (x) sum_w dP(w).
    weight_on: a neighbour of another species than the centre weighs -Z_j (SO3.py:381-385).
    atom_ids: centres to evaluate, in the given order (SO3.py:354-401); other rows of x stay zero."""
    positions = np.asarray(positions, float)
    numbers = np.asarray(numbers)
    n = len(positions)
    pairs = neighbor_pairs(positions, cell, pbc, rcut)
    cell = np.asarray(cell, float).reshape(3, 3)
    d = nmax * (nmax + 1) // 2 * (lmax + 1)
    tril = np.tril_indices(nmax)
    # seq: for each centre the sorted set {neighbours} U {centre}  (SO3.py:389-401)
    nb_sets = [set([i]) for i in range(n)]
    for (i, j, *_s) in pairs:
        nb_sets[i].add(j)
    centres = list(range(n)) if atom_ids is None else [int(i) for i in atom_ids]
    seq = np.array([[i, j] for i in centres for j in sorted(nb_sets[i])], dtype=np.int64).reshape(-1, 2)
    row_of = {(int(a), int(b)): k for k, (a, b) in enumerate(seq)}
    x = np.zeros((n, d))
    dxdr = np.zeros((len(seq), d, 3))
    pstress = np.zeros((len(seq), d, 3, 3))
    vol = abs(np.linalg.det(cell))
    if not pairs:
        return (x, dxdr, seq, -pstress / vol) if stress else (x, dxdr, seq)
    pi = np.array([p[0] for p in pairs])
    pj = np.array([p[1] for p in pairs])
    S = np.array([p[2:] for p in pairs], dtype=float)
    rvec = positions[pj] + S @ cell - positions[pi]
    C, dC = expansion_coefficients(rvec, nmax, lmax, rcut, alpha)
    ls = np.arange(lmax + 1)
    norm = np.sqrt(2 * np.sqrt(2) * np.pi / np.sqrt(2 * ls + 1))
    wgt = numbers[pj].astype(float)
    if weight_on:
        wgt = np.where(numbers[pj] != numbers[pi], -wgt, wgt)
    C = C * wgt[:, None, None, None] * norm[None, None, :, None]
    dC = dC * wgt[:, None, None, None, None] * norm[None, None, :, None, None]
    for i in centres:
        sel = np.nonzero(pi == i)[0]
        if len(sel) == 0:
            continue
        ctot = C[sel].sum(axis=0)
        Pm = np.einsum('alm,blm->abl', ctot, np.conj(ctot)).real
        x[i] = Pm[tril].reshape(-1)
        dP = np.einsum('walmx,blm->wablx', dC[sel], np.conj(ctot))
        dP = (dP + np.conj(np.transpose(dP, (0, 2, 1, 3, 4)))).real
        dPt = dP[:, tril[0], tril[1]].reshape(len(sel), d, 3)
        for k, w in enumerate(sel):
            dxdr[row_of[(i, int(pj[w]))]] += dPt[k]
            if stress:
                pstress[row_of[(i, int(pj[w]))]] -= np.einsum('n,dm->dnm', positions[i] + rvec[w], dPt[k])
        ii = row_of[(i, i)]
        own = [row_of[(i, j)] for j in sorted(nb_sets[i])]
        dxdr[ii] -= dxdr[own].sum(axis=0)
        if stress:
            pstress[ii] += np.einsum('n,dm->dnm', positions[i], dPt.sum(axis=0))
    if stress:
        return x, dxdr, seq, -pstress / vol
    return x, dxdr, seq


class SO3Oracle:
    """Descriptor-protocol wrapper (calculate(atoms) -> dict) around so3_calculate."""

    def __init__(self, nmax=3, lmax=4, rcut=5.0, alpha=2.0, stress=False):
        self.nmax, self.lmax, self.rcut, self.alpha, self.stress = nmax, lmax, rcut, alpha, stress

    def calculate(self, atoms, atom_ids=None, use_mpi=False):
        r = so3_calculate(atoms.positions, np.asarray(atoms.cell), atoms.pbc, atoms.numbers,
                          self.nmax, self.lmax, self.rcut, self.alpha, stress=self.stress)
        return {'x': r[0], 'dxdr': r[1], 'rdxdr': r[3] if self.stress else None, 'elements': list(atoms.symbols), 'seq': r[2]}
