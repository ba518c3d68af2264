"""Seeded synthetic training / test sets of the benchmark configurations.

S4 = 200 x Cu fcc 3x3x3 (108 atoms), S5 = 340 x Cu fcc 2x2x2 (32 atoms): a = 3.61 A, positions +
N(0, 0.05 A) per coordinate with numpy.random.default_rng(seed0 + k) for structure k,
SO3(nmax=3, lmax=4, rcut=5.0).  The training data are injected in the packed layout the reference's
``utilities.get_data`` / ``list_to_tuple`` produce (utilities.py:142-163, 340-390): energy item
(x, E/n_atoms, Z), force item (x[seq[ids,0]], dxdr[ids], F[i], Z[seq[ids,0]]) with
ids = argwhere(seq[:,1] == i) — not through ``add_structure``, whose de-duplication collapses
near-identical Cu environments to about one force centre per structure.

Labels come from an Einstein-crystal potential E = k/2 sum |u_i|^2, F_i = -k u_i (u = displacement
from the lattice site); they do not influence the covariance build.
"""
import numpy as np
import torch

from .utilities import SimpleAtoms

A_CU = 3.61
K_SPRING = 2.0   # eV / A^2


def cu_fcc(nrep, seed, noise=0.05, a=A_CU):
    """One noisy Cu fcc nrep^3 supercell; returns (SimpleAtoms, energy, forces)."""
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) * a
    site = np.concatenate([base + np.array([i, j, k]) * a
                           for i in range(nrep) for j in range(nrep) for k in range(nrep)])
    u = np.random.default_rng(seed).normal(scale=noise, size=site.shape)
    atoms = SimpleAtoms([29] * len(site), site + u, np.eye(3) * a * nrep)
    return atoms, 0.5 * K_SPRING * float((u ** 2).sum()), -K_SPRING * u


def structures(n, nrep, seed0):
    return [cu_fcc(nrep, seed0 + k) for k in range(n)]


def packed_from_batch(des, atoms_list, centres_per_structure=None, chunk=64):
    """Descriptors of `atoms_list` on the device -> packed (energy tuple, force tuple) of CUDA tensors.

    energy: (X [A, d], ELE [A] int32, indices [S])
    force : (X [R, d], dXdR [R, d, 3], ELE [R] int32, indices [NF]) with one force centre per atom
            (or the first `centres_per_structure` atoms of every structure), rows of a centre in
            increasing `seq` order exactly like argwhere(seq[:, 1] == i).
    """
    from .batch import rows_from_batch
    Es, Fs = [], []
    for s0 in range(0, len(atoms_list), chunk):
        part = atoms_list[s0:s0 + chunk]
        r = des.calculate_batch(part, to_host=False)
        centres = None
        if centres_per_structure is not None:
            centres = [range(min(centres_per_structure, len(a))) for a in part]
        E, F = rows_from_batch(r, centres)
        Es.append(E)
        Fs.append(F)
    E = (torch.cat([e[0] for e in Es]), torch.cat([e[1] for e in Es]), sum((e[2] for e in Es), []))
    F = (torch.cat([f[0] for f in Fs]), torch.cat([f[1] for f in Fs]), torch.cat([f[2] for f in Fs]), sum((f[3] for f in Fs), []))
    return E, F


def targets(labelled, centres_per_structure=None):
    """y column of the GP: per-atom energies, then (Fx, Fy, Fz) per force centre (gaussianprocess.py:472-488)."""
    e = [E / len(a) for a, E, _ in labelled]
    f = [F[:centres_per_structure] if centres_per_structure is not None else F for _, _, F in labelled]
    return np.concatenate((np.asarray(e), np.concatenate(f).reshape(-1))).reshape(-1, 1)


def to_host(packed, pin=False):
    """CUDA-tensor packed tuple -> numpy packed tuple (optionally views of pinned host tensors)."""
    out, keep = [], []
    for t in packed:
        if isinstance(t, torch.Tensor):
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=pin)
            h.copy_(t)
            keep.append(h)
            out.append(h.numpy())
        else:
            out.append(t)
    return tuple(out), keep


def pair_counts(ele, indices, symmetric):
    """Same-species row pairs of a packed side against itself: full block, or the I <= J group blocks
    a symmetric build evaluates.  (ele: numpy int array per row.)"""
    ele = np.asarray(ele)
    indices = np.asarray(indices, dtype=np.int64)
    grp = np.repeat(np.arange(len(indices)), indices)
    full = upper = 0
    for z in np.unique(ele):
        c = np.bincount(grp[ele == z], minlength=len(indices)).astype(np.float64)
        tot = c.sum()
        full += tot * tot
        # sum_{I<=J} c_I c_J = (tot^2 + sum c_I^2) / 2
        upper += 0.5 * (tot * tot + (c ** 2).sum())
    return int(upper if symmetric else full)
