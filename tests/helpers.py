"""Seeded synthetic inputs shared by the parity tests."""
import numpy as np


def make_force(rng, n_groups, d=30, lo=3, hi=40, species=(13, 79), scale=1.0, zero_rows=0):
    """List of (x, dxdr, ele) force data with ragged group sizes."""
    out = []
    base = np.abs(rng.normal(size=d)) + 0.5
    for _ in range(n_groups):
        n = int(rng.integers(lo, hi + 1))
        x = (base[None, :] + 0.3 * rng.normal(size=(n, d))) * scale
        dx = rng.normal(size=(n, d, 3)) * scale
        ele = rng.choice(species, size=n)
        for _z in range(zero_rows):
            x[int(rng.integers(0, n))] = 0.0
        out.append((x, dx, ele))
    return out


def make_energy(rng, n_groups, d=30, lo=4, hi=60, species=(13, 79), scale=1.0):
    out = []
    base = np.abs(rng.normal(size=d)) + 0.5
    for _ in range(n_groups):
        n = int(rng.integers(lo, hi + 1))
        x = (base[None, :] + 0.3 * rng.normal(size=(n, d))) * scale
        ele = rng.choice(species, size=n)
        out.append((x, ele))
    return out


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / den) if den > 0 else float(np.abs(a - b).max())


def entry_err(a, b, d_rows, d_cols):
    """Per-entry error on the Cauchy-Schwarz scale of each entry: max_ij |a_ij - b_ij| / sqrt(k(i,i) k(j,j)), with d_rows / d_cols
    the prior variances k(i,i) of the row side and k(j,j) of the column side.  |K_ij| <= sqrt(K_ii K_jj) for a covariance, so
    this is the relative error of every entry measured against the natural magnitude of ITS sum of pair terms (the absolute
    floor tol * |A| |B| of SURVEY.md 7.3): small entries of a block with large neighbours are checked too, entries that are
    sums of cancelling terms are not held to their own (arbitrarily small) value."""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.sqrt(np.abs(np.asarray(d_rows))[:, None] * np.abs(np.asarray(d_cols))[None, :])
    scale = np.where(scale > 0, scale, 1.0)
    return float((np.abs(a - b) / scale).max()) if a.size else 0.0
