"""GPU parity: the device SO3 descriptor against the reference's golden vectors and the oracle.
seq (neighbour indexing) must be bit-exact; x / dxdr within 1e-10 relative."""
import os

import numpy as np
import pytest

from helpers import rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_so3_vs_golden():
    from gpr_calculator_b200.SO3 import SO3
    from gpr_calculator_b200.utilities import SimpleAtoms
    g = np.load(os.path.join(GOLD, "so3.npz"))
    for k in range(3):
        prm = g["s%d_prm" % k]
        at = SimpleAtoms(g["s%d_numbers" % k], g["s%d_pos" % k], g["s%d_cell" % k], g["s%d_pbc" % k])
        r = SO3(nmax=int(prm[0]), lmax=int(prm[1]), rcut=float(prm[2]), alpha=float(prm[3])).calculate(at)
        assert r["seq"].dtype == np.int64 and np.array_equal(r["seq"], g["s%d_seq" % k])
        assert rel_err(r["x"], g["s%d_x" % k]) <= 1e-10
        assert rel_err(r["dxdr"], g["s%d_dxdr" % k]) <= 1e-10
        assert r["elements"] == at.symbols and r["rdxdr"] is None


def _cu(nrep, seed, noise=0.05, a=3.61):
    from gpr_calculator_b200.utilities import SimpleAtoms
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) * a
    pos = np.concatenate([base + np.array([i, j, k]) * a for i in range(nrep) for j in range(nrep) for k in range(nrep)])
    rng = np.random.default_rng(seed)
    return SimpleAtoms([29] * len(pos), pos + rng.normal(scale=noise, size=pos.shape), np.eye(3) * a * nrep)


def test_so3_vs_oracle_and_batch():
    """Cu fcc 2x2x2 (the S5 unit: images collapse onto few unique neighbours), a molecule in a
    non-periodic box, parameter sweeps; batch == one by one; translation invariance."""
    from oracle import so3 as oso3
    from gpr_calculator_b200.SO3 import SO3
    from gpr_calculator_b200.utilities import SimpleAtoms
    strucs = [_cu(2, 2000), _cu(1, 5),
              SimpleAtoms([1, 1, 8], [[0, 0, 0], [0, 0, 0.96], [8, 8, 8]], np.eye(3) * 20, pbc=(False, False, False)),
              SimpleAtoms([13] * 3 + [79], np.random.default_rng(1).uniform(0, 5, (4, 3)), np.diag([5.7, 5.7, 13.0]), pbc=(True, True, False))]
    for prm in ((3, 4, 5.0, 2.0), (4, 3, 4.0, 1.5), (2, 6, 3.5, 2.0), (1, 0, 3.0, 1.0)):
        des = SO3(nmax=prm[0], lmax=prm[1], rcut=prm[2], alpha=prm[3])
        batch = des.calculate_batch(strucs)
        for at, rb in zip(strucs, batch):
            x, dxdr, seq = oso3.so3_calculate(at.positions, at.cell, at.pbc, at.numbers, *prm)
            r = des.calculate(at)
            assert np.array_equal(r["seq"], seq) and np.array_equal(rb["seq"], seq)
            assert rel_err(r["x"], x) <= 1e-10 and rel_err(r["dxdr"], dxdr) <= 1e-10
            assert np.array_equal(rb["x"], r["x"]) and np.array_equal(rb["dxdr"], r["dxdr"])
    # sum_j dx_i/dr_j = 0 (translation invariance): rows of each centre cancel
    des = SO3(nmax=3, lmax=4, rcut=5.0)
    r = des.calculate(strucs[0])
    tot = np.zeros((len(strucs[0]), 30, 3))
    np.add.at(tot, r["seq"][:, 0], r["dxdr"])
    assert np.abs(tot).max() <= 1e-9 * np.abs(r["dxdr"]).max()


def test_so3_finite_difference():
    """dxdr is the true derivative of x: central differences on one coordinate."""
    from gpr_calculator_b200.SO3 import SO3
    at = _cu(1, 9, noise=0.1)
    des = SO3(nmax=3, lmax=4, rcut=4.0)
    r = des.calculate(at)
    j, c, h = 2, 1, 1e-5
    ap, am = at.copy(), at.copy()
    ap.positions[j, c] += h
    am.positions[j, c] -= h
    fd = (des.calculate(ap)["x"] - des.calculate(am)["x"]) / (2 * h)
    for (i, jj), row in zip(r["seq"], r["dxdr"]):
        if jj == j:
            assert np.abs(row[:, c] - fd[i]).max() <= 1e-6 * max(1.0, np.abs(fd[i]).max())


def test_so3_options_vs_golden():
    """SO3(weight_on=True), derivative=False and calculate(atom_ids=...) against the reference's own output."""
    from gpr_calculator_b200.SO3 import SO3
    from gpr_calculator_b200.utilities import SimpleAtoms
    g = np.load(os.path.join(GOLD, "so3_options.npz"))
    for k in range(2):
        prm = g["s%d_prm" % k]
        kw = dict(nmax=int(prm[0]), lmax=int(prm[1]), rcut=float(prm[2]), alpha=float(prm[3]))
        at = SimpleAtoms(g["s%d_numbers" % k], g["s%d_pos" % k], g["s%d_cell" % k], g["s%d_pbc" % k])
        r = SO3(weight_on=True, **kw).calculate(at)
        assert np.array_equal(r["seq"], g["s%d_w_seq" % k])
        assert rel_err(r["x"], g["s%d_w_x" % k]) <= 1e-10 and rel_err(r["dxdr"], g["s%d_w_dxdr" % k]) <= 1e-10
        r = SO3(weight_on=True, derivative=False, **kw).calculate(at)
        assert r["dxdr"] is None and rel_err(r["x"], g["s%d_nod_x" % k]) <= 1e-10
        r = SO3(**kw).calculate(at, atom_ids=list(g["s%d_ids" % k]))
        assert r["seq"].dtype == np.int64 and np.array_equal(r["seq"], g["s%d_sub_seq" % k])
        assert rel_err(r["x"], g["s%d_sub_x" % k]) <= 1e-10 and rel_err(r["dxdr"], g["s%d_sub_dxdr" % k]) <= 1e-10


def test_get_data_from_database_on_device():
    """utilities.get_data with the device descriptor (one batched pass over the database rows) against the
    reference's own get_data output (tests/golden/getdata.*)."""
    from test_host_logic import _check_get_data
    from gpr_calculator_b200.SO3 import SO3
    _check_get_data(SO3(nmax=3, lmax=4, rcut=5.0), 1e-10)
