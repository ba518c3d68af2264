"""Stress path: rdxdr descriptor, *_stress covariance blocks, k_total_with_stress and
predict_structure(stress=True) against golden vectors produced by the reference
(tests/golden/gen_golden_stress.py).  CPU part: the oracle against the golden vectors."""
import os

import numpy as np
import pytest

from helpers import rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-10


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLD, "stress.npz"))


def _data(g):
    F1 = (g["F1_x"], g["F1_dxdr9"], g["F1_ele"], list(g["F1_ind"]))
    F2 = (g["F2_x"], g["F2_dxdr"], g["F2_ele"], list(g["F2_ind"]))
    E2 = (g["E2_x"], g["E2_ele"], list(g["E2_ind"]))
    E1 = (g["E1_x"], g["E1_ele"], list(g["E1_ind"]))
    return E1, E2, F1, F2


# ---- CPU: the oracle is pinned by the reference's own output -------------------------------------------
def test_oracle_rdxdr_vs_golden(g):
    from oracle import so3 as oso3
    for k in range(2):
        x, dxdr, seq, rd = oso3.so3_calculate(g["s%d_pos" % k], g["s%d_cell" % k], g["s%d_pbc" % k], g["s%d_numbers" % k],
                                              3, 4, 5.0, 2.0, stress=True)
        assert np.array_equal(seq, g["s%d_seq" % k])
        assert rel_err(dxdr, g["s%d_dxdr" % k]) <= TOL and rel_err(rd, g["s%d_rdxdr" % k]) <= TOL


def test_oracle_stress_blocks_vs_golden(g, oracle_libs):
    E1, E2, F1, F2 = _data(g)
    sig, l, zeta = g["params"]
    for backend in ("port", "ref") if oracle_libs.have_ref() else ("port",):
        O, OD = oracle_libs.RBFOracle(backend), oracle_libs.DotOracle(backend)
        C, Cs = O.kff_C(F1, F2, sig, l, zeta, stress=True, tol=1e-12)
        assert rel_err(C, g["rbf_kff_C"]) <= TOL and rel_err(Cs, g["rbf_kff_Cs"]) <= TOL
        C, Cs = O.kef_C(E2, F1, sig, l, zeta, stress=True)
        assert rel_err(C, g["rbf_kef_C"]) <= TOL and rel_err(Cs, g["rbf_kef_Cs"]) <= TOL
        C, Cs = OD.kff_C(F1, F2, 2.0, 1.5, zeta, stress=True)
        assert rel_err(C, g["dot_kff_C"]) <= TOL and rel_err(Cs, g["dot_kff_Cs"]) <= TOL
        C, Cs = OD.kef_C(E2, F1, 2.0, 1.5, zeta, stress=True)
        assert rel_err(C, g["dot_kef_C"]) <= TOL and rel_err(Cs, g["dot_kef_Cs"]) <= TOL


# ---- GPU ---------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_so3_rdxdr_vs_golden(g):
    from gpr_calculator_b200.SO3 import SO3
    from gpr_calculator_b200.utilities import SimpleAtoms
    des = SO3(nmax=3, lmax=4, rcut=5.0, alpha=2.0, stress=True)
    for k in range(2):
        at = SimpleAtoms(g["s%d_numbers" % k], g["s%d_pos" % k], g["s%d_cell" % k], g["s%d_pbc" % k])
        r = des.calculate(at)
        assert np.array_equal(r["seq"], g["s%d_seq" % k])
        assert r["rdxdr"].shape == g["s%d_rdxdr" % k].shape
        assert rel_err(r["dxdr"], g["s%d_dxdr" % k]) <= TOL and rel_err(r["rdxdr"], g["s%d_rdxdr" % k]) <= TOL
    # batch == one by one, and stress=False still returns rdxdr None
    ats = [SimpleAtoms(g["s%d_numbers" % k], g["s%d_pos" % k], g["s%d_cell" % k], g["s%d_pbc" % k]) for k in range(2)]
    for rb, k in zip(des.calculate_batch(ats), range(2)):
        assert rel_err(rb["rdxdr"], g["s%d_rdxdr" % k]) <= TOL
    assert SO3(nmax=3, lmax=4, rcut=5.0).calculate(ats[1])["rdxdr"] is None


@pytest.mark.gpu
def test_stress_blocks_vs_golden(g):
    from gpr_calculator_b200.kernels import rbf_kernel as rk, dot_kernel as dk, RBF_mb
    from gpr_calculator_b200.utilities import tuple_to_list
    E1, E2, F1, F2 = _data(g)
    sig, l, zeta = g["params"]
    C, Cs = rk.kff_C(F1, F2, sig, l, zeta, stress=True, tol=1e-12)
    assert rel_err(C, g["rbf_kff_C"]) <= TOL and rel_err(Cs, g["rbf_kff_Cs"]) <= TOL
    C, Cs = rk.kef_C(E2, F1, sig, l, zeta, stress=True)
    assert rel_err(C, g["rbf_kef_C"]) <= TOL and rel_err(Cs, g["rbf_kef_Cs"]) <= TOL
    C, Cs = rk.kef_C(E2, F1, sig, l, zeta, stress=True, transpose=True)
    assert rel_err(C, g["rbf_kfe_C"]) <= TOL and rel_err(Cs, g["rbf_kse_C"]) <= TOL
    C, Cs = dk.kff_C(F1, F2, 2.0, 1.5, zeta, stress=True)
    assert rel_err(C, g["dot_kff_C"]) <= TOL and rel_err(Cs, g["dot_kff_Cs"]) <= TOL
    C, Cs = dk.kef_C(E2, F1, 2.0, 1.5, zeta, stress=True)
    assert rel_err(C, g["dot_kef_C"]) <= TOL and rel_err(Cs, g["dot_kef_Cs"]) <= TOL
    rbf = RBF_mb(para=[sig, l], zeta=2)
    for force in (F1, tuple_to_list(F1)):                  # packed tuple and the list form predict_structure builds
        C, C1 = rbf.k_total_with_stress({"energy": E1, "force": force}, {"energy": E2, "force": F2}, 1e-12)
        assert rel_err(C, g["RBF_stress_C"]) <= TOL and rel_err(C1, g["RBF_stress_C1"]) <= TOL


@pytest.mark.gpu
def test_predict_structure_with_stress_vs_golden(g):
    import io
    import contextlib
    from gpr_calculator_b200.gaussianprocess import GP
    from gpr_calculator_b200.kernels import RBF_mb
    from gpr_calculator_b200.SO3 import SO3
    from gpr_calculator_b200.utilities import SimpleAtoms, convert_train_data
    mk = lambda pos: SimpleAtoms(g["numbers"], pos, g["cell"], g["pbc"])   # noqa: E731
    des = SO3(nmax=3, lmax=4, rcut=5.0, stress=True)
    labelled = [(mk(g["t%d_pos" % k]), float(g["t%d_E" % k]), g["t%d_F" % k]) for k in range(3)]
    gp = GP(kernel=RBF_mb(para=[2.0, 0.8], zeta=2.0), descriptor=des, noise_e=0.002, noise_f=0.1, log_file=None)
    with contextlib.redirect_stdout(io.StringIO()):
        gp.fit(TrainData=convert_train_data(labelled, des), opt=False, show=False)
    test = mk(g["test_pos"])
    E, F, S = gp.predict_structure(test, stress=True, return_std=False, f_tol=1e-12)
    assert abs(E - g["pred_E"]) <= 1e-8 and np.abs(F - g["pred_F"]).max() <= 1e-8
    assert S.shape == (13, 6) and np.abs(S - g["pred_S"]).max() <= 1e-8 * max(1.0, np.abs(g["pred_S"]).max())
    E2, F2, S2, E_std, F_std = gp.predict_structure(test, return_std=True, f_tol=1e-12)     # default stress=True
    assert abs(E2 - E) <= 1e-12 and np.array_equal(S2, S) and np.all(np.isfinite(F_std)) and np.isfinite(E_std)
    E3, F3, _ = gp.predict_structure(test, stress=False, f_tol=1e-12)
    assert abs(E3 - E) <= 1e-9 and np.abs(F3 - F).max() <= 1e-9
    plain = GP(kernel=RBF_mb(para=[2.0, 0.8], zeta=2.0), descriptor=SO3(nmax=3, lmax=4, rcut=5.0), log_file=None)
    with pytest.raises(ValueError):
        plain.predict_structure(test)            # stress=True without rdxdr: loud, not silent
