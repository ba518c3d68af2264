"""Golden vectors of the stress path, produced by the UNMODIFIED reference under the
stubs of oracle/ref_harness.py (build container only):  python tests/golden/gen_golden_stress.py

stress.npz:
    SO3(stress=True).calculate -> rdxdr for two structures
    rbf / dot kef_C(stress=True), kff_C(stress=True) on 9-column force data (3 force + 6 Voigt columns)
    RBF_mb.k_total_with_stress
    GP.predict_structure(stress=True) energy, forces and per-atom stress
"""
import io
import os
import sys
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from oracle import ref_harness as rh   # noqa: E402
from helpers import make_force, make_energy   # noqa: E402
from gen_golden import slab, toy_labels   # noqa: E402


def nine_columns(rng, data):
    """append 6 random Voigt columns to the dxdr of every (x, dxdr, ele) item"""
    return [(x, np.concatenate((dx, rng.normal(size=(len(x), x.shape[1], 6))), axis=2), ele) for x, dx, ele in data]


def main():
    os.chdir("/tmp")
    m = rh.modules()
    lt = m.utilities.list_to_tuple
    out = {}
    # ---- descriptor --------------------------------------------------------------------------------
    strucs = [slab(11, n_fixed=0), rh.Atoms([29] * 4, np.array([[0, 0, 0], [1.8, 1.8, 0], [1.8, 0, 1.8], [0, 1.8, 1.8]]) + 0.03 *
                                            np.random.default_rng(3).normal(size=(4, 3)), np.eye(3) * 3.61)]
    for k, at in enumerate(strucs):
        r = m.SO3(nmax=3, lmax=4, rcut=5.0, alpha=2.0, stress=True).calculate(at)
        out["s%d_numbers" % k], out["s%d_pos" % k], out["s%d_cell" % k], out["s%d_pbc" % k] = at.numbers, at.positions, np.asarray(at.cell), at.pbc
        out["s%d_rdxdr" % k], out["s%d_dxdr" % k], out["s%d_seq" % k] = r["rdxdr"], r["dxdr"], r["seq"]
    # ---- covariance blocks with stress columns --------------------------------------------------------
    rng = np.random.default_rng(20261019)
    F1l = nine_columns(rng, make_force(rng, 5, lo=2, hi=11))
    F2 = lt(make_force(rng, 4, lo=1, hi=13))
    E2 = lt(make_energy(rng, 3, lo=3, hi=12), mode="energy")
    F1 = lt(F1l, stress=True)
    out["F1_x"], out["F1_dxdr9"], out["F1_ele"], out["F1_ind"] = F1[0], F1[1], F1[2], np.array(F1[3])
    out["F2_x"], out["F2_dxdr"], out["F2_ele"], out["F2_ind"] = F2[0], F2[1], F2[2], np.array(F2[3])
    out["E2_x"], out["E2_ele"], out["E2_ind"] = E2[0], E2[1], np.array(E2[2])
    sig, l, zeta = 1.3, 0.7, 2.0
    out["params"] = np.array([sig, l, zeta])
    C, Cs = m.rbf_kernel.kff_C(F1, F2, sig, l, zeta, stress=True, tol=1e-12)
    out["rbf_kff_C"], out["rbf_kff_Cs"] = C, Cs
    C, Cs = m.rbf_kernel.kef_C(E2, F1, sig, l, zeta, stress=True)
    out["rbf_kef_C"], out["rbf_kef_Cs"] = C, Cs
    C, Cs = m.rbf_kernel.kef_C(E2, F1, sig, l, zeta, stress=True, transpose=True)
    out["rbf_kfe_C"], out["rbf_kse_C"] = C, Cs
    C, Cs = m.dot_kernel.kff_C(F1, F2, 2.0, 1.5, zeta, stress=True)
    out["dot_kff_C"], out["dot_kff_Cs"] = C, Cs
    C, Cs = m.dot_kernel.kef_C(E2, F1, 2.0, 1.5, zeta, stress=True)
    out["dot_kef_C"], out["dot_kef_Cs"] = C, Cs
    rbf = m.RBF_mb(para=[sig, l], zeta=2)
    E1 = lt(make_energy(rng, 1, lo=9, hi=9), mode="energy")
    out["E1_x"], out["E1_ele"], out["E1_ind"] = E1[0], E1[1], np.array(E1[2])
    C, C1 = rbf.k_total_with_stress({"energy": E1, "force": F1}, {"energy": E2, "force": F2}, 1e-12)
    out["RBF_stress_C"], out["RBF_stress_C1"] = C, C1
    # ---- GP.predict_structure(stress=True) --------------------------------------------------------------
    train = [slab(100 + k, n_fixed=0) for k in range(3)]
    test = slab(200, n_fixed=0)
    des = m.SO3(nmax=3, lmax=4, rcut=5.0, stress=True)
    labelled = []
    for k, at in enumerate(train):
        E, F = toy_labels(at, 300 + k)
        labelled.append((at, E, F))
        out["t%d_pos" % k], out["t%d_E" % k], out["t%d_F" % k] = at.positions, E, F
    out["numbers"], out["cell"], out["pbc"] = train[0].numbers, np.asarray(train[0].cell), train[0].pbc
    out["test_pos"] = test.positions
    tdata = m.utilities.convert_train_data(labelled, des)
    gp = m.GP(kernel=m.RBF_mb(para=[2.0, 0.8], zeta=2.0), descriptor=des, noise_e=0.002, noise_f=0.1, log_file="/tmp/gpr_golden.log")
    with contextlib.redirect_stdout(io.StringIO()):
        gp.fit(TrainData=tdata, opt=False, show=False)
    # gp.predict_structure(test, stress=True) itself fails in the reference: it hands kff_C an object ndarray that
    # kff_C does not convert (rbf_kernel.py:213-218 vs gaussianprocess.py:854).  The same steps with the force data
    # as a list (gaussianprocess.py:846-891) pin the stress prediction:
    d = des.calculate(test)
    ele = np.array([rh.SYMBOLS.index(sym) for sym in d['elements']])
    data = {"energy": lt([(d['x'], ele)], mode='energy'), "force": []}
    for i in range(len(test)):
        ids = np.argwhere(d['seq'][:, 1] == i).flatten()
        _i = d['seq'][ids, 0]
        _rdxdr = d['rdxdr'][ids].reshape(len(ids), d['x'].shape[1], 9)[:, :, [0, 4, 8, 1, 2, 5]]
        data["force"].append((d['x'][_i, :], np.concatenate((d['dxdr'][ids], _rdxdr), axis=2), ele[_i]))
    K_trans, K_trans1 = gp.kernel.k_total_with_stress(data, gp.get_train_x(), 1e-12)
    y_mean = K_trans.dot(gp.alpha_)[:, 0]
    E = y_mean[0] * len(test)
    F = y_mean[1:].reshape([len(test), 3])
    S = K_trans1.dot(gp.alpha_)[:, 0].reshape([len(test), 6])
    out["pred_E"], out["pred_F"], out["pred_S"] = E, F, S
    np.savez_compressed(os.path.join(HERE, "stress.npz"), **out)
    print("stress.npz", os.path.getsize(os.path.join(HERE, "stress.npz")) // 1024, "KiB", "S", S.shape)


if __name__ == "__main__":
    main()
