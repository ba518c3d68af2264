"""Golden vectors for the training-set bookkeeping either side of the hot path, from the UNMODIFIED reference
(build container only; same harness and structures as gen_golden.py):

    python tests/golden/gen_golden_selection.py   ->   tests/golden/selection.npz

  * GP.add_structure (gaussianprocess.py:921-1002): which force centres of a new labelled structure enter the training
    set, in what order, for the default thresholds, a tight force threshold, an N_max cap and an untrained model;
    the (E, E1, E_std, F, F1, F_std) error tuple it returns; the queue counters afterwards.
  * GP.predict(X, return_cov=True) (gaussianprocess.py:363-366).
  * CUR (gaussianprocess.py:1165-1182) on covariance blocks with exactly duplicated training points.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, HERE)

from oracle import ref_harness as rh   # noqa: E402
from gen_golden import slab, toy_labels   # noqa: E402


def main():
    os.chdir("/tmp")
    m = rh.modules()
    out = {}
    des = m.SO3(nmax=3, lmax=4, rcut=5.0)
    train = [slab(100 + k) for k in range(3)]
    labelled = [(at,) + toy_labels(at, 300 + k) for k, at in enumerate(train)]
    new = slab(400, n_fixed=0)
    new.positions[12] += np.array([0.4, -0.3, 0.25])          # a displaced adatom: some force rows exceed the thresholds
    E_new, F_new = toy_labels(new, 401)
    out["new_pos"], out["new_E"], out["new_F"] = new.positions, E_new, F_new

    def fitted():
        gp = m.GP(kernel=m.RBF_mb(para=[2.0, 0.8], zeta=2.0), descriptor=des, noise_e=0.002, noise_f=0.1, log_file="/tmp/gpr_golden.log")
        with contextlib.redirect_stdout(io.StringIO()):
            gp.fit(TrainData=m.utilities.convert_train_data(labelled, des), opt=False, show=False)
        return gp

    cases = {"default": {}, "tight": {"tol_f_var": 0.05}, "capped": {"tol_f_var": 0.05, "N_max": 2}, "noforce": {"add_force": False}}
    for tag, kw in cases.items():
        gp = fitted()
        with contextlib.redirect_stdout(io.StringIO()):
            pts, n_pts, err = gp.add_structure((new.copy(), float(E_new), F_new.copy()), **kw)
        out[tag + "_force_in"] = np.array(gp.train_db[-1][4], dtype=np.int64)
        out[tag + "_n_pts"] = n_pts
        out[tag + "_counters"] = np.array([gp.N_energy, gp.N_forces, gp.N_energy_queue, gp.N_forces_queue, gp.N_queue])
        out[tag + "_err_E"] = np.array([err[0], err[1], err[2]])
        out[tag + "_err_F"], out[tag + "_err_F1"], out[tag + "_err_Fstd"] = np.asarray(err[3]), np.asarray(err[4]), np.asarray(err[5])
        out[tag + "_y_train"] = gp.y_train
    # untrained model: every centre is a candidate, new_pt() removes the near-duplicates
    gp = m.GP(kernel=m.RBF_mb(para=[2.0, 0.8], zeta=2.0), descriptor=des, noise_e=0.002, noise_f=0.1, log_file="/tmp/gpr_golden.log")
    with contextlib.redirect_stdout(io.StringIO()):
        for k, (at, E, F) in enumerate(labelled):
            gp.add_structure((at.copy(), float(E), F.copy()))
            out["fresh%d_force_in" % k] = np.array(gp.train_db[-1][4], dtype=np.int64)
    out["fresh_counters"] = np.array([gp.N_energy, gp.N_forces, gp.N_energy_queue, gp.N_forces_queue, gp.N_queue])

    # predict(return_cov=True) on the new structure's energy + three force centres
    gp = fitted()
    d = m.utilities.convert_train_data([(new.copy(), float(E_new), F_new.copy())], des)
    # (packed tuples: k_total(X) of :365 unpacks them, rbf_kernel.py:30)
    lt = m.utilities.list_to_tuple
    X = {"energy": lt([(d["energy"][0][0], d["energy"][0][2])], mode="energy"), "force": lt([(f[0], f[1], f[3]) for f in d["force"][9:12]])}
    y_mean, y_cov = gp.predict(X, return_cov=True)
    out["cov_mean"], out["cov_cov"] = y_mean, y_cov

    # CUR on blocks with exact duplicates (third structure = first)
    lab2 = [labelled[0], labelled[1], labelled[0]]
    gp = m.GP(kernel=m.RBF_mb(para=[2.0, 0.8], zeta=2.0), descriptor=des, noise_e=0.002, noise_f=0.1, log_file="/tmp/gpr_golden.log")
    with contextlib.redirect_stdout(io.StringIO()):
        gp.fit(TrainData=m.utilities.convert_train_data(lab2, des), opt=False, show=False)
    K = gp.kernel.k_total(gp.train_x)
    n_e = len(gp.train_x["energy"][-1])
    out["cur_K"] = K
    out["cur_n_e"] = n_e
    for tol in (1e-8, 1e-4):
        out["cur_e_%g" % tol] = m.gp_module.CUR(K[:n_e, :n_e], tol)
        out["cur_f_%g" % tol] = m.gp_module.CUR(K[n_e:, n_e:], tol)
        # the leverage scores themselves (the projector diagonal onto the low eigen-space: unique when the cut is clean)
        L, U = np.linalg.eigh(K[n_e:, n_e:])
        out["cur_f_omega_%g" % tol] = (U[:, L < tol] ** 2).sum(axis=1)
    path = os.path.join(HERE, "selection.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")
    for tag in cases:
        print(tag, "force_in", out[tag + "_force_in"], "n_pts", out[tag + "_n_pts"], "counters", out[tag + "_counters"])
    for k in range(3):
        print("fresh", k, out["fresh%d_force_in" % k])
    print("cur", {k: (v if v.size < 12 else v.shape) for k, v in out.items() if k.startswith("cur_e") or k.startswith("cur_f_1")})


if __name__ == "__main__":
    main()
