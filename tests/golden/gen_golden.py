"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference: its Python + its C++ compiled into oracle/_ref) under the stubs of
oracle/ref_harness.py.  Runs only in the build container (the GPU box has no /root/reference).

    python tests/golden/gen_golden.py

Outputs (float64, small):
    kernels.npz   packed synthetic inputs + kee/kef/kff (+grad) of RBF and Dot, RBF_mb / Dot_mb
                  k_total, k_total_with_grad, diag
    so3.npz       three small structures + x / dxdr / seq of reference SO3.calculate
    gp.npz        a 3-structure training set -> reference GP: LML + gradient, fit(opt=False) alpha,
                  predict_structure E/F/std, fit(opt=True) trajectory
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

from oracle import ref_harness as rh   # noqa: E402
from helpers import make_force, make_energy   # noqa: E402


def slab(seed, n_fixed=8):
    """Al12Au slab like examples/database/initial.traj: 12 Al in three layers + Au adatom, pbc TTF."""
    rng = np.random.default_rng(seed)
    a = 2.8638
    pos = []
    for layer, z in enumerate((4.0, 6.025, 8.05)):
        for ix in range(2):
            for iy in range(2):
                off = 0.5 * a * (layer % 2)
                pos.append([ix * a + off, iy * a + off, z])
    pos.append([1.4, 1.4, 10.0 + 0.2 * rng.normal()])
    pos = np.array(pos) + rng.normal(scale=0.05, size=(13, 3))
    cons = [rh.FixAtoms(list(range(n_fixed)))] if n_fixed else []
    return rh.Atoms([13] * 12 + [79], pos, np.diag([2 * a, 2 * a, 13.75]), pbc=(True, True, False), constraints=cons)


def toy_labels(atoms, seed):
    """Smooth synthetic labels (values are irrelevant to the covariance path)."""
    rng = np.random.default_rng(seed)
    E = -3.0 * len(atoms) + rng.normal(scale=0.05)
    F = rng.normal(scale=0.3, size=(len(atoms), 3))
    return float(E), F


def main():
    os.chdir("/tmp")
    m = rh.modules()
    lt = m.utilities.list_to_tuple

    # ---- kernels.npz -------------------------------------------------------------------------
    rng = np.random.default_rng(20261018)
    F1 = lt(make_force(rng, 6, lo=2, hi=11)); F2 = lt(make_force(rng, 5, lo=1, hi=13, zero_rows=1))
    E1 = lt(make_energy(rng, 4, lo=3, hi=12), mode="energy"); E2 = lt(make_energy(rng, 3, lo=3, hi=20), mode="energy")
    out = {}
    for name, t in (("F1", F1), ("F2", F2)):
        out[name + "_x"], out[name + "_dxdr"], out[name + "_ele"], out[name + "_ind"] = t[0], t[1], t[2], np.array(t[3])
    for name, t in (("E1", E1), ("E2", E2)):
        out[name + "_x"], out[name + "_ele"], out[name + "_ind"] = t[0], t[1], np.array(t[2])
    sig, l = 1.3, 0.7
    for zeta in (2.0, 3.0):
        z = "z%d" % int(zeta)
        out["rbf_kee_" + z] = m.rbf_kernel.kee_C(E1, E2, sig, l, zeta)
        for k, v in zip(("K", "Ks", "Kl"), m.rbf_kernel.kee_C(E1, E2, sig, l, zeta, grad=True)):
            out["rbf_kee_grad_%s_%s" % (k, z)] = v
        out["rbf_kef_" + z] = m.rbf_kernel.kef_C(E1, F2, sig, l, zeta)
        for k, v in zip(("K", "Ks", "Kl"), m.rbf_kernel.kef_C(E1, F2, sig, l, zeta, grad=True)):
            out["rbf_kef_grad_%s_%s" % (k, z)] = v
        out["rbf_kff_" + z] = m.rbf_kernel.kff_C(F1, F2, sig, l, zeta, tol=1e-12)
        for k, v in zip(("K", "Ks", "Kl"), m.rbf_kernel.kff_C(F1, F2, sig, l, zeta, grad=True)):
            out["rbf_kff_grad_%s_%s" % (k, z)] = v
        out["dot_kee_" + z] = m.dot_kernel.kee_C(E1, E2, 2.0, 1.5, zeta)
        out["dot_kef_" + z] = m.dot_kernel.kef_C(E1, F2, 2.0, 1.5, zeta)
        out["dot_kff_" + z] = m.dot_kernel.kff_C(F1, F2, 2.0, 1.5, zeta)
    # a loose pair cut that actually removes pairs
    out["rbf_kff_tol_l02"] = m.rbf_kernel.kff_C(F1, F2, 1.0, 0.2, 2.0, tol=1.0)
    out["params"] = np.array([sig, l])
    # kernel objects on a training-like dict
    data = {"energy": E1, "force": F1}
    data2 = {"energy": E2, "force": F2}
    rbf = m.RBF_mb(para=[sig, l], zeta=2)
    out["RBF_k_total"] = rbf.k_total(data)
    out["RBF_k_total_rect"] = rbf.k_total(data2, data, f_tol=1e-12)
    K, dK = rbf.k_total_with_grad(data)
    out["RBF_k_grad_K"], out["RBF_k_grad_dK"] = K, dK
    tl = m.utilities.tuple_to_list
    # force data as a list, the form predict_structure passes (gaussianprocess.py:854-870); with the
    # packed tuple the reference takes NF = len(tuple) = 4 (RBF_mb.py:87) and truncates
    out["RBF_diag"] = rbf.diag({"energy": E1, "force": tl(F1)})
    dot = m.Dot_mb(para=[2.0, 1.5], zeta=3)
    out["Dot_k_total"] = dot.k_total(data)
    K, dK = dot.k_total_with_grad(data)
    out["Dot_k_grad_K"], out["Dot_k_grad_dK"] = K, dK
    out["Dot_diag"] = dot.diag({"energy": tl(E1, mode="energy"), "force": tl(F1)})
    np.savez_compressed(os.path.join(HERE, "kernels.npz"), **out)

    # ---- so3.npz ------------------------------------------------------------------------------
    out = {}
    strucs = [slab(11), rh.Atoms([29] * 4, np.array([[0, 0, 0], [1.8, 1.8, 0], [1.8, 0, 1.8], [0, 1.8, 1.8]]) + 0.03 *
                                 np.random.default_rng(3).normal(size=(4, 3)), np.eye(3) * 3.61),
              rh.Atoms([1, 1, 16, 46, 46], [[0, 0, 0.3], [0.2, 1.5, 0.1], [0.9, 0.7, 0.9], [3.0, 3.0, 3.0], [5.5, 3.1, 2.7]],
                       [[8.0, 0.5, 0], [0, 8.0, 0], [0, 0, 9.0]], pbc=(True, True, True))]
    prms = [(3, 4, 5.0, 2.0), (3, 4, 5.0, 2.0), (2, 3, 4.0, 1.5)]
    for k, (at, prm) in enumerate(zip(strucs, prms)):
        r = m.SO3(nmax=prm[0], lmax=prm[1], rcut=prm[2], alpha=prm[3]).calculate(at)
        out["s%d_numbers" % k], out["s%d_pos" % k], out["s%d_cell" % k], out["s%d_pbc" % k] = at.numbers, at.positions, np.asarray(at.cell), at.pbc
        out["s%d_prm" % k] = np.array(prm)
        out["s%d_x" % k], out["s%d_dxdr" % k], out["s%d_seq" % k] = r["x"], r["dxdr"], r["seq"]
    np.savez_compressed(os.path.join(HERE, "so3.npz"), **out)

    # ---- gp.npz -------------------------------------------------------------------------------
    out = {}
    train = [slab(100 + k) for k in range(3)]
    test = slab(200)
    des = m.SO3(nmax=3, lmax=4, rcut=5.0)
    labelled = []
    for k, at in enumerate(train):
        E, F = toy_labels(at, 300 + k)
        labelled.append((at, E, F))
        out["t%d_pos" % k], out["t%d_E" % k], out["t%d_F" % k] = at.positions, E, F
    out["numbers"], out["cell"], out["pbc"], out["fixed"] = train[0].numbers, np.asarray(train[0].cell), train[0].pbc, np.arange(8)
    out["test_pos"] = test.positions
    tdata = m.utilities.convert_train_data(labelled, des)
    gp = m.GP(kernel=m.RBF_mb(para=[1.0, 0.1], zeta=2.0), descriptor=des, noise_e=0.002, noise_f=0.1, log_file="/tmp/gpr_golden.log")
    with contextlib.redirect_stdout(io.StringIO()):
        gp.fit(TrainData=tdata, opt=False, show=False)
    out["N"] = len(gp.y_train)
    out["y_train"] = gp.y_train
    for tag, prm in (("a", [1.0, 0.1]), ("b", [2.0, 0.8])):
        lml, grad = gp.log_marginal_likelihood(np.array(prm), eval_gradient=True)
        out["lml_" + tag], out["lml_grad_" + tag] = lml, grad
    with contextlib.redirect_stdout(io.StringIO()):
        gp.kernel.update([2.0, 0.8])
        gp.fit(opt=False, show=False)
    out["alpha_b"] = gp.alpha_
    out["K_b"] = gp.kernel.k_total(gp.train_x)
    E, F, S, E_std, F_std = gp.predict_structure(test, stress=False, return_std=True, f_tol=1e-12)
    out["pred_E"], out["pred_F"], out["pred_E_std"], out["pred_F_std"] = E, F, E_std, F_std
    # (validate_data(return_std=True) on the packed training tuple crashes in the reference:
    #  RBF_mb.diag takes NF = len(tuple) = 4, RBF_mb.py:87 — so only the means are pinned here)
    Ev, Ep, Fv, Fp = gp.validate_data()
    out["val_E_pred"], out["val_F_pred"] = Ep, Fp
    # optimisation trajectory from the reference's initial guess
    buf = io.StringIO()
    gp.kernel.update([1.0, 0.1])
    with contextlib.redirect_stdout(buf):
        gp.fit(opt=True, show=True, maxiter=10)
    losses = [[float(v) for v in line.split()[1:]] for line in buf.getvalue().splitlines() if line.startswith("Loss:")]
    out["opt_trace"] = np.array(losses)
    out["opt_params"] = np.array(gp.kernel.parameters())
    E, F, S, E_std, F_std = gp.predict_structure(test, stress=False, return_std=True, f_tol=1e-12)
    out["opt_pred_E"], out["opt_pred_F"], out["opt_pred_E_std"], out["opt_pred_F_std"] = E, F, E_std, F_std
    # Dot kernel GP
    gpd = m.GP(kernel=m.Dot_mb(para=[2, 2.0], zeta=2.0), descriptor=des, noise_e=0.002, noise_f=0.1, log_file="/tmp/gpr_golden.log")
    with contextlib.redirect_stdout(io.StringIO()):
        gpd.fit(TrainData=tdata, opt=False, show=False)
    lml, grad = gpd.log_marginal_likelihood(np.array([2.0, 2.0]), eval_gradient=True)
    out["dot_lml"], out["dot_lml_grad"] = lml, grad
    out["dot_alpha"] = gpd.alpha_
    np.savez_compressed(os.path.join(HERE, "gp.npz"), **out)
    for f in ("kernels.npz", "so3.npz", "gp.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
