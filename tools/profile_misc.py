"""One pass of each non-K_ff kernel family for ncu (python tools/profile_misc.py PHASE):
  so3      SO3.calculate_batch of 64 Cu32 structures (so3_neighbors / so3_radial / so3_power)
  pack     gprb_pack_create of the S5 force rows (prep_rows_kernel) from device arrays
  kef      K_ef / K_fe with gradient at S5 (cov_mma_kernel NB = 1)
  kee      K_ee with gradient of the Pd4/MgO-shaped energy set: 155 groups x 220 rows, three species (cov_mma_kernel NB = 0)
  lml      gprb_lml_eval at N = 9 700 (trace_kernel, final sums; library potrf / trsm in between)
  predict  K* rows + mean / variance of a 32-structure batch at N = 9 700 (two-stage K_ff, predict_rows kernels)
Each phase runs its work twice: a warm-up pass, then the pass to capture between cudaProfilerStart / cudaProfilerStop
(ncu --profile-from-start off -k regex:<kernels of the phase> -c <limit>; tools/r02_ncu_misc.sh)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpr_calculator_b200 import _lib, device as gdev, synthetic as syn   # noqa: E402
from gpr_calculator_b200.SO3 import SO3                                   # noqa: E402
from gpr_calculator_b200.gaussianprocess import GP                        # noqa: E402
from gpr_calculator_b200.kernels import RBF_mb                            # noqa: E402

phase = sys.argv[1]
des = SO3(nmax=3, lmax=4, rcut=5.0)
st = gdev.stream


def training(n):
    labelled = syn.structures(n, 2, 2000)
    E_dev, F_dev = syn.packed_from_batch(des, [a for a, _, _ in labelled])
    return labelled, E_dev, F_dev


def count():
    return _lib.load().gprb_launch_count()


def capture(k):
    """profiler range around the second pass"""
    if k == 1:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()


def done():
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if phase == "so3":
    atoms = [a for a, _, _ in syn.structures(64, 2, 2000)]
    for k in range(2):
        capture(k)
        n0 = count()
        des.calculate_batch(atoms, to_host=False)
        torch.cuda.synchronize()
        print("launches", count() - n0)
    done()
elif phase == "pack":
    _, E_dev, F_dev = training(340)
    for k in range(2):
        capture(k)
        n0 = count()
        p = gdev.Pack(F_dev[0], F_dev[2], F_dev[3], dxdr=F_dev[1])
        torch.cuda.synchronize()
        print("launches", count() - n0, "rows", p.n_rows)
    done()
elif phase == "kef":
    _, E_dev, F_dev = training(340)
    e, f = gdev.Pack(E_dev[0], E_dev[1], E_dev[2]), gdev.Pack(F_dev[0], F_dev[2], F_dev[3], dxdr=F_dev[1])
    NE, NF = e.n_groups, f.n_groups
    Kef, Kfe = gdev.empty(NE, 3 * NF), gdev.empty(3 * NF, NE)
    dKef, dKfe = gdev.empty(NE, 3 * NF), gdev.empty(3 * NF, NE)
    for k in range(2):
        capture(k)
        _lib.call("gprb_kef", _lib.RBF, e.handle, f.handle, 1.0, 0.1, 2.0, 0, NF, gdev.ptr(Kef), 3 * NF, gdev.ptr(Kfe), NE,
                  gdev.ptr(dKef), 3 * NF, gdev.ptr(dKfe), NE, st())
        torch.cuda.synchronize()
    done()
    print("pairs", e.pair_count(f))
elif phase == "kee":
    rng = np.random.default_rng(0)
    base = np.abs(rng.normal(size=30)) + 0.5
    G, n = 155, 220
    X = torch.as_tensor(base[None, :] + 0.3 * rng.normal(size=(G * n, 30)), device="cuda")
    ele = torch.as_tensor(np.tile(np.array([12] * 108 + [8] * 108 + [46] * 4, dtype=np.int32), G), device="cuda")
    e = gdev.Pack(X, ele, [n] * G)
    K, dK = gdev.empty(G, G), gdev.empty(G, G)
    for k in range(2):
        capture(k)
        _lib.call("gprb_kee", _lib.RBF, e.handle, e.handle, 1.0, 0.5, 2.0, 0, G, gdev.ptr(K), G, gdev.ptr(dK), G, st())
        torch.cuda.synchronize()
    done()
    print("pairs", e.pair_count(e))
elif phase in ("lml", "predict"):
    labelled, E_dev, F_dev = training(100)
    gp = GP(kernel=RBF_mb(para=[1.0, 0.1], zeta=2.0), descriptor=des, noise_e=0.002, noise_f=0.1, log_file=None)
    gp.train_x = {"energy": gdev.Pack(E_dev[0], E_dev[1], E_dev[2]), "force": gdev.Pack(F_dev[0], F_dev[2], F_dev[3], dxdr=F_dev[1])}
    y = syn.targets(labelled)
    gp.train_y = {"energy": list(y[:100, 0]), "force": y[100:, 0].reshape(-1, 3)}
    gp.update_y_train()
    gp.N_energy, gp.N_forces = 100, 3200
    if phase == "lml":
        for k in range(2):
            capture(k)
            print(gp.log_marginal_likelihood(np.array([1.0, 0.1]), eval_gradient=True))
        done()
    else:
        gp.fit(opt=False, show=False)
        tests = [a for a, _, _ in syn.structures(32, 2, 3000)]
        for k in range(2):
            capture(k)
            os.environ["GPRB_VARIANCE_ROUTE"] = "chol"
            gp.predict_structures(tests, return_std=True, f_tol=1e-12, batch=32)
            os.environ["GPRB_VARIANCE_ROUTE"] = "inverse"
            gp.predict_structures(tests, return_std=True, f_tol=1e-12, batch=32)
            torch.cuda.synchronize()
        done()
else:
    raise SystemExit("unknown phase " + phase)
