"""Variance of m test rows against the S5 training set: explicit inverse (cuBLAS gemm) vs Cholesky factor (cuBLAS trsm).
Decides GP.CHOL_VARIANCE_MIN_ROWS.   python tools/predict_routes.py [n_struct]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpr_calculator_b200 import device as gdev, synthetic as syn   # noqa: E402
from gpr_calculator_b200.SO3 import SO3                             # noqa: E402
from gpr_calculator_b200.gaussianprocess import GP                  # noqa: E402
from gpr_calculator_b200.kernels import RBF_mb                      # noqa: E402

n_struct = int(sys.argv[1]) if len(sys.argv) > 1 else 340
des = SO3(nmax=3, lmax=4, rcut=5.0)
labelled = syn.structures(n_struct, 2, 2000)
E_dev, F_dev = syn.packed_from_batch(des, [a for a, _, _ in labelled])
gp = GP(kernel=RBF_mb(para=[1.0, 0.1], zeta=2.0), descriptor=des, noise_e=0.002, noise_f=0.1, log_file=None)
gp.train_x = {"energy": gdev.Pack(E_dev[0], E_dev[1], E_dev[2]), "force": gdev.Pack(F_dev[0], F_dev[2], F_dev[3], dxdr=F_dev[1])}
gp.y_train = syn.targets(labelled)
K, _, _ = gp._build_K(grad=False)
rows = K[:4096].clone()                      # realistic K* rows: rows of the training covariance itself
gp._alpha_dev = gp._factor(K, 0.002, 0.1)
gp._L_dev, gp._Kinv_dev = K, None
gp.set_K_inv()
N = K.shape[0]
for m in (97, 291, 582, 1067, 2037, 3104):
    Ks = rows[:m].contiguous()
    diag = torch.full((m,), 1e3, dtype=torch.float64, device="cuda")
    res = {}
    for route in ("inverse", "trmm", "chol"):
        os.environ["GPRB_VARIANCE_ROUTE"] = "inverse" if route == "trmm" else route
        os.environ.pop("GPRB_PREDICT_TRMM", None)
        if route == "trmm":
            os.environ["GPRB_PREDICT_TRMM"] = "1"
        for _ in range(2):
            mean, var = gp._mean_var(Ks, diag)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            mean, var = gp._mean_var(Ks, diag)
        e1.record()
        torch.cuda.synchronize()
        res[route] = (e0.elapsed_time(e1) / 5, var.cpu().numpy())
    d = np.abs(res["inverse"][1] - res["chol"][1]).max() / np.abs(1e3 - res["inverse"][1]).max()
    d2 = np.abs(res["inverse"][1] - res["trmm"][1]).max() / np.abs(1e3 - res["inverse"][1]).max()
    print("m=%5d N=%d  inverse (gemm) %.2f ms   inverse, half product (trmm) %.2f ms   chol (trsm) %.2f ms   max rel diff of "
          "k*^T K^-1 k*: chol %.2e, trmm %.2e" % (m, N, res["inverse"][0], res["trmm"][0], res["chol"][0], d, d2), flush=True)
