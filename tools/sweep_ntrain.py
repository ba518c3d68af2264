"""BASELINE.json metric as a function of N_train on one B200: covariance build (K and dK/dl) GFLOP/s and wall time,
end-to-end likelihood evaluation, and E/F/sigma predictions per second, for growing prefixes of the S5 workload.

    python tools/sweep_ntrain.py [out.json]      # one JSON line per N_train
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gpr_calculator_b200 import device as gdev, synthetic as syn   # noqa: E402
from gpr_calculator_b200.SO3 import SO3                             # noqa: E402
from gpr_calculator_b200.gaussianprocess import GP                  # noqa: E402
from gpr_calculator_b200.kernels import RBF_mb                      # noqa: E402

SIGMA, ELL = 1.0, 0.1


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    des = SO3(nmax=3, lmax=4, rcut=5.0)
    tests = [a for a, _, _ in syn.structures(64, 2, 3000)]
    lines = []
    for n_struct in (10, 42, 85, 170, 340):
        labelled = syn.structures(n_struct, 2, 2000)
        E_dev, F_dev = syn.packed_from_batch(des, [a for a, _, _ in labelled])
        e_pack = gdev.Pack(E_dev[0], E_dev[1], E_dev[2])
        f_pack = gdev.Pack(F_dev[0], F_dev[2], F_dev[3], dxdr=F_dev[1])
        N = e_pack.n_groups + 3 * f_pack.n_groups
        p_ff = syn.pair_counts(F_dev[2].cpu().numpy(), F_dev[3], symmetric=True)
        p_ee = syn.pair_counts(E_dev[1].cpu().numpy(), E_dev[2], symmetric=False)
        p_ef = e_pack.pair_count(f_pack)
        flops = 32.0 * 30 * p_ff + 8.0 * 30 * p_ef + 2.0 * 30 * p_ee
        gp = GP(kernel=RBF_mb(para=[SIGMA, ELL], zeta=2.0), descriptor=des, noise_e=0.002, noise_f=0.1, log_file=None)
        gp.train_x = {"energy": e_pack, "force": f_pack}
        gp.y_train = syn.targets(labelled)
        theta = np.array([SIGMA, ELL])
        for _ in range(3):
            gp._build_K(grad=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3 if n_struct >= 170 else 10
        e0.record()
        for _ in range(reps):
            gp._build_K(grad=True)
        e1.record()
        torch.cuda.synchronize()
        build_ms = e0.elapsed_time(e1) / reps
        gp.log_marginal_likelihood(theta, eval_gradient=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            gp.log_marginal_likelihood(theta, eval_gradient=True)
        torch.cuda.synchronize()
        lml_ms = (time.perf_counter() - t0) / reps * 1e3
        K, _, _ = gp._build_K(grad=False)
        gp._alpha_dev = gp._factor(K, 0.002, 0.1)
        gp._L_dev, gp._Kinv_dev = K, None
        gp.set_K_inv()
        gp.predict_structures(tests[:32], return_std=True, f_tol=1e-12, batch=32)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        gp.predict_structures(tests, return_std=True, f_tol=1e-12, batch=32)
        torch.cuda.synchronize()
        per_s = len(tests) / (time.perf_counter() - t0)
        gp.predict_structure(tests[0], stress=False, return_std=True, f_tol=1e-12)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for a in tests[:8]:
            gp.predict_structure(a, stress=False, return_std=True, f_tol=1e-12)
        torch.cuda.synchronize()
        single_ms = (time.perf_counter() - t0) / 8 * 1e3
        line = {"n_structures": n_struct, "N_train": N, "force_rows": int(F_dev[0].shape[0]), "k_build_ms": round(build_ms, 3),
                "k_build_gflops": round(flops / build_ms * 1e-6, 1), "lml_grad_call_ms": round(lml_ms, 3),
                "predict_structures_per_s": round(per_s, 2), "predict_single_call_ms": round(single_ms, 3)}
        print(json.dumps(line), flush=True)
        lines.append(line)
        del gp, K, e_pack, f_pack, E_dev, F_dev
        torch.cuda.empty_cache()
    if out_path:
        with open(out_path, "w") as fh:
            for line in lines:
                fh.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
