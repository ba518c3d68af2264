"""Per-step wall times of the end-to-end LML call of bench.py (diagnostic: where do host stalls come from)."""
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

n_struct = int(sys.argv[1]) if len(sys.argv) > 1 else 340
labelled = syn.structures(n_struct, 2, 2000)
des = SO3(nmax=3, lmax=4, rcut=5.0)
E_dev, F_dev = syn.packed_from_batch(des, [a for a, _, _ in labelled])
y = syn.targets(labelled)
E_host, k1 = syn.to_host(E_dev, pin=True)
F_host, k2 = syn.to_host(F_dev, pin=True)
del E_dev, F_dev
gp = GP(kernel=RBF_mb(para=[1.0, 0.1], zeta=2.0), descriptor=des, noise_e=0.002, noise_f=0.1, log_file=None)
gp.train_x = {"energy": E_host, "force": F_host}
gp.y_train = y
theta = np.array([1.0, 0.1])
for it in range(10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gdev.clear_cache()
    t1 = time.perf_counter()
    e, f = gdev.packs_of(gp.train_x)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    res = gp.log_marginal_likelihood(theta, eval_gradient=True)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    free, total = torch.cuda.mem_get_info()
    print("step %d: clear %.1f ms, packs %.1f ms, lml %.1f ms, total %.1f ms | free %.1f GB torch reserved %.1f GB"
          % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3, free / 2**30,
             torch.cuda.memory_reserved() / 2**30), flush=True)
    del e, f
