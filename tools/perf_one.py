"""One gprb_kff launch configuration for profiling: python tools/perf_one.py NF rows_lo rows_hi grad(0/1) mode reps"""
import sys, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpr_calculator_b200 import _lib
from gpr_calculator_b200.device import Pack, empty, ptr, stream
NF, lo, hi, grad, mode, reps = (int(v) for v in sys.argv[1:7])
rng = np.random.default_rng(0)
ind = rng.integers(lo, hi + 1, size=NF)
R = int(ind.sum())
x = torch.rand(R, 30, dtype=torch.float64, device='cuda') + 0.5
dx = torch.randn(R, 30, 3, dtype=torch.float64, device='cuda')
ele = torch.full((R,), 29, dtype=torch.int32, device='cuda')
p = Pack(x, ele, ind, dxdr=dx)
K = empty(3 * NF, 3 * NF); dK = empty(3 * NF, 3 * NF) if grad else None
def go():
    _lib.call("gprb_kff", _lib.RBF, p.handle, p.handle, 1.0, 0.5, 2.0, 0 if grad else 1, 1e-10, mode, 0, NF,
              ptr(K), 3 * NF, ptr(dK), 3 * NF, stream())
go(); torch.cuda.synchronize()
best = 1e9
for _ in range(reps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); go(); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
pairs = p.pair_count(p)
print(json.dumps({"NF": NF, "rows": R, "ms": best, "alg_tflops_fullblock": 32 * 30 * pairs / best * 1e-9}))
