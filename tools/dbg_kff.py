"""Small staged debug run of the covariance kernels (each stage prints before it runs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import make_force, make_energy, rel_err
from gpr_calculator_b200.kernels import rbf_kernel as rk
from gpr_calculator_b200.utilities import list_to_tuple
from oracle import kernels as ok
ok.build(ref=False, port=True)
O = ok.RBFOracle("port")
rng = np.random.default_rng(1)
for (n1, n2, lo, hi) in ((1, 1, 3, 3), (2, 3, 5, 11), (9, 7, 3, 40), (13, 5, 1, 2), (3, 4, 70, 150), (40, 50, 20, 36)):
    F1, F2 = list_to_tuple(make_force(rng, n1, lo=lo, hi=hi)), list_to_tuple(make_force(rng, n2, lo=lo, hi=hi))
    E1 = list_to_tuple(make_energy(rng, 3, lo=lo, hi=hi + 20), mode="energy")
    for grad in (False, True):
        print("kff", n1, n2, lo, hi, grad, flush=True)
        got = rk.kff_C(F1, F2, 1.3, 0.7, 2.0, grad=grad); torch.cuda.synchronize()
        ref = O.kff_C(F1, F2, 1.3, 0.7, 2.0, grad=grad)
        got, ref = (got, ref) if grad else ((got,), (ref,))
        print("   err", [rel_err(a, b) for a, b in zip(got, ref)], flush=True)
        print("kef", flush=True)
        got = rk.kef_C(E1, F2, 1.3, 0.7, 2.0, grad=grad); torch.cuda.synchronize()
        ref = O.kef_C(E1, F2, 1.3, 0.7, 2.0, grad=grad)
        got, ref = (got, ref) if grad else ((got,), (ref,))
        print("   err", [rel_err(a, b) for a, b in zip(got, ref)], flush=True)
print("done")
