import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from helpers import *
from oracle import kernels as ok
from gpr_calculator_b200.kernels import rbf_kernel as rk, dot_kernel as dk
from gpr_calculator_b200.utilities import list_to_tuple
rng = np.random.default_rng(1)
F1 = list_to_tuple(make_force(rng, 9)); F2 = list_to_tuple(make_force(rng, 7, zero_rows=1))
E1 = list_to_tuple(make_energy(rng, 5), mode='energy'); E2 = list_to_tuple(make_energy(rng, 4), mode='energy')
O = ok.RBFOracle('ref'); OD = ok.DotOracle('ref')
sig, l = 1.3, 0.7
for zeta in (2.0, 3.0, 2.5):
  for grad in (False, True):
    a = rk.kff_C(F1, F2, sig, l, zeta, grad=grad, tol=1e-12); b = O.kff_C(F1, F2, sig, l, zeta, grad=grad, tol=1e-12)
    a = a if grad else (a,); b = b if grad else (b,)
    print('kff', zeta, grad, [rel_err(x, y) for x, y in zip(a, b)])
    a = rk.kef_C(E1, F2, sig, l, zeta, grad=grad); b = O.kef_C(E1, F2, sig, l, zeta, grad=grad)
    a = a if grad else (a,); b = b if grad else (b,)
    print('kef', zeta, grad, [rel_err(x, y) for x, y in zip(a, b)])
    a = rk.kee_C(E1, E2, sig, l, zeta, grad=grad); b = O.kee_C(E1, E2, sig, l, zeta, grad=grad)
    a = a if grad else (a,); b = b if grad else (b,)
    print('kee', zeta, grad, [rel_err(x, y) for x, y in zip(a, b)])
  print('dot kff', zeta, rel_err(dk.kff_C(F1, F2, 2.0, 1.5, zeta), OD.kff_C(F1, F2, 2.0, 1.5, zeta)))
  print('dot kef', zeta, rel_err(dk.kef_C(E1, F2, 2.0, 1.5, zeta), OD.kef_C(E1, F2, 2.0, 1.5, zeta)))
  print('dot kee', zeta, rel_err(dk.kee_C(E1, E2, 2.0, 1.5, zeta), OD.kee_C(E1, E2, 2.0, 1.5, zeta)))
# big groups (split) and tiny
F3 = list_to_tuple(make_force(rng, 3, lo=70, hi=150)); F4 = list_to_tuple(make_force(rng, 11, lo=1, hi=5))
for A, B in ((F3, F4), (F4, F3), (F3, F3)):
    print('kff big/tiny', rel_err(rk.kff_C(A, B, sig, l, 2.0), O.kff_C(A, B, sig, l, 2.0)))
print('kef big', rel_err(rk.kef_C(E1, F3, sig, l, 2.0), O.kef_C(E1, F3, sig, l, 2.0)))
