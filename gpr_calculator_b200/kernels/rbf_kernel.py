"""Function-level RBF covariance blocks on the B200 (mirror of gpr_calc/kernels/rbf_kernel.py).

``kee_C / kef_C / kff_C`` keep the reference signatures and return numpy arrays, so parity tests
read like calls into the reference's cffi wrappers; the work is done by libgpr_b200.so.
"""
import torch

from .. import _lib
from ..device import Pack, energy_pack, force_pack, empty, ptr, stream, require_cuda, c_vp


def _host(t):
    return t.cpu().numpy()


def kee_C(X1, X2, sigma=1.0, l=1.0, zeta=2.0, grad=False):
    """Energy-energy block [m1, m2]; grad=True also returns dK/dsigma, dK/dl (rbf_kernel.py:7-85)."""
    require_cuda()
    e1, e2 = energy_pack(X1), energy_pack(X2)
    K = empty(e1.n_groups, e2.n_groups)
    dK = empty(e1.n_groups, e2.n_groups) if grad else None
    _lib.call("gprb_kee", _lib.RBF, e1.handle, e2.handle, float(sigma), float(l), float(zeta), 0, e1.n_groups,
              ptr(K), e2.n_groups, ptr(dK), e2.n_groups, stream())
    if grad:
        C = _host(K)
        return C, (2 / sigma) * C, _host(dK)
    return _host(K)


def kef_C(X1, X2, sigma=1.0, l=1.0, zeta=2.0, grad=False, stress=False, transpose=False):
    """Energy-force block [m1, 3 m2] (or its transpose) (rbf_kernel.py:87-189)."""
    require_cuda()
    if stress:
        raise NotImplementedError("stress blocks are not part of the B200 hot path yet (SURVEY.md §8f)")
    e, f = energy_pack(X1), force_pack(X2)
    K = empty(e.n_groups, 3 * f.n_groups)
    dK = empty(e.n_groups, 3 * f.n_groups) if grad else None
    _lib.call("gprb_kef", _lib.RBF, e.handle, f.handle, float(sigma), float(l), float(zeta), 0, f.n_groups,
              ptr(K), 3 * f.n_groups, c_vp(0), 0, ptr(dK), 3 * f.n_groups, c_vp(0), 0, stream())
    C = _host(K)
    if transpose:
        C = C.T
    if grad:
        C_l = _host(dK)
        return C, (2 / sigma) * C, (C_l.T if transpose else C_l)
    return C


def kff_C(X1, X2, sigma=1.0, l=1.0, zeta=2.0, grad=False, stress=False, diag=False, tol=1e-12):
    """Force-force block [3 m1, 3 m2] (rbf_kernel.py:191-337).  The non-grad variant applies the
    reference's `dK_dD > tol` pair cut; the grad variant does not (rbf_kernel.cpp:395 vs :534)."""
    require_cuda()
    if stress:
        raise NotImplementedError("stress blocks are not part of the B200 hot path yet (SURVEY.md §8f)")
    f1, f2 = force_pack(X1), force_pack(X2)
    K = empty(3 * f1.n_groups, 3 * f2.n_groups)
    dK = empty(3 * f1.n_groups, 3 * f2.n_groups) if grad else None
    _lib.call("gprb_kff", _lib.RBF, f1.handle, f2.handle, float(sigma), float(l), float(zeta),
              0 if grad else 1, float(tol), _lib.FF_FULL, 0, f1.n_groups,
              ptr(K), 3 * f2.n_groups, ptr(dK), 3 * f2.n_groups, stream())
    C = _host(K)
    if grad:
        return C, (2 / sigma) * C, _host(dK)
    return C
