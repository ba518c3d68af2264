"""Function-level RBF covariance blocks on the B200 (mirror of gpr_calc/kernels/rbf_kernel.py).

``kee_C / kef_C / kff_C`` keep the reference signatures and return numpy arrays, so parity tests
read like calls into the reference's cffi wrappers; the work is done by libgpr_b200.so.
"""
import numpy as np

from .. import _lib
from ..device import Pack, energy_pack, force_pack, stress_packs, interleave_stress, empty, ptr, stream, require_cuda, c_vp


def _host(t):
    return t.cpu().numpy()


def _empty_block(p1, p2, r1, r2, n_out):
    """A side without groups (an empty list / tuple) gives an empty block, like the reference's loops."""
    if p1 is not None and p2 is not None:
        return None
    shape = ((0 if p1 is None else p1.n_groups) * r1, (0 if p2 is None else p2.n_groups) * r2)
    return np.zeros(shape) if n_out == 1 else tuple(np.zeros(shape) for _ in range(n_out))


def kee_C(X1, X2, sigma=1.0, l=1.0, zeta=2.0, grad=False):
    """Energy-energy block [m1, m2]; grad=True also returns dK/dsigma, dK/dl (rbf_kernel.py:7-85)."""
    require_cuda()
    e1, e2 = energy_pack(X1), energy_pack(X2)
    if e1 is None or e2 is None:
        return _empty_block(e1, e2, 1, 1, 3 if grad else 1)
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
        # 9-column force data: C [m1, 3 m2] and C_s [m1, 6 m2] (rbf_kef_many_stress, rbf_kernel.cpp:256-338)
        e = energy_pack(X1)
        packs = stress_packs(X2)
        blocks = []
        for f in packs:
            K = empty(e.n_groups, 3 * f.n_groups)
            _lib.call("gprb_kef", _lib.RBF, e.handle, f.handle, float(sigma), float(l), float(zeta), 0, f.n_groups,
                      ptr(K), 3 * f.n_groups, c_vp(0), 0, c_vp(0), 0, c_vp(0), 0, stream())
            blocks.append(K)
        m2 = packs[0].n_groups
        Cs = interleave_stress(blocks[1].T.contiguous(), blocks[2].T.contiguous(), m2).T
        C, Cs = _host(blocks[0]), _host(Cs.contiguous())
        return (C.T, Cs.T) if transpose else (C, Cs)
    e, f = energy_pack(X1), force_pack(X2)
    if e is None or f is None:
        out = _empty_block(e, f, 1, 3, 3 if grad else 1)
        return (tuple(o.T for o in out) if grad else out.T) if transpose else out
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
        # X1 carries 9 columns: C [3 m1, 3 m2] and C_s [6 m1, 3 m2] (rbf_kff_many_stress, rbf_kernel.cpp:642-822);
        # the pair cut `dK_dD > tol` applies as in the plain variant
        packs = stress_packs(X1)
        f2 = force_pack(X2)
        blocks = []
        for f1 in packs:
            K = empty(3 * f1.n_groups, 3 * f2.n_groups)
            _lib.call("gprb_kff", _lib.RBF, f1.handle, f2.handle, float(sigma), float(l), float(zeta), 1, float(tol),
                      _lib.FF_FULL, 0, f1.n_groups, ptr(K), 3 * f2.n_groups, c_vp(0), 0, stream())
            blocks.append(K)
        return _host(blocks[0]), _host(interleave_stress(blocks[1], blocks[2], packs[0].n_groups))
    f1, f2 = force_pack(X1), force_pack(X2)
    if f1 is None or f2 is None:
        return _empty_block(f1, f2, 3, 3, 3 if grad else 1)
    K = empty(3 * f1.n_groups, 3 * f2.n_groups)
    dK = empty(3 * f1.n_groups, 3 * f2.n_groups) if grad else None
    _lib.call("gprb_kff", _lib.RBF, f1.handle, f2.handle, float(sigma), float(l), float(zeta),
              0 if grad else 1, float(tol), _lib.FF_FULL, 0, f1.n_groups,
              ptr(K), 3 * f2.n_groups, ptr(dK), 3 * f2.n_groups, stream())
    C = _host(K)
    if grad:
        return C, (2 / sigma) * C, _host(dK)
    return C
