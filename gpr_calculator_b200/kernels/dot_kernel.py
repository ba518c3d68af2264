"""Function-level Dot covariance blocks on the B200 (mirror of gpr_calc/kernels/dot_kernel.py)."""
import numpy as np

from .. import _lib
from ..device import energy_pack, force_pack, stress_packs, interleave_stress, empty, ptr, stream, require_cuda, c_vp


def _host(t):
    return t.cpu().numpy()


def _empty_block(p1, p2, r1, r2, n_out):
    """A side without groups (an empty list / tuple) gives an empty block, like the reference's loops."""
    shape = ((0 if p1 is None else p1.n_groups) * r1, (0 if p2 is None else p2.n_groups) * r2)
    return np.zeros(shape) if n_out == 1 else tuple(np.zeros(shape) for _ in range(n_out))


def kee_C(X1, X2, sigma=1.0, sigma0=1.0, zeta=2.0, grad=False):
    """dot_kernel.py:9-63.  d/dsigma0 is the reference's constant 0.8*2*sigma^2*sigma0 (:58)."""
    require_cuda()
    e1, e2 = energy_pack(X1), energy_pack(X2)
    if e1 is None or e2 is None:
        return _empty_block(e1, e2, 1, 1, 3 if grad else 1)
    K = empty(e1.n_groups, e2.n_groups)
    _lib.call("gprb_kee", _lib.DOT, e1.handle, e2.handle, float(sigma), float(sigma0), float(zeta), 0, e1.n_groups,
              ptr(K), e2.n_groups, c_vp(0), 0, stream())
    C = _host(K)
    if grad:
        return C, 2 * C / sigma, 0.8 * 2 * sigma ** 2 * sigma0 * np.ones(C.shape)
    return C


def kef_C(X1, X2, sigma=1.0, sigma0=1.0, zeta=2.0, grad=False, stress=False, transpose=False):
    """dot_kernel.py:66-160 (sigma0 does not enter; d/dsigma0 = 0, :154)."""
    require_cuda()
    if stress:     # dot_kef_many_stress (dot_kernel.cpp:133-216): C [m1, 3 m2], C_s [m1, 6 m2]
        e = energy_pack(X1)
        packs = stress_packs(X2)
        blocks = []
        for f in packs:
            K = empty(e.n_groups, 3 * f.n_groups)
            _lib.call("gprb_kef", _lib.DOT, e.handle, f.handle, float(sigma), float(sigma0), float(zeta), 0, f.n_groups,
                      ptr(K), 3 * f.n_groups, c_vp(0), 0, c_vp(0), 0, c_vp(0), 0, stream())
            blocks.append(K)
        Cs = interleave_stress(blocks[1].T.contiguous(), blocks[2].T.contiguous(), packs[0].n_groups).T
        C, Cs = _host(blocks[0]), _host(Cs.contiguous())
        return (C.T, Cs.T) if transpose else (C, Cs)
    e, f = energy_pack(X1), force_pack(X2)
    if e is None or f is None:
        out = _empty_block(e, f, 1, 3, 3 if grad else 1)
        return (tuple(o.T for o in out) if grad else out.T) if transpose else out
    K = empty(e.n_groups, 3 * f.n_groups)
    _lib.call("gprb_kef", _lib.DOT, e.handle, f.handle, float(sigma), float(sigma0), float(zeta), 0, f.n_groups,
              ptr(K), 3 * f.n_groups, c_vp(0), 0, c_vp(0), 0, c_vp(0), 0, stream())
    C = _host(K)
    if transpose:
        C = C.T
    if grad:
        return C, 2 * C / sigma, np.zeros(C.shape)
    return C


def kff_C(X1, X2, sigma=1.0, sigma0=1.0, zeta=2.0, grad=False, stress=False):
    """dot_kernel.py:162-270 (no pair cut; d/dsigma0 = 0, :265)."""
    require_cuda()
    if stress:     # dot_kff_many_stress (dot_kernel.cpp:338-508): C [3 m1, 3 m2], C_s [6 m1, 3 m2]
        packs = stress_packs(X1)
        f2 = force_pack(X2)
        blocks = []
        for f1 in packs:
            K = empty(3 * f1.n_groups, 3 * f2.n_groups)
            _lib.call("gprb_kff", _lib.DOT, f1.handle, f2.handle, float(sigma), float(sigma0), float(zeta), 0, 0.0,
                      _lib.FF_FULL, 0, f1.n_groups, ptr(K), 3 * f2.n_groups, c_vp(0), 0, stream())
            blocks.append(K)
        return _host(blocks[0]), _host(interleave_stress(blocks[1], blocks[2], packs[0].n_groups))
    f1, f2 = force_pack(X1), force_pack(X2)
    if f1 is None or f2 is None:
        return _empty_block(f1, f2, 3, 3, 3 if grad else 1)
    K = empty(3 * f1.n_groups, 3 * f2.n_groups)
    _lib.call("gprb_kff", _lib.DOT, f1.handle, f2.handle, float(sigma), float(sigma0), float(zeta), 0, 0.0,
              _lib.FF_FULL, 0, f1.n_groups, ptr(K), 3 * f2.n_groups, c_vp(0), 0, stream())
    C = _host(K)
    if grad:
        return C, 2 * C / sigma, np.zeros(C.shape)
    return C
