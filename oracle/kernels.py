"""CPU ORACLE (test infrastructure only) — Python drivers for the covariance pair loops.

Two interchangeable native back ends, both driven through ctypes on numpy buffers:

* ``port``  — ``oracle/liboracle_port.so`` built from ``oracle/oracle_kernels.c`` (our C restatement)
* ``ref``   — ``oracle/_ref/lib{rbf,dot}_ref.so``: the UNMODIFIED reference C++
              (gpr_calc/kernels/rbf_kernel.cpp, dot_kernel.cpp) compiled by ``oracle/Makefile``

The functions ``kee_C / kef_C / kff_C`` restate what the reference's cffi wrappers do around
the native call (gpr_calc/kernels/rbf_kernel.py:7-337, dot_kernel.py:9-270): expansion of
``indices`` into per-row group ids, output reshapes, the 1/n_I (1/n_J) normalisations and the
derived ``dK/dsigma = 2K/sigma``.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_D, _I = ctypes.c_double, ctypes.c_int
_P = ctypes.c_void_p

_LIBS = {}


def build(ref=True, port=True, quiet=True):
    """Compile the oracle libraries (building the checker is not using it)."""
    targets = []
    if port:
        targets.append("port")
    if ref:
        targets.append("ref")
    out = subprocess.run(["make", "-s", "-C", HERE] + targets, capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def have_ref():
    return os.path.exists(os.path.join(HERE, "_ref", "librbf_ref.so"))


def _lib(name):
    if name not in _LIBS:
        path = {"port": os.path.join(HERE, "liboracle_port.so"),
                "rbf_ref": os.path.join(HERE, "_ref", "librbf_ref.so"),
                "dot_ref": os.path.join(HERE, "_ref", "libdot_ref.so")}[name]
        if not os.path.exists(path):
            build(ref=name != "port", port=name == "port")
        lib = ctypes.CDLL(path)
        if name == "port":
            lib.orc_rbf_kee.argtypes = [_I, _I, _I, _I, _D, _D, _D] + [_P] * 8
            lib.orc_rbf_kef.argtypes = [_I, _I, _I, _I, _I, _D, _D, _D] + [_P] * 9
            lib.orc_rbf_kff.argtypes = [_I, _I, _I, _I, _I, _I, _I, _D, _D, _D, _I, _D] + [_P] * 10
            lib.orc_dot_kee.argtypes = [_I, _I, _I, _I, _D, _D, _D] + [_P] * 7
            lib.orc_dot_kef.argtypes = [_I, _I, _I, _I, _I, _D] + [_P] * 8
            lib.orc_dot_kff.argtypes = [_I, _I, _I, _I, _I, _I, _I, _D] + [_P] * 9
        elif name == "rbf_ref":      # rbf_kernel.h:4-38
            lib.rbf_kee_many.argtypes = [_I, _I, _I, _I, _D, _D, _D] + [_P] * 7
            lib.rbf_kee_many_with_grad.argtypes = [_I, _I, _I, _I, _D, _D, _D] + [_P] * 8
            lib.rbf_kef_many.argtypes = [_I, _I, _I, _I, _D, _D, _D] + [_P] * 8
            lib.rbf_kef_many_with_grad.argtypes = [_I, _I, _I, _I, _D, _D, _D] + [_P] * 8
            lib.rbf_kef_many_stress.argtypes = [_I, _I, _I, _I, _D, _D, _D] + [_P] * 8
            lib.rbf_kff_many.argtypes = [_I, _I, _I, _I, _I, _I, _D, _D, _D, _D] + [_P] * 9
            lib.rbf_kff_many_with_grad.argtypes = [_I, _I, _I, _I, _I, _I, _D, _D, _D] + [_P] * 10
            lib.rbf_kff_many_stress.argtypes = [_I, _I, _I, _I, _I, _I, _D, _D, _D, _D] + [_P] * 9
        else:                        # dot_kernel.h:4-27
            lib.dot_kee_many.argtypes = [_I, _I, _I, _I, _D, _D, _D] + [_P] * 7
            lib.dot_kef_many.argtypes = [_I, _I, _I, _I, _D] + [_P] * 8
            lib.dot_kef_many_stress.argtypes = [_I, _I, _I, _I, _D] + [_P] * 8
            lib.dot_kff_many.argtypes = [_I, _I, _I, _I, _I, _I, _D] + [_P] * 9
            lib.dot_kff_many_stress.argtypes = [_I, _I, _I, _I, _I, _I, _D] + [_P] * 9
        _LIBS[name] = lib
    return _LIBS[name]


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_P)


# ----------------------------------------------------------------------------------------------
# packed <-> ragged layout (gpr_calc/utilities.py:340-406)
# ----------------------------------------------------------------------------------------------
def list_to_tuple(data, stress=False, include_value=False, mode="force"):
    rows = sum(fd[0].shape[0] for fd in data)
    d = data[-1][0].shape[1]
    X = np.zeros([rows, d])
    ELE, indices, values = [], [], []
    if mode == "force":
        dXdR = np.zeros([rows, d, 9 if stress else 3])
    c = 0
    for fd in data:
        if mode == "force":
            if include_value:
                x, dxdr, f, ele = fd
                values.append(f)
            else:
                x, dxdr, ele = fd
            dXdR[c:c + x.shape[0]] = dxdr
        else:
            if include_value:
                x, e, ele = fd
                values.append(e)
            else:
                x, ele = fd
        n = x.shape[0]
        indices.append(n)
        X[c:c + n] = x
        ELE.extend(ele)
        c += n
    ELE = np.ravel(ELE)
    if mode == "force":
        return (X, dXdR, ELE, indices, values) if include_value else (X, dXdR, ELE, indices)
    return (X, ELE, indices, values) if include_value else (X, ELE, indices)


def tuple_to_list(data, mode="force"):
    out, c = [], 0
    if mode == "force":
        X, dXdR, ELE, indices = data
        for n in indices:
            out.append((X[c:c + n], dXdR[c:c + n], ELE[c:c + n]))
            c += n
    else:
        X, ELE, indices = data
        for n in indices:
            out.append((X[c:c + n], ELE[c:c + n]))
            c += n
    return out


def _group_ids(indices):
    return np.repeat(np.arange(len(indices), dtype=np.int32), np.asarray(indices, dtype=np.int64)).astype(np.int32)


def _as_energy(X):
    if isinstance(X, list):
        X = list_to_tuple(X, mode="energy")
    x, ele, ind = X
    return _f64(x), _i32(ele), list(ind)


def _as_force(X, stress=False):
    if isinstance(X, (list, np.ndarray)):
        X = list_to_tuple(list(X), stress=stress)
    x, dxdr, ele, ind = X
    return _f64(x), _f64(dxdr), _i32(ele), list(ind)


# ----------------------------------------------------------------------------------------------
# RBF wrappers  (gpr_calc/kernels/rbf_kernel.py)
# ----------------------------------------------------------------------------------------------
class RBFOracle:
    """kee_C / kef_C / kff_C of the RBF kernel on a chosen native back end."""

    def __init__(self, backend="port"):
        assert backend in ("port", "ref")
        self.backend = backend

    # rbf_kernel.py:7-85
    def kee_C(self, X1, X2, sigma=1.0, l=1.0, zeta=2.0, grad=False):
        x1, e1, i1 = _as_energy(X1)
        x2, e2, i2 = _as_energy(X2)
        g1, g2 = _group_ids(i1), _group_ids(i2)
        m1, m2, d = len(i1), len(i2), x1.shape[1]
        out = np.zeros((m1, m2))
        dout = np.zeros((m1, m2)) if grad else None
        zeta, s2, l2 = float(zeta), float(sigma * sigma), float(l * l)
        if self.backend == "port":
            _lib("port").orc_rbf_kee(len(x1), len(x2), d, m2, zeta, s2, l2, _ptr(x1), _ptr(e1), _ptr(g1),
                                     _ptr(x2), _ptr(e2), _ptr(g2), _ptr(out), _ptr(dout))
        elif grad:
            _lib("rbf_ref").rbf_kee_many_with_grad(len(x1), len(x2), d, m2, zeta, s2, l2, _ptr(x1), _ptr(e1),
                                                   _ptr(g1), _ptr(x2), _ptr(e2), _ptr(g2), _ptr(out), _ptr(dout))
        else:
            _lib("rbf_ref").rbf_kee_many(len(x1), len(x2), d, m2, zeta, s2, l2, _ptr(x1), _ptr(e1), _ptr(g1),
                                         _ptr(x2), _ptr(e2), _ptr(g2), _ptr(out))
        nn = np.array(i1)[:, None] * np.array(i2)[None, :]
        C = out / nn
        if grad:
            C_l = dout / nn * (1.0 / (l * l2))
            return C, (2 / sigma) * C, C_l
        return C

    # rbf_kernel.py:87-189
    def kef_C(self, X1, X2, sigma=1.0, l=1.0, zeta=2.0, grad=False, stress=False, transpose=False):
        x1, e1, i1 = _as_energy(X1)
        x2, dx2, e2, i2 = _as_force(X2, stress=stress)
        g1, g2 = _group_ids(i1), _group_ids(i2)
        m1, m2, d = len(i1), len(i2), x1.shape[1]
        nc = 9 if stress else 3
        zeta, s2, l2 = float(zeta), float(sigma * sigma), float(l * l)
        if self.backend == "port":
            out = np.zeros((m1, m2, nc))
            dout = np.zeros((m1, m2, nc)) if grad else None
            _lib("port").orc_rbf_kef(len(x1), len(x2), d, m2, nc, zeta, s2, float(l), _ptr(x1), _ptr(e1), _ptr(g1),
                                     _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out), _ptr(dout))
        elif stress:
            out = np.zeros((m1, m2, 9))
            _lib("rbf_ref").rbf_kef_many_stress(len(x1), len(x2), d, m2, zeta, s2, l2, _ptr(x1), _ptr(e1), _ptr(g1),
                                                _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out))
        elif grad:
            buf = np.zeros((m1, m2, 6))
            _lib("rbf_ref").rbf_kef_many_with_grad(len(x1), len(x2), d, m2, zeta, s2, float(l), _ptr(x1), _ptr(e1),
                                                   _ptr(g1), _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(buf))
            out, dout = buf[:, :, :3].copy(), buf[:, :, 3:].copy()
        else:
            out = np.zeros((m1, m2, 3))
            _lib("rbf_ref").rbf_kef_many(len(x1), len(x2), d, m2, zeta, s2, l2, _ptr(x1), _ptr(e1), _ptr(g1),
                                         _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out))
        nI = np.array(i1)[:, None, None]
        out = out / nI
        C = out[:, :, :3].reshape(m1, m2 * 3)
        if stress:
            Cs = out[:, :, 3:].reshape(m1, m2 * 6)
        elif grad:
            C_l = (dout / nI).reshape(m1, m2 * 3)
            C_s = (2 / sigma) * C
        else:
            Cs = np.zeros((m1, m2 * 6))
        if transpose:
            C = C.T
            if not grad:
                Cs = Cs.T
        if grad:
            return C, C_s, C_l
        if stress:
            return C, Cs
        return C

    # rbf_kernel.py:191-337
    def kff_C(self, X1, X2, sigma=1.0, l=1.0, zeta=2.0, grad=False, stress=False, tol=1e-12,
              n2_window=None):
        x1, dx1, e1, i1 = _as_force(X1, stress=stress)
        x2, dx2, e2, i2 = _as_force(X2, stress=False) if not stress else _as_force(X2, stress=stress)
        if stress:
            dx2 = _f64(dx2[:, :, :3])
        g1, g2 = _group_ids(i1), _group_ids(i2)
        m1, m2, d = len(i1), len(i2), x1.shape[1]
        nc1 = 9 if stress else 3
        zeta, s2, l2 = float(zeta), float(sigma * sigma), float(l * l)
        lo, hi = (0, len(x2)) if n2_window is None else n2_window
        out = np.zeros((m1, nc1, m2 * 3))
        dout = np.zeros((m1, nc1, m2 * 3)) if grad else None
        if self.backend == "port":
            _lib("port").orc_rbf_kff(len(x1), len(x2), lo, hi, d, m2, nc1, zeta, s2, float(l),
                                     0 if grad else 1, float(tol),
                                     _ptr(x1), _ptr(dx1), _ptr(e1), _ptr(g1),
                                     _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out), _ptr(dout))
        elif stress:
            _lib("rbf_ref").rbf_kff_many_stress(len(x1), len(x2), lo, hi, d, m2, zeta, s2, l2, float(tol),
                                                _ptr(x1), _ptr(dx1), _ptr(e1), _ptr(g1),
                                                _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out))
        elif grad:
            _lib("rbf_ref").rbf_kff_many_with_grad(len(x1), len(x2), lo, hi, d, m2, zeta, s2, float(l),
                                                   _ptr(x1), _ptr(dx1), _ptr(e1), _ptr(g1),
                                                   _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out), _ptr(dout))
        else:
            _lib("rbf_ref").rbf_kff_many(len(x1), len(x2), lo, hi, d, m2, zeta, s2, l2, float(tol),
                                         _ptr(x1), _ptr(dx1), _ptr(e1), _ptr(g1),
                                         _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out))
        C = out[:, :3, :].reshape(m1 * 3, m2 * 3)
        if grad:
            return C, (2 / sigma) * C, dout.reshape(m1 * 3, m2 * 3)
        if stress:
            return C, out[:, 3:, :].reshape(m1 * 6, m2 * 3)
        return C


# ----------------------------------------------------------------------------------------------
# Dot wrappers  (gpr_calc/kernels/dot_kernel.py)
# ----------------------------------------------------------------------------------------------
class DotOracle:
    def __init__(self, backend="port"):
        assert backend in ("port", "ref")
        self.backend = backend

    # dot_kernel.py:9-63
    def kee_C(self, X1, X2, sigma=1.0, sigma0=1.0, zeta=2.0, grad=False):
        x1, e1, i1 = _as_energy(X1)
        x2, e2, i2 = _as_energy(X2)
        g1, g2 = _group_ids(i1), _group_ids(i2)
        m1, m2, d = len(i1), len(i2), x1.shape[1]
        out = np.zeros((m1, m2))
        args = (len(x1), len(x2), d, m2, float(zeta), float(sigma ** 2), float(sigma0 ** 2), _ptr(x1), _ptr(e1),
                _ptr(g1), _ptr(x2), _ptr(e2), _ptr(g2), _ptr(out))
        if self.backend == "port":
            _lib("port").orc_dot_kee(*args)
        else:
            _lib("dot_ref").dot_kee_many(*args)
        C = out / (np.array(i1)[:, None] * np.array(i2)[None, :])
        if grad:
            return C, 2 * C / sigma, 0.8 * 2 * sigma ** 2 * sigma0 * np.ones([m1, m2])
        return C

    # dot_kernel.py:66-160
    def kef_C(self, X1, X2, sigma=1.0, sigma0=1.0, zeta=2.0, grad=False, stress=False, transpose=False):
        x1, e1, i1 = _as_energy(X1)
        x2, dx2, e2, i2 = _as_force(X2, stress=stress)
        g1, g2 = _group_ids(i1), _group_ids(i2)
        m1, m2, d = len(i1), len(i2), x1.shape[1]
        nc = 9 if stress else 3
        out = np.zeros((m1, m2, nc))
        if self.backend == "port":
            _lib("port").orc_dot_kef(len(x1), len(x2), d, m2, nc, float(zeta), _ptr(x1), _ptr(e1), _ptr(g1),
                                     _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out))
        else:
            fn = _lib("dot_ref").dot_kef_many_stress if stress else _lib("dot_ref").dot_kef_many
            fn(len(x1), len(x2), d, m2, float(zeta), _ptr(x1), _ptr(e1), _ptr(g1),
               _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out))
        out = out / np.array(i1)[:, None, None] * (-sigma * sigma)
        C = out[:, :, :3].reshape(m1, m2 * 3)
        Cs = out[:, :, 3:].reshape(m1, m2 * 6) if stress else np.zeros((m1, m2 * 6))
        if transpose:
            C, Cs = C.T, Cs.T
        if grad:
            return C, 2 * C / sigma, np.zeros([m1, m2 * 3])
        if stress:
            return C, Cs
        return C

    # dot_kernel.py:162-270 (single rank: column window = all rows)
    def kff_C(self, X1, X2, sigma=1.0, sigma0=1.0, zeta=2.0, grad=False, stress=False, n2_window=None):
        x1, dx1, e1, i1 = _as_force(X1, stress=stress)
        x2, dx2, e2, i2 = _as_force(X2, stress=False) if not stress else _as_force(X2, stress=stress)
        if stress:
            dx2 = _f64(dx2[:, :, :3])
        g1, g2 = _group_ids(i1), _group_ids(i2)
        m1, m2, d = len(i1), len(i2), x1.shape[1]
        nc1 = 9 if stress else 3
        lo, hi = (0, len(x2)) if n2_window is None else n2_window
        out = np.zeros((m1, nc1, m2 * 3))
        if self.backend == "port":
            _lib("port").orc_dot_kff(len(x1), len(x2), lo, hi, d, m2, nc1, float(zeta),
                                     _ptr(x1), _ptr(dx1), _ptr(e1), _ptr(g1),
                                     _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out))
        else:
            fn = _lib("dot_ref").dot_kff_many_stress if stress else _lib("dot_ref").dot_kff_many
            fn(len(x1), len(x2), lo, hi, d, m2, float(zeta), _ptr(x1), _ptr(dx1), _ptr(e1), _ptr(g1),
               _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out))
        out *= sigma * sigma * zeta
        C = out[:, :3, :].reshape(m1 * 3, m2 * 3)
        if grad:
            return C, 2 * C / sigma, np.zeros([m1 * 3, m2 * 3])
        if stress:
            return C, out[:, 3:, :].reshape(m1 * 6, m2 * 3)
        return C
