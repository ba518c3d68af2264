"""CPU: the C-ABI library loads and exports every symbol include/gpr_b200.h declares
(no compute calls: there is no GPU here)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "gpr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gprb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    names = _declared()
    for must in ("gprb_pack_create", "gprb_kff", "gprb_kef", "gprb_kee", "gprb_chol_factor", "gprb_predict",
                 "gprb_so3_neighbors", "gprb_so3_radial", "gprb_so3_power"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from gpr_calculator_b200 import _lib
    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), "libgpr_b200.so does not export %s" % name
        assert name in _lib.SIGNATURES, "no ctypes signature for %s" % name
    assert set(_lib.SIGNATURES) == set(_declared())
    assert lib.gprb_version() >= 100


def test_no_cuda_means_loud_failure():
    import torch
    import pytest
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from gpr_calculator_b200.device import require_cuda
    from gpr_calculator_b200.kernels import RBF_mb
    with pytest.raises(RuntimeError):
        require_cuda()
    import numpy as np
    x = np.ones((2, 30))
    with pytest.raises(RuntimeError):
        RBF_mb().k_total({"energy": (x, np.array([1, 1]), [2])})


def test_product_never_imports_oracle():
    """The product package must not reference the oracle (a routed-through oracle voids parity)."""
    pkg = os.path.join(ROOT, "gpr_calculator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text, f
