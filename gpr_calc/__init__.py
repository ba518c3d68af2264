"""Import-path alias of the reference package: ``gpr_calc.*`` resolves to ``gpr_calculator_b200.*``.

Scripts written against MaterSim/GPR_calculator (README.md:34-71: ``from gpr_calc.gaussianprocess import GP``,
``from gpr_calc.calculator import GPR``, ``from gpr_calc.kernels.RBF_mb import RBF_mb``, ``from gpr_calc.SO3 import SO3``)
run unedited with this directory's parent on ``sys.path``.  Every aliased module IS the B200 module (same object), so the
covariance path below it is libgpr_b200.so; there is no second implementation here.  ``gpr_calc.NEB`` (ASE NEB drivers and
plotting, SURVEY.md §2: out of scope) is not provided: keep the reference's ``NEB.py``, it only talks to the calculator API.
"""
import importlib
import sys

_ALIASED = ("gaussianprocess", "calculator", "SO3", "utilities", "kernels", "kernels.base", "kernels.RBF_mb", "kernels.Dot_mb",
            "kernels.rbf_kernel", "kernels.dot_kernel")

for _name in _ALIASED:
    _mod = importlib.import_module("gpr_calculator_b200." + _name)
    sys.modules[__name__ + "." + _name] = _mod
    if "." not in _name:
        globals()[_name] = _mod
del _name, _mod
