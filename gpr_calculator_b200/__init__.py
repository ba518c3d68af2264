"""gpr_calculator_b200 — B200-native covariance hot path of MaterSim/GPR_calculator.

Public surface mirrors the reference package ``gpr_calc`` for the hot path:
``GP`` (gaussianprocess), ``RBF_mb`` / ``Dot_mb`` (kernels), ``SO3`` (descriptor) and ``GPR``
(ASE calculator adapter, imported lazily because it needs ASE).
"""
__version__ = "0.1.0"
