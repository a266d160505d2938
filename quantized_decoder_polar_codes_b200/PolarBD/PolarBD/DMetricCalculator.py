from .._cpp.libPolarBD import DMetricCalculator  # noqa: F401  (PolarEncoder/PolarBD/PolarBD/DMetricCalculator.py)
