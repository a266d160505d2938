from .._cpp.libPolarBD import CASCLDecoder  # noqa: F401  (PolarEncoder/PolarBD/PolarBD/CASCLWithRNTI.py)
