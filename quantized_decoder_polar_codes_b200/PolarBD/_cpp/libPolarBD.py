"""Stands in for the reference's compiled module PolarBD._cpp.libPolarBD (PolarEncoder/PolarBD/_cpp/_libPolarBD.cpp:9-14):
the same two class names, implemented by the CUDA kernels behind include/polar_b200.h (kinds PD_BD_DMETRIC, PD_BD_CASCL)."""
from quantized_decoder_polar_codes_b200._libPolarDecoder import BDCASCLDecoder as CASCLDecoder  # noqa: F401
from quantized_decoder_polar_codes_b200._libPolarDecoder import BDDMetricCalculator as DMetricCalculator  # noqa: F401
