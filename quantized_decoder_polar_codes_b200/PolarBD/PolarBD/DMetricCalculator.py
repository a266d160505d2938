"""`from PolarBD.PolarBD.DMetricCalculator import DMetricCalculator` on the B200 build (kind PD_BD_DMETRIC)."""
from quantized_decoder_polar_codes_b200._libPolarDecoder import BDDMetricCalculator as DMetricCalculator  # noqa: F401
