"""`from PolarBD.PolarBD.CASCLWithRNTI import CASCLDecoder` on the B200 build (kind PD_BD_CASCL)."""
from quantized_decoder_polar_codes_b200._libPolarDecoder import BDCASCLDecoder as CASCLDecoder  # noqa: F401
