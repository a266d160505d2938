from quantized_decoder_polar_codes_b200.encoder import CRCEnc  # noqa: F401
