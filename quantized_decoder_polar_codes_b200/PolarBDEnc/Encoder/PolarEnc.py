from quantized_decoder_polar_codes_b200.encoder import PolarEnc  # noqa: F401
