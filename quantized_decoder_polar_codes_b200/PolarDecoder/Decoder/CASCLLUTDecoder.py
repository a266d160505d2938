from quantized_decoder_polar_codes_b200._libPolarDecoder import CASCLLUTDecoder  # noqa: F401
