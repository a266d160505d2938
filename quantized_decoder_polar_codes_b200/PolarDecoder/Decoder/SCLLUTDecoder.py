from quantized_decoder_polar_codes_b200._libPolarDecoder import SCLLUTDecoder  # noqa: F401
