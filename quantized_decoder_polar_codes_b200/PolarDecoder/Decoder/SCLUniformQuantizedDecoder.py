from quantized_decoder_polar_codes_b200._libPolarDecoder import SCLUniformQuantizedDecoder  # noqa: F401
