from quantized_decoder_polar_codes_b200._libPolarDecoder import SCLloydQuantizedDecoder  # noqa: F401
