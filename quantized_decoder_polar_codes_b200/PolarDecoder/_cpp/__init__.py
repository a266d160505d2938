from quantized_decoder_polar_codes_b200 import _libPolarDecoder  # noqa: F401
