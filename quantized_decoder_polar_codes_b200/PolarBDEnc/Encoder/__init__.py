"""Import-path shim for the reference drivers' `from PolarBDEnc.Encoder.PolarEnc import PolarEnc` /
`from PolarBDEnc.Encoder.CRCEnc import CRCEnc` (mainFPDecoder.py:12-13); see quantized_decoder_polar_codes_b200/encoder.py."""
