"""B200-native batched polar decoders behind the reference's `PolarDecoder` API.

    from quantized_decoder_polar_codes_b200 import SCLLUTDecoder        # same ctor kwargs as the reference
    dec = SCLLUTDecoder(N=..., K=..., L=8, frozen_bits=..., message_bits=..., LUT_f=..., LUT_g=..., virtual_channel_llr=...)
    bits = dec.decode(symbols)            # (N,) -> (K,)  like the reference;  (B,N) -> (B,K) batched

The 15 classes live in the compiled pybind11 module `_libPolarDecoder` (host C++), which drives the CUDA
kernels through the C ABI of include/polar_b200.h (libpolar_b200.so).  There is no CPU fallback: importing
this package without the built extension raises, and constructing a decoder without a B200 raises.
To use the reference's own import paths (`from PolarDecoder.Decoder.SCLUTDecoder import SCLUTDecoder`) call
`install_reference_import_paths()` or put this directory on sys.path.
"""
import os as _os
import sys as _sys

_HERE = _os.path.dirname(_os.path.abspath(__file__))

# `python -m quantized_decoder_polar_codes_b200.build` imports this package before the extension exists
_BUILDING = (__name__ + ".build") in getattr(_sys, "orig_argv", [])

try:
    if not _BUILDING:
        from . import _libPolarDecoder  # noqa: F401
except ImportError as _e:  # pragma: no cover
    raise ImportError(
        "quantized_decoder_polar_codes_b200: the native extension is not built "
        "(run `python -m quantized_decoder_polar_codes_b200.build`); there is no CPU fallback. "
        f"Original error: {_e}") from _e

if _BUILDING:
    _libPolarDecoder = None

DECODER_CLASSES = (
    "SCDecoder", "FastSCDecoder", "SCLDecoder", "FastSCLDecoder", "CASCLDecoder",
    "SCLUTDecoder", "FastSCLUTDecoder", "SCLLUTDecoder", "FastSCLLUTDecoder", "CASCLLUTDecoder",
    "CAFastSCLLUTDecoder", "SCUniformQuantizedDecoder", "SCLUniformQuantizedDecoder",
    "SCLloydQuantizedDecoder", "SCLLloydQuantizedDecoder",
)
if not _BUILDING:
    for _n in DECODER_CLASSES:
        globals()[_n] = getattr(_libPolarDecoder, _n)
    del _n
    # blind-detection helpers of the reference's PolarEncoder/PolarBD package
    BDDMetricCalculator = _libPolarDecoder.BDDMetricCalculator
    BDCASCLDecoder = _libPolarDecoder.BDCASCLDecoder

LIB_PATH = _os.path.join(_HERE, "libpolar_b200.so")


def install_reference_import_paths():
    """Make `import PolarDecoder.Decoder.<Name>` (the reference's package layout,
    PolarDecoder/PolarDecoder/Decoder/*.py) resolve to this build."""
    if _HERE not in _sys.path:
        _sys.path.insert(0, _HERE)
    import PolarDecoder  # noqa: F401
    return PolarDecoder


__all__ = list(DECODER_CLASSES) + ["BDDMetricCalculator", "BDCASCLDecoder", "install_reference_import_paths", "LIB_PATH"]
