"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares; the pybind11 mirror exposes the reference's 15 classes with the reference's keyword names; host
validation errors are raised without touching a GPU.  (No compute calls here: there is no CPU decode path.)"""
import ctypes
import os
import re

import numpy as np
import pytest

import common

ROOT = common.ROOT


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "polar_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pd_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from quantized_decoder_polar_codes_b200 import capi
    lib = ctypes.CDLL(capi.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"libpolar_b200.so does not export {s}"
    assert "sm_100a" in capi.lib().pd_version().decode()


# keyword names of the reference constructors (PolarDecoder/PolarDecoder/_cpp/py_interface/py_*.cpp)
REF_SIGNATURES = {
    "SCDecoder": ["N", "K", "frozen_bits", "message_bits"],
    "FastSCDecoder": ["N", "K", "frozen_bits", "message_bits", "node_type"],
    "SCLDecoder": ["N", "K", "L", "frozen_bits", "message_bits"],
    "FastSCLDecoder": ["N", "K", "L", "frozen_bits", "message_bits", "node_type"],
    "CASCLDecoder": ["N", "K", "A", "L", "frozen_bits", "message_bits", "crc_n", "crc_p"],
    "SCLUTDecoder": ["N", "K", "frozen_bits", "message_bits", "LUT_f", "LUT_g", "virtual_channel_llr"],
    "FastSCLUTDecoder": ["N", "K", "frozen_bits", "message_bits", "node_type", "LUT_Fs", "LUT_Gs", "virtual_channel_llr"],
    "SCLLUTDecoder": ["N", "K", "L", "frozen_bits", "message_bits", "LUT_f", "LUT_g", "virtual_channel_llr"],
    "FastSCLLUTDecoder": ["N", "K", "L", "frozen_bits", "message_bits", "node_type", "LUT_f", "LUT_g", "virtual_channel_llr"],
    "CASCLLUTDecoder": ["N", "K", "A", "L", "frozen_bits", "message_bits", "crc_n", "crc_p", "LUT_f", "LUT_g", "virtual_channel_llr"],
    "CAFastSCLLUTDecoder": ["N", "K", "A", "L", "frozen_bits", "message_bits", "node_type", "LUT_f", "LUT_g", "virtual_channel_llr"],
    "SCUniformQuantizedDecoder": ["N", "K", "frozen_bits", "message_bits", "decoder_r_f", "decoder_r_g", "v"],
    "SCLUniformQuantizedDecoder": ["N", "K", "L", "frozen_bits", "message_bits", "decoder_r_f", "decoder_r_g", "v"],
    "SCLloydQuantizedDecoder": ["N", "K", "frozen_bits", "message_bits", "boundaries_f", "boundaries_g", "reconstruction_f", "reconstruction_g", "v"],
    "SCLLloydQuantizedDecoder": ["N", "K", "L", "frozen_bits", "message_bits", "boundaries_f", "boundaries_g", "reconstruction_f", "reconstruction_g", "v"],
}
DECODE_ARG = {k: ("channel_quantized_symbols" if "LUT" in k and k != "FastSCLUTDecoder" else "llr") for k in REF_SIGNATURES}


@pytest.mark.parametrize("name", sorted(REF_SIGNATURES))
def test_pybind_surface_matches_reference(name):
    import quantized_decoder_polar_codes_b200 as q
    cls = getattr(q, name)
    doc = cls.__init__.__doc__
    got = re.findall(r"(\w+): ", doc.split("->")[0])
    got = [g for g in got if g != "self"]
    assert got[: len(REF_SIGNATURES[name])] == REF_SIGNATURES[name], doc
    assert got[len(REF_SIGNATURES[name]):] == ["device"]
    assert DECODE_ARG[name] + ":" in cls.decode.__doc__


def test_reference_signatures_agree_with_compiled_reference(refmod):
    for name, kws in REF_SIGNATURES.items():
        doc = getattr(refmod, name).__init__.__doc__
        got = [g for g in re.findall(r"(\w+): ", doc.split("->")[0]) if g != "self"]
        assert got == kws, name
        assert DECODE_ARG[name] + ":" in getattr(refmod, name).decode.__doc__


def test_reference_import_paths():
    import quantized_decoder_polar_codes_b200 as q
    q.install_reference_import_paths()
    from PolarDecoder.Decoder.SCLLUTDecoder import SCLLUTDecoder
    from PolarDecoder.Decoder.CAFastSCLLUTDecoder import CAFastSCLLUTDecoder  # noqa: F401
    assert SCLLUTDecoder is q.SCLLUTDecoder


def test_host_validation_errors_without_gpu():
    import quantized_decoder_polar_codes_b200 as q
    fm, mm = common.sim.frozen_mask(16, 8)
    with pytest.raises(ValueError, match="power of two"):
        q.SCDecoder(N=12, K=6, frozen_bits=fm[:12], message_bits=mm[:12])
    with pytest.raises(ValueError, match="non-frozen"):
        q.SCDecoder(N=16, K=7, frozen_bits=fm, message_bits=mm)
    with pytest.raises(ValueError, match="L="):
        q.SCLDecoder(N=16, K=8, L=64, frozen_bits=fm, message_bits=mm)
    nt = -np.ones(31, np.int32)
    nt[0] = 1
    with pytest.raises(ValueError, match="degenerate"):
        q.FastSCDecoder(N=16, K=8, frozen_bits=fm, message_bits=mm, node_type=nt)
    with pytest.raises(ValueError, match="crc_n"):
        q.CASCLDecoder(N=16, K=8, A=4, L=2, frozen_bits=fm, message_bits=mm, crc_n=40, crc_p=[0, 40])


def test_no_cpu_fallback_in_product_sources():
    """The product package must not import or link the oracle."""
    pkg = os.path.join(ROOT, "quantized_decoder_polar_codes_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "polar_oracle" not in txt and "oracle/" not in txt.replace("oracle/_ref", ""), f


def _build_c_example(tmp_path):
    import subprocess
    exe = tmp_path / "sc_roundtrip"
    pkg = os.path.join(ROOT, "quantized_decoder_polar_codes_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "sc_roundtrip.c"), "-L" + pkg, "-lpolar_b200", "-Wl,-rpath," + pkg, "-o", str(exe)])
    return exe


def test_plain_c_caller_builds_and_links(tmp_path):
    """include/polar_b200.h is a C header and libpolar_b200.so a C-ABI library: a C99 program links against it."""
    import subprocess
    exe = _build_c_example(tmp_path)
    assert subprocess.check_output([str(exe), "--version"]).decode().startswith("polar_b200")


@pytest.mark.gpu
def test_plain_c_caller_round_trip(tmp_path):
    """examples/sc_roundtrip.c: encode (CRC + polar) -> noiseless channel -> SC and CA-SCL decode, all through the C ABI."""
    import subprocess
    out = subprocess.check_output([str(_build_c_example(tmp_path))]).decode()
    assert out.startswith("ok: 1000 frames")
