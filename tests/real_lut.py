"""Loader for tests/golden/real_lut_n128.npz (real MinDistortion tables from the reference's generator code)."""
import os

import numpy as np

import common
from quantized_decoder_polar_codes_b200 import simulation as sim

N, QD = 128, 16
_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "real_lut_n128.npz")
KINDS_PLAIN = ["SCLUTDecoder", "FastSCLUTDecoder", "SCLLUTDecoder", "FastSCLLUTDecoder"]
KINDS_CRC = ["CASCLLUTDecoder", "CAFastSCLLUTDecoder"]
CASES = [(f"A32_eb{e}", k) for e in (1, 2, 3) for k in KINDS_PLAIN] + [(f"A32crc_eb{e}", k) for e in (1, 2, 3) for k in KINDS_CRC]


def load():
    return np.load(_PATH)


def build_kwargs(z, tag, kind):
    crc = "crc" in tag
    A = 32
    K = A + 24 if crc else A
    fm, mm = sim.frozen_mask(N, K)
    f = [z["lut_f"][p].astype(np.int32)[None] for p in range(N - 1)]
    g = [z["lut_g"][p].astype(np.int32)[None] for p in range(N - 1)]
    kw = dict(N=N, K=K, frozen_bits=fm, message_bits=mm, virtual_channel_llr=z["llr_quanta"])
    if kind in common.LIST_KINDS:
        kw["L"] = 8
    if kind.startswith("CA"):
        kw["A"] = A
    if kind == "CASCLLUTDecoder":
        kw.update(crc_n=24, crc_p=list(sim.CRC24_LOC))
    if "Fast" in kind:
        kw["node_type"] = sim.identify_nodes(N, fm)
    nf, ng = ("LUT_Fs", "LUT_Gs") if kind == "FastSCLUTDecoder" else ("LUT_f", "LUT_g")
    kw[nf], kw[ng] = f, g
    x = z[tag + "/x"]
    want = np.unpackbits(z[f"{tag}/{kind}"], axis=1)[:, :A if crc else K]
    return kw, x, want, z[tag + "/msg"]
