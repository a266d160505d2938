"""Generates tests/golden/real_lut_n128.npz: REAL MinDistortion lookup tables (BASELINE.json configs 2/3:
N=128, QDecoder=QChannel=16, DesignSNR=3.0 dB) produced by the REFERENCE's own generator code
(QuantizeDensityEvolution/QLLRDensityEvolution_MinDistortion.py: LLRQuantizerSC.run, driven exactly like
GenerateLookUpTable_LLRDomain.py:33-55), the channel quantizers the driver builds per Eb/N0
(mainQuantizedDecoder_LLRDomain.py:130-145), seeded AWGN frames quantized with the driver's rule (:167-176),
and the outputs of the compiled reference decoders on them.

The reference's C++ `quantizers` package needs OpenCV (absent here), so `LLRQuantizer.find_OptLS_quantizer` is
served by the reference's own pure-numpy restatement QuantizeDensityEvolution/MinDistortionQuantizer.py.
Run in the build container only:  python tests/golden/make_real_luts.py
"""
import os
import sys
import types
from bisect import bisect_left

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, REF)

# ---- shim: quantizers.quantizer.LLROptLSQuantizer.LLRQuantizer over the reference's numpy DP ----
from QuantizeDensityEvolution import MinDistortionQuantizer as _mdq  # noqa: E402


class LLRQuantizer:
    def find_OptLS_quantizer(self, density, quanta, M, K):
        density = np.asarray(density, dtype=np.float64).ravel()
        quanta = np.asarray(quanta, dtype=np.float64).ravel()
        if density.shape[0] <= K:   # nothing to compress: identity (sorted) mapping padded to K symbols
            order = np.argsort(quanta)
            lut = np.zeros(density.shape[0], np.int32)
            lut[order] = np.arange(density.shape[0])
            d = np.zeros(K); q = np.zeros(K)
            d[: density.shape[0]] = density[order]; q[: density.shape[0]] = quanta[order]
            return d[None], q[None], lut[None], 0.0
        d, q, lut = _mdq.find_OptLS_quantizer(density, quanta, K)
        return d[None], q[None], lut[None], 0.0


for name in ["quantizers", "quantizers.quantizer", "quantizers.quantizer.LLROptLSQuantizer"]:
    sys.modules[name] = types.ModuleType(name)
sys.modules["quantizers.quantizer.LLROptLSQuantizer"].LLRQuantizer = LLRQuantizer

from QuantizeDensityEvolution.QLLRDensityEvolution_MinDistortion import LLRQuantizerSC  # noqa: E402
from utils import channel_llr_density_table  # noqa: E402  (the reference's utils.py)

from oracle import polar_oracle as po  # noqa: E402
import common  # noqa: E402
from quantized_decoder_polar_codes_b200 import simulation as sim  # noqa: E402

N, QD, QC, QCU, DESIGN_DB = 128, 16, 16, 128, 3.0


def channel_quantizer(sigma):
    E = 2 / sigma ** 2
    D = np.sqrt(2 * E)
    pyx, interval_x, quanta = channel_llr_density_table(QCU, -E - 3 * D, E + 3 * D, E, -E, D)
    dens, q, lut, _ = LLRQuantizer().find_OptLS_quantizer(pyx, quanta, QCU, QC)
    return dens.squeeze(), q.squeeze(), lut.squeeze(), interval_x


def quantize(llr, interval_x, channel_lut):
    out = np.zeros(llr.shape, np.int32)
    flat, o = llr.ravel(), out.ravel()
    xs = list(interval_x[:-1])
    for i, v in enumerate(flat):
        if v <= interval_x[0]:
            o[i] = 0
        elif v >= interval_x[-1]:
            o[i] = QC - 1
        else:
            o[i] = channel_lut[bisect_left(xs, v) - 1]
    return out


def main():
    cache = "/tmp/real_lut_cache.npz"   # the density evolution takes ~4 minutes in pure Python
    if os.path.exists(cache):
        c = np.load(cache)
        f, g, llr_quanta = c["f"], c["g"], c["q"]
    else:
        sigma_d = np.sqrt(1 / 10 ** (DESIGN_DB / 10))
        dens, quanta, _, _ = channel_quantizer(sigma_d)
        llr_density, llr_quanta, lut_fs, lut_gs = LLRQuantizerSC(N, QD).run(channel_llr_density=dens, channel_llr_quanta=quanta)
        f = np.stack([np.asarray(lut_fs[p][0], np.uint8) for p in range(N - 1)])
        g = np.stack([np.asarray(lut_gs[p][0], np.uint8) for p in range(N - 1)])
        np.savez(cache, f=f, g=g, q=np.asarray(llr_quanta))
    out = {"lut_f": f, "lut_g": g, "llr_quanta": np.asarray(llr_quanta, np.float64)}
    ref = po.load_reference()
    rng = np.random.default_rng(2024)
    LUT_f = [np.broadcast_to(f[p].astype(np.int32), (N >> (int(np.log2(p + 1)) + 1), QD, QD)) for p in range(N - 1)]
    LUT_g = [np.broadcast_to(g[p].astype(np.int32), (N >> (int(np.log2(p + 1)) + 1), 2, QD, QD)) for p in range(N - 1)]
    for A, crc in [(32, False), (32, True)]:
        K = A + 24 if crc else A
        fm, mm = sim.frozen_mask(N, K)
        nt = sim.identify_nodes(N, fm)
        kinds = ["CASCLLUTDecoder", "CAFastSCLLUTDecoder"] if crc else ["SCLUTDecoder", "FastSCLUTDecoder", "SCLLUTDecoder", "FastSCLLUTDecoder"]
        for eb in [1.0, 2.0, 3.0]:
            sigma = sim.awgn_sigma(eb, A / N)
            _, _, clut, ix = channel_quantizer(sigma)
            B = 400
            msg = rng.integers(0, 2, (B, A), dtype=np.uint8)
            word = sim.crc_attach(msg) if crc else msg
            x = quantize(sim.awgn_llr(sim.polar_encode(word, fm), sigma, rng), ix, clut)
            tag = f"A{A}{'crc' if crc else ''}_eb{eb:.0f}"
            out[tag + "/x"] = x.astype(np.uint8)
            out[tag + "/chan_edges"] = np.asarray(ix, np.float64)
            out[tag + "/chan_lut"] = np.asarray(clut, np.uint8)
            out[tag + "/msg"] = msg
            for kind in kinds:
                kw = dict(N=N, K=K, frozen_bits=fm, message_bits=mm, virtual_channel_llr=llr_quanta)
                if kind in common.LIST_KINDS:
                    kw["L"] = 8
                if kind.startswith("CA"):
                    kw["A"] = A
                if kind == "CASCLLUTDecoder":
                    kw.update(crc_n=24, crc_p=list(sim.CRC24_LOC))
                if "Fast" in kind:
                    kw["node_type"] = nt
                nf, ng = ("LUT_Fs", "LUT_Gs") if kind == "FastSCLUTDecoder" else ("LUT_f", "LUT_g")
                kw[nf], kw[ng] = LUT_f, LUT_g
                y = common.ref_decode(ref, kind, kw, x)
                out[f"{tag}/{kind}"] = np.packbits(y, axis=1)
                bler = (y != msg).any(axis=1).mean()
                print(f"{tag:12s} {kind:22s} BLER {bler:.3f}")
    np.savez_compressed(os.path.join(HERE, "real_lut_n128.npz"), **out)
    sym = all((g[p, 1] == g[p, 0][::-1]).all() for p in range(N - 1))
    print("g tables mirror-symmetric (g1[a][b] == g0[Q-1-a][b]) at every node:", sym)


if __name__ == "__main__":
    main()
