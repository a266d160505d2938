"""Generates tests/golden/lutgen_golden.npz: inputs and outputs of the REFERENCE's LLR-domain table generator
(QuantizeDensityEvolution/QLLRDensityEvolution_MinDistortion.py: LLRQuantizerSC.run, driven like
GenerateLookUpTable_LLRDomain.py:33-55) for small codes, plus stand-alone quantizer problems, as golden vectors for
quantized_decoder_polar_codes_b200/lutgen.py and pd_optls_quantize.

The generator's `LLRQuantizer.find_OptLS_quantizer` is C++ on OpenCV (cannot be built here); it is served by the
reference's own numpy restatement QuantizeDensityEvolution/MinDistortionQuantizer.py (same shim as make_real_luts.py).
Run in the build container only:  python tests/golden/make_lutgen_golden.py   (~6 minutes)
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, REF)

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


def channel(design_db, qcu, qc):
    sigma = np.sqrt(1 / 10 ** (design_db / 10))
    E = 2 / sigma ** 2
    D = np.sqrt(2 * E)
    pyx, interval_x, quanta = channel_llr_density_table(qcu, -E - 3 * D, E + 3 * D, E, -E, D)
    dens, q, lut, _ = LLRQuantizer().find_OptLS_quantizer(pyx, quanta, qcu, qc)
    return np.asarray(pyx, np.float64).ravel(), np.asarray(quanta, np.float64).ravel(), dens.squeeze(), q.squeeze(), lut.squeeze()


def main():
    out = {}
    rng = np.random.default_rng(7)
    # stand-alone quantizer problems (sorted unique quanta, as np.unique hands them over)
    for i, (M, K) in enumerate([(17, 16), (40, 16), (64, 4), (129, 16), (256, 16), (300, 8), (33, 2), (520, 16)]):
        q = np.unique(np.round(rng.standard_normal(M) * 40) / 8 if i % 2 else rng.standard_normal(M) * 3)
        d = rng.random(q.size)
        d /= d.sum()
        od, oq, lut = _mdq.find_OptLS_quantizer(d.copy(), q.copy(), K)
        out[f"q{i}/d"], out[f"q{i}/q"], out[f"q{i}/K"] = d, q, np.int32(K)
        out[f"q{i}/od"], out[f"q{i}/oq"], out[f"q{i}/lut"] = od, oq, lut.astype(np.int32)
    out["nq"] = np.int32(8)
    # whole generator runs
    for tag, N, v, qcu, db in [("n16v4", 16, 4, 32, 2.0), ("n32v8", 32, 8, 64, 3.0), ("n64v16", 64, 16, 128, 3.0)]:
        pyx, cq, dens, quanta, clut = channel(db, qcu, v)
        llr_density, llr_quanta, lut_fs, lut_gs = LLRQuantizerSC(N, v).run(channel_llr_density=dens, channel_llr_quanta=quanta)
        out[tag + "/chan_pyx"], out[tag + "/chan_cells"] = pyx, cq
        out[tag + "/chan_density"], out[tag + "/chan_quanta"], out[tag + "/chan_lut"] = dens, quanta, clut.astype(np.int32)
        out[tag + "/llr_density"], out[tag + "/llr_quanta"] = np.asarray(llr_density), np.asarray(llr_quanta)
        out[tag + "/lut_f"] = np.stack([np.asarray(lut_fs[p][0], np.int32) for p in range(N - 1)])
        out[tag + "/lut_g"] = np.stack([np.asarray(lut_gs[p][0], np.int32) for p in range(N - 1)])
        print(tag, "done", flush=True)
    # channel inputs of the N=128 fixture (real_lut_n128.npz holds the tables the reference generator made from them)
    pyx, cq, _, _, _ = channel(3.0, 128, 16)
    out["n128v16/chan_pyx"], out["n128v16/chan_cells"] = pyx, cq
    np.savez_compressed(os.path.join(HERE, "lutgen_golden.npz"), **out)


if __name__ == "__main__":
    main()
