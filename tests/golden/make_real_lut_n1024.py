"""MinDistortion lookup tables for the north-star shape (N=1024, QDecoder=QChannel=16, DesignSNR=3.0 dB) from the
REFERENCE's generator code (see make_real_luts.py for the shim) -> quantized_decoder_polar_codes_b200/data/
mindistortion_n1024_q16_3dB.npz, used by bench.py so that the benchmark decodes with real tables.  Also stores
the channel quantizer (interval edges + lut) the driver builds at Eb/N0 = 1..4 dB for A=512.
Pure-Python density evolution: ~30-40 minutes.  Run in the build container only."""
import os
import runpy
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
mod = runpy.run_path(os.path.join(HERE, "make_real_luts.py"), run_name="shim_only")   # installs the shim, no main()
LLRQuantizerSC, channel_quantizer = mod["LLRQuantizerSC"], mod["channel_quantizer"]
sys.path.insert(0, ROOT)
from quantized_decoder_polar_codes_b200 import simulation as sim  # noqa: E402

N, QD, DESIGN_DB = 1024, 16, 3.0


def main():
    sigma_d = np.sqrt(1 / 10 ** (DESIGN_DB / 10))
    dens, quanta, _, _ = channel_quantizer(sigma_d)
    llr_density, llr_quanta, lut_fs, lut_gs = LLRQuantizerSC(N, QD).run(channel_llr_density=dens, channel_llr_quanta=quanta)
    out = {
        "lut_f": np.stack([np.asarray(lut_fs[p][0], np.uint8) for p in range(N - 1)]),
        "lut_g": np.stack([np.asarray(lut_gs[p][0], np.uint8) for p in range(N - 1)]),
        "llr_quanta": np.asarray(llr_quanta, np.float64),
    }
    for eb in [1.0, 2.0, 3.0, 4.0]:
        for A in (512,):
            _, _, clut, ix = channel_quantizer(sim.awgn_sigma(eb, A / N))
            out[f"chan_A{A}_eb{eb:.0f}/lut"] = clut.astype(np.uint8)
            out[f"chan_A{A}_eb{eb:.0f}/edges"] = np.asarray(ix, np.float64)
    np.savez_compressed(os.path.join(ROOT, "quantized_decoder_polar_codes_b200", "data", "mindistortion_n1024_q16_3dB.npz"), **out)
    print("done")


if __name__ == "__main__":
    main()
