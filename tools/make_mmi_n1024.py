#!/usr/bin/env python
"""Probability-domain (MMI) lookup tables for BASELINE config 4 (N=1024, QDecoder=QChannel=16, DesignSNR=3.0 dB) made by this
package's own generator (lutgen.MMILUTGenerator: the reference's QDensityEvolutionMMI.run, quantizer passes on the GPU) ->
quantized_decoder_polar_codes_b200/data/mmi_n1024_q16_3dB.npz.  Also stores the MMI channel quantizer the probability-domain
driver builds at Eb/N0 = 1..4 dB for A=512 (mainQuantizedDecoder_ProbabilityDomain.py:137-152; symbols are cut at
interval_x[channel_lut], :176).  Needs a B200:  python tools/make_mmi_n1024.py [out.npz]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantized_decoder_polar_codes_b200 import lutgen, simulation as sim  # noqa: E402

N, QD, QC, QCU, DESIGN_DB = 1024, 16, 16, 128, 3.0


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "quantized_decoder_polar_codes_b200", "data", "mmi_n1024_q16_3dB.npz")
    t0 = time.time()
    sigma_d = np.sqrt(1 / 10 ** (DESIGN_DB / 10))                       # GenerateLookUpTable_ProbabilityDomain.py:42-43
    pzx, _, _ = lutgen.mmi_channel_quantizer(sigma_d, QCU, QC)
    lut_f, lut_g, llrs, _ = lutgen.MMILUTGenerator(N, QD).run(pzx)
    out = {"lut_f": np.stack(lut_f).astype(np.uint8), "lut_g": np.stack(lut_g).astype(np.uint8), "llrs": llrs, "design_pzx": pzx}
    for eb in [1.0, 2.0, 3.0, 4.0]:
        _, interval_x, channel_lut = lutgen.mmi_channel_quantizer(sim.awgn_sigma(eb, 512 / N), QCU, QC)
        out[f"chan_A512_eb{eb:.0f}/edges"] = np.asarray(interval_x[channel_lut], np.float64)
    np.savez_compressed(out_path, **out)
    print("MMI tables N=%d written to %s in %.1f s" % (N, out_path, time.time() - t0))


if __name__ == "__main__":
    main()
