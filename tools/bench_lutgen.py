#!/usr/bin/env python
"""Wall time of the GPU table generator (lutgen.MinDistortionLUTGenerator) for N=128 and N=1024, Q=16, and the share spent
in the batched quantizer kernel (pd_optls_quantize) vs the numpy glue.  The reference's pure-Python path needs ~4 minutes
for N=128 (measured in the build container) and scales with N.  python tools/bench_lutgen.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from quantized_decoder_polar_codes_b200 import lutgen  # noqa: E402

v = 16
# the channel the reference driver designs for 3 dB (128 uniform LLR cells -> 16 symbols), inputs from the golden file
gold = np.load(os.path.join(ROOT, "tests", "golden", "lutgen_golden.npz"))
od, oq, _ = lutgen.optls_quantize_batch([gold["n128v16/chan_pyx"]], [gold["n128v16/chan_cells"]], v)
ref = np.load(os.path.join(ROOT, "quantized_decoder_polar_codes_b200", "data", "mindistortion_n1024_q16_3dB.npz"))
for N in (128, 1024):
    gen = lutgen.MinDistortionLUTGenerator(N, v)
    kernel_s = [0.0]
    orig = lutgen.optls_quantize_batch

    def timed(*a, **k):
        t = time.perf_counter()
        r = orig(*a, **k)
        kernel_s[0] += time.perf_counter() - t
        return r
    lutgen.optls_quantize_batch = timed
    t0 = time.perf_counter()
    dens, quan, lut_f, lut_g = gen.run(od[0], oq[0])
    wall = time.perf_counter() - t0
    lutgen.optls_quantize_batch = orig
    print(json.dumps({"N": N, "v": v, "nodes": N - 1, "quantizer_problems": 2 * (N - 1), "wall_s": round(wall, 3),
                      "pd_optls_quantize_s": round(kernel_s[0], 3), "numpy_glue_s": round(wall - kernel_s[0], 3),
                      "equals_reference_tables": bool((lut_f == ref["lut_f"]).all() and (lut_g == ref["lut_g"]).all() and (quan == ref["llr_quanta"]).all()) if N == 1024 else None}))
