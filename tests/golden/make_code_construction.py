"""tests/golden/code_construction.npz: frozen sets and node types produced by the REFERENCE's own code
(PolarCodesUtils/CodeConstruction.py PW / GA, IdentifyNodes.py NodeIdentifier.run) for the shapes the benchmark and the
BASELINE configs use.  Run in the build container only."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from quantized_decoder_polar_codes_b200 import compat, simulation as sim  # noqa: E402

compat.install(decoders=False, encoder=False, quantizers=False)
REF = "/root/reference"
sys.path.insert(0, REF)
from PolarCodesUtils.CodeConstruction import PolarCodeConstructor  # noqa: E402
from PolarCodesUtils.IdentifyNodes import NodeIdentifier  # noqa: E402

out = {}
for N, K in [(128, 64), (1024, 512), (1024, 536)]:
    fb, mb, fmask, _ = PolarCodeConstructor(N, K, os.path.join(REF, "reliable sequence.txt")).PW()
    out[f"pw_{N}_{K}/frozen"] = fmask.astype(np.int32)
    out[f"pw_{N}_{K}/node_type"] = NodeIdentifier(N, K, fb, mb, use_new_node=False).run().astype(np.int32)
sigma = sim.awgn_sigma(2.0, 0.5)
_, _, fmask, _ = PolarCodeConstructor(2048, 1024, os.path.join(REF, "reliable sequence.txt")).GA(sigma)
out["ga_2048_1024/frozen"], out["ga_2048_1024/sigma"] = fmask.astype(np.int32), np.float64(sigma)
from QuantizeDensityEvolution.QLLRDensityEvolution_OptUniform import LLRLSUniformQuantizer  # noqa: E402
out["uq_2048_16/r_f"], out["uq_2048_16/r_g"] = LLRLSUniformQuantizer(2048, 16).generate_uniform_quantizers(sigma)
np.savez_compressed(os.path.join(HERE, "code_construction.npz"), **out)
print("written")
