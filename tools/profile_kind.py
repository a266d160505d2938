#!/usr/bin/env python
"""One decode launch of a named shape from tools/bench_kinds.py (for ncu): python tools/profile_kind.py C4 [frames]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
import common  # noqa: E402
import quantized_decoder_polar_codes_b200 as q  # noqa: E402
from quantized_decoder_polar_codes_b200 import capi  # noqa: E402
from bench_kinds import SHAPES  # noqa: E402

name, kind, ckw, frames = [s for s in SHAPES if sys.argv[1] in s[0]][0]
F = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
kw, x, _ = common.make_case(kind, B=min(F, 2048), seed=1, ebn0_db=3.0, **ckw)
lut = "LUT" in kind
x = np.tile(x, (-(-F // x.shape[0]), 1))[:F]
d_in = torch.from_numpy(x.astype(np.uint8) if lut else x.astype(np.float64)).cuda()
dec = getattr(q, kind)(**kw)
d_out = torch.empty((F, capi.lib().pd_out_len(dec._handle)), dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    capi.decode_device(dec, d_in.data_ptr(), capi.PD_U8 if lut else capi.PD_F64, F, d_out.data_ptr(), stream)
capi.sync_check(dec, stream)
print(name, dec.kernel, "ok")
