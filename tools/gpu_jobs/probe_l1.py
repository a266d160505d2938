"""GPU probe: SC-LUT / FastSC-LUT (L=1) at N=1024 with growing batches, each in its own process under a timeout."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import sys, time, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r + "/tests")
import common
import quantized_decoder_polar_codes_b200 as q
from quantized_decoder_polar_codes_b200 import capi
kind, N, K, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
kw, x, _ = common.make_case(kind, N=N, K=K, B=min(B, 2048), seed=1, tables="minsum", ebn0_db=3.0)
x = np.tile(x, (-(-B // x.shape[0]), 1))[:B]
d_in = torch.from_numpy(x.astype(np.uint8)).cuda()
dec = getattr(q, kind)(**kw)
ko = capi.lib().pd_out_len(dec._handle)
d_out = torch.empty((B, ko), dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    capi.decode_device(dec, d_in.data_ptr(), capi.PD_U8, B, d_out.data_ptr(), s)
    capi.sync_check(dec, s)
    dt = time.time() - t0
print("OK", kind, N, B, "%%.3g frames/s" %% (B / dt), "wave", capi.wave_frames(dec, capi.PD_U8), flush=True)
''' % (ROOT, ROOT)

for kind, N, K in [("SCLUTDecoder", 1024, 512), ("SCLUTDecoder", 512, 256), ("FastSCLUTDecoder", 1024, 512)]:
    for B in [2048, 32768, 113664, 131072, 262144]:
        for env in [{}, {"POLAR_B200_WARPS_PER_CTA": "1"}]:
            try:
                r = subprocess.run([sys.executable, "-c", CHILD, kind, str(N), str(K), str(B)], capture_output=True, text=True, timeout=60,
                                   env=dict(os.environ, **env))
                print(env, (r.stdout.strip().splitlines() or ["(no output)"])[-1], r.stderr.strip()[-300:], flush=True)
            except subprocess.TimeoutExpired:
                print(env, "TIMEOUT", kind, N, B, flush=True)
