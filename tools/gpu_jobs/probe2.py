"""GPU probe 2: which knob removes the L=1 full-residency failure.  Each case in its own process under a timeout."""
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
from oracle import polar_oracle as po
kind, N, K, L, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
kw, x, _ = common.make_case(kind, N=N, K=K, L=L, B=min(B, 1024), seed=1, tables="minsum", ebn0_db=3.0)
want = po.OracleDecoder(kind, **kw).decode(x[:64].astype(np.int32))
x = np.tile(x, (-(-B // x.shape[0]), 1))[:B]
d_in = torch.from_numpy(x.astype(np.uint8)).cuda()
dec = getattr(q, kind)(**kw)
ko = capi.lib().pd_out_len(dec._handle)
d_out = torch.empty((B, ko), dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.time()
    capi.decode_device(dec, d_in.data_ptr(), capi.PD_U8, B, d_out.data_ptr(), s)
    capi.sync_check(dec, s)
    dt = time.time() - t0
got = d_out.cpu().numpy()
bad = int((got[:64] != want).any(axis=1).sum()) + int((got[-1024:][:64] != want).any(axis=1).sum()) if B %% 1024 == 0 else int((got[:64] != want).any(axis=1).sum())
print("OK", kind, N, "L", L, "B", B, "%%.3g frames/s" %% (B / dt), "wave", capi.wave_frames(dec, capi.PD_U8), "mismatch(first/last 64)", bad, flush=True)
''' % (ROOT, ROOT)


def run(tag, kind, N, K, L, B, env, timeout=45, prefix=()):
    try:
        r = subprocess.run(list(prefix) + [sys.executable, "-c", CHILD, kind, str(N), str(K), str(L), str(B)], capture_output=True, text=True,
                           timeout=timeout, env=dict(os.environ, **env))
        out = (r.stdout.strip().splitlines() or ["(no output)"])
        print(tag, env, out[-1], "|", r.stderr.strip()[-200:].replace("\n", " "), flush=True)
        return r
    except subprocess.TimeoutExpired:
        print(tag, env, "TIMEOUT", kind, N, L, B, flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "knobs"):
        for env in [{"POLAR_B200_CTAS_PER_SM": "5"}, {"POLAR_B200_NO_TMA": "1"}, {"POLAR_B200_WARPS_PER_CTA": "2"}, {"POLAR_B200_WARPS_PER_CTA": "3"}, {}]:
            run("E1", "SCLUTDecoder", 1024, 512, 1, 113664, env)
    if which in ("flags",):
        for env in [{"POLAR_B200_NO_TMA": "4"}, {"POLAR_B200_NO_TMA": "12"}, {"POLAR_B200_NO_TMA": "8"}, {"POLAR_B200_NO_TMA": "1"}]:
            run("F", "SCLUTDecoder", 1024, 512, 1, 113664, env)
            run("F", "SCLUTDecoder", 256, 128, 1, 1 << 19, env)
    if which in ("verify",):
        for env in [{"POLAR_B200_NO_TMA": "2"}, {"POLAR_B200_NO_TMA": "2", "POLAR_B200_WARPS_PER_CTA": "2"}]:
            run("V", "SCLUTDecoder", 1024, 512, 1, 113664, env)
            run("V", "SCLUTDecoder", 256, 128, 1, 1 << 19, env)
        run("V", "SCLLUTDecoder", 1024, 512, 8, 14208 * 4, {"POLAR_B200_NO_TMA": "2"})
    if which in ("all", "lists"):
        run("E4", "SCLLUTDecoder", 512, 256, 8, 1 << 17, {})
        run("E4", "SCLLUTDecoder", 256, 128, 8, 1 << 18, {})
        run("E4", "SCLLUTDecoder", 1024, 512, 2, 1 << 17, {})
        run("E4", "SCLLUTDecoder", 1024, 512, 4, 1 << 17, {})
        run("E4", "SCLLUTDecoder", 512, 256, 4, 1 << 18, {})
        run("E4", "SCLUTDecoder", 256, 128, 1, 1 << 19, {})
    if which in ("all", "memcheck"):
        r = run("E5", "FastSCLUTDecoder", 1024, 512, 1, 262144, {}, timeout=280,
                prefix=("compute-sanitizer", "--tool", "memcheck", "--print-limit", "5", "--launch-timeout", "120"))
        if r is not None:
            print(r.stdout[-3000:])
