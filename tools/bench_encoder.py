#!/usr/bin/env python
"""Device-resident throughput of the batched polar encoder (csrc/pb_enc.cuh), a pure streaming kernel:
algorithmic bytes per frame = in_len + N (one byte per bit).  python tools/bench_encoder.py [--frames F]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import torch
    from quantized_decoder_polar_codes_b200 import simulation as sim
    from quantized_decoder_polar_codes_b200.encoder import PD_ENC_CRC_POLAR, PD_ENC_POLAR, _Handle
    peak = 6650.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    stream = torch.cuda.current_stream().cuda_stream
    for name, N, A, crc in [("polar N=1024 K=512", 1024, 512, 0), ("crc24+polar N=1024 A=512 K=536", 1024, 512, 24),
                            ("polar N=128 K=64", 128, 64, 0), ("polar N=4096 K=2048", 4096, 2048, 0)]:
        K = A + crc
        fm = sim.frozen_mask(N, K)[0] if N <= 1024 else np.r_[np.ones(N - K, np.int32), np.zeros(K, np.int32)]
        h = _Handle(N, K, A, fm, crc, list(sim.CRC24_LOC) if crc else None)
        F = max(1024, args.frames * 1024 // N)
        d_in = torch.randint(0, 2, (F, A), dtype=torch.uint8, device="cuda")
        d_out = torch.empty((F, N), dtype=torch.uint8, device="cuda")
        mode = PD_ENC_CRC_POLAR if crc else PD_ENC_POLAR
        for _ in range(3):
            h.run_device(mode, d_in.data_ptr(), F, d_out.data_ptr(), stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.reps):
            h.run_device(mode, d_in.data_ptr(), F, d_out.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        gbs = F * (A + N) / ms / 1e6
        print(json.dumps({"op": name, "frames": F, "ms": round(ms, 4), "frames_per_s": F / ms * 1e3, "alg_bytes_per_frame": A + N,
                          "alg_GBps": round(gbs, 1), "hbm_peak_GBps": peak, "frac": round(gbs / peak, 3)}))


if __name__ == "__main__":
    main()
