#!/usr/bin/env python
"""Device-resident throughput of every BASELINE.json config shape (and a few more) on one GPU.
    python tools/bench_kinds.py [--frames F] [--only substr]
Prints one JSON line per shape: frames/s, info Gbit/s, kernel used, algorithmic GB/s vs the measured HBM peak."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SHAPES = [
    ("C1 float SC N=128 A=64", "SCDecoder", dict(N=128, K=64, tables="channel"), 1 << 18),
    ("C2 SC-LUT N=128 A=32 Q=16", "SCLUTDecoder", dict(N=128, K=32, tables="minsum"), 1 << 20),
    ("C3 SCL-LUT N=128 A=32 L=8 Q=16", "SCLLUTDecoder", dict(N=128, K=32, L=8, tables="minsum"), 1 << 19),
    ("NS SCL-LUT N=1024 A=512 L=8 Q=16", "SCLLUTDecoder", dict(N=1024, K=512, L=8, tables="minsum"), 1 << 17),
    ("SC-LUT N=1024 A=512", "SCLUTDecoder", dict(N=1024, K=512, tables="minsum"), 1 << 18),
    ("CASCL-LUT N=1024 A=512 L=8", "CASCLLUTDecoder", dict(N=1024, K=536, A=512, L=8, tables="minsum"), 1 << 17),
    ("C4 CAFastSCL-LUT N=1024 A=512 L=8", "CAFastSCLLUTDecoder", dict(N=1024, K=536, A=512, L=8, tables="minsum"), 1 << 17),
    ("FastSC-LUT N=1024 A=512", "FastSCLUTDecoder", dict(N=1024, K=512, tables="minsum"), 1 << 18),
    ("float SCL N=1024 A=512 L=8", "SCLDecoder", dict(N=1024, K=512, L=8, tables="channel"), 1 << 16),
    ("float FastSCL N=1024 A=512 L=8", "FastSCLDecoder", dict(N=1024, K=512, L=8, tables="channel"), 1 << 16),
    ("C5 SCL-Uniform N=2048 A=1024 L=32 v=16", "SCLUniformQuantizedDecoder", dict(N=2048, K=1024, L=32, construction="pw"), 1 << 13),
    ("SCL-Lloyd N=1024 A=512 L=8", "SCLLloydQuantizedDecoder", dict(N=1024, K=512, L=8), 1 << 15),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import common
    import quantized_decoder_polar_codes_b200 as q
    from quantized_decoder_polar_codes_b200 import capi
    peak = 6650.0
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    stream = torch.cuda.current_stream().cuda_stream
    for name, kind, ckw, frames in SHAPES:
        if args.only and args.only not in name:
            continue
        F = args.frames or frames
        kw, x, _ = common.make_case(kind, B=min(F, 2048), seed=1, ebn0_db=3.0, **ckw)
        lut = "LUT" in kind
        x = np.tile(x, (-(-F // x.shape[0]), 1))[:F]
        d_in = torch.from_numpy(x.astype(np.uint8) if lut else x.astype(np.float64)).cuda()
        dec = getattr(q, kind)(**kw)
        kout = capi.lib().pd_out_len(dec._handle)
        d_out = torch.empty((F, kout), dtype=torch.uint8, device="cuda")
        dt = capi.PD_U8 if lut else capi.PD_F64

        def run():
            capi.decode_device(dec, d_in.data_ptr(), dt, F, d_out.data_ptr(), stream)
        run()
        capi.sync_check(dec, stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(args.reps):
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        fps = F / (best / 1e3)
        a = kw.get("A", kw["K"])
        bpf = kw["N"] * (1 if lut else 8) + kout
        print(json.dumps({"shape": name, "class": kind, "kernel": dec.kernel, "frames": F, "ms": round(best, 3), "frames_per_s": fps,
                          "info_gbit_s": fps * a / 1e9, "alg_bytes_per_frame": bpf, "alg_GBps": fps * bpf / 1e9,
                          "hbm_frac_of_measured": fps * bpf / 1e9 / peak}), flush=True)
    bench_bd(args, peak)


def bench_bd(args, peak):
    """The two blind-detection kinds (pd_decode_bd_device): D-metric N=1024 K=512, CA-SCL+RNTI N=512 A=100 L=8."""
    import torch
    import ctypes as C
    import quantized_decoder_polar_codes_b200 as q
    from quantized_decoder_polar_codes_b200 import capi
    from quantized_decoder_polar_codes_b200 import simulation as sim
    lib = capi.lib()
    stream = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(0)
    for name, N, K, A, L, F in [("BD D-metric N=1024 K=512", 1024, 512, 0, 1, 1 << 18), ("BD CA-SCL+RNTI N=512 A=100 L=8", 512, 124, 100, 8, 1 << 16)]:
        if args.only and args.only not in name:
            continue
        fm, mm = sim.frozen_mask(N, K)
        if A:
            dec = q.BDCASCLDecoder(N, K, A, L, fm, mm, 24, list(sim.CRC24_LOC))
        else:
            dec = q.BDDMetricCalculator(N, K, fm, mm, sim.identify_nodes(N, fm))
        msg = rng.integers(0, 2, (2048, K), dtype=np.uint8)
        llr = sim.awgn_llr(sim.polar_encode(msg, fm), sim.awgn_sigma(2.0, K / N), rng)
        d_in = torch.from_numpy(np.tile(llr, (F // 2048, 1))).cuda()
        d_bits = torch.empty((F, max(A, 1)), dtype=torch.uint8, device="cuda")
        d_metric = torch.empty(F, dtype=torch.float64, device="cuda")
        d_pass = torch.empty(F, dtype=torch.uint8, device="cuda")
        d_rnti = torch.zeros(16, dtype=torch.int32, device="cuda")

        def run():
            capi.check(lib.pd_decode_bd_device(dec._handle, d_in.data_ptr(), F, d_rnti.data_ptr(), 16 if A else 0,
                                               d_bits.data_ptr(), d_metric.data_ptr(), d_pass.data_ptr(), stream))
        run()
        capi.sync_check(dec, stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(args.reps):
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        fps = F / (best / 1e3)
        bpf = 8 * N + (A + 9 if A else 8)
        print(json.dumps({"shape": name, "kernel": dec.kernel, "frames": F, "ms": round(best, 3), "frames_per_s": fps,
                          "alg_bytes_per_frame": bpf, "alg_GBps": fps * bpf / 1e9, "hbm_frac_of_measured": fps * bpf / 1e9 / peak}), flush=True)


if __name__ == "__main__":
    main()
