#!/usr/bin/env python
"""Benchmark of the decode hot path on the north-star shape (BASELINE.json): SCL-LUT, N=1024, A=K=512, L=8,
QDecoder=16, synthetic AWGN frames.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch F]

A "step" decodes one batch of F frames per GPU (inputs > L2 so no flush is needed).  Prints ONE JSON line:
  value     decoded frames/s over all GPUs, inputs resident in HBM (pd_decode_device), CUDA-event timed, max over ranks
  e2e       the same metric through the host-buffer C-ABI call pd_decode (pinned host in/out, H2D+D2H inside)
  roofline  algorithmic bytes (N symbol bytes in + K bit bytes out per frame) / kernel time vs measured HBM peak
  cpu_baseline  the compiled reference (oracle/_ref), one process per host core, on a bounded sample (N=1 only)
`--impl reference` times the reference's own CPU implementation on the same workload instead.
Multi-GPU: one process per GPU (torchrun); frames are sharded, no data-path collective; the only exchange is the
NCCL all-reduce of the two error counters (bit / block errors), inside the timed region.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N, K, L, Q = 1024, 512, 8, 16
EBN0_DB = 2.0
WORKLOAD = "SCL-LUT N=1024 A=K=512 L=8 QDecoder=QChannel=16, MinDistortion LUTs (reference generator, design 3 dB), AWGN Eb/N0=2.0 dB through the driver's channel quantizer, NR-sequence frozen set"
METRIC = "decoded frames/s (info Gbit/s = frames/s*512/1e9), N=1024 L=8 SCL-LUT"
BYTES_PER_FRAME = N + K  # uint8 symbols in + uint8 bits out


def make_workload(frames, seed):
    """Real MinDistortion tables for N=1024, Q=16, design SNR 3 dB, produced by the reference's own generator code
    (tests/golden/make_real_lut_n1024.py), the channel quantizer the reference driver builds at this Eb/N0
    (mainQuantizedDecoder_LLRDomain.py:130-145) and its per-symbol rule (:167-176), vectorised."""
    from quantized_decoder_polar_codes_b200 import simulation as sim
    rng = np.random.default_rng(seed)
    z = np.load(os.path.join(ROOT, "quantized_decoder_polar_codes_b200", "data", "mindistortion_n1024_q16_3dB.npz"))
    fm, mm = sim.frozen_mask(N, K)
    f = [z["lut_f"][p].astype(np.int32)[None] for p in range(N - 1)]
    g = [z["lut_g"][p].astype(np.int32)[None] for p in range(N - 1)]
    llr_tab = z["llr_quanta"]
    edges, clut = z[f"chan_A{K}_eb{EBN0_DB:.0f}/edges"], z[f"chan_A{K}_eb{EBN0_DB:.0f}/lut"]
    msg = rng.integers(0, 2, (frames, K), dtype=np.uint8)
    cw = sim.polar_encode(msg, fm)
    llr = sim.awgn_llr(cw, sim.awgn_sigma(EBN0_DB, K / N), rng)
    idx = np.clip(np.searchsorted(edges[:-1], llr, side="left") - 1, 0, clut.size - 1)
    sym = np.where(llr <= edges[0], 0, np.where(llr >= edges[-1], Q - 1, clut[idx])).astype(np.uint8)
    kw = dict(N=N, K=K, L=L, frozen_bits=fm, message_bits=mm, LUT_f=f, LUT_g=g, virtual_channel_llr=llr_tab)
    return kw, sym, msg


def tile_frames(sym, msg, frames):
    reps = -(-frames // sym.shape[0])
    return np.tile(sym, (reps, 1))[:frames], np.tile(msg, (reps, 1))[:frames]


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md): NVML polled every 10 ms from a
    thread (nvidia-smi as a fallback, which is too slow for sub-second regions)."""
    SMI_Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.sm, self.mx, self.reasons = [], 0.0, set()
        self.stop = False
        self.th = None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            phys = gpu_index
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                phys = int(vis.split(",")[gpu_index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll_nvml(self):
        n = self.nvml
        bits = {"hw_slowdown": n.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": n.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": n.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": n.nvmlClocksThrottleReasonSwPowerCap}
        while not self.stop:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                self.mx = max(self.mx, float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def _poll_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.SMI_Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                r = [t.strip() for t in out.split(",")]
                self.sm.append(float(r[0]))
                self.mx = max(self.mx, float(r[1]))
                for nme, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        self.th = threading.Thread(target=self._poll_nvml if self.nvml else self._poll_smi, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx or None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml" if self.nvml else "nvidia-smi"}


# ---------------------------------------------------------------------------------------------------
def _ref_worker(args):
    kw_small, sym, core = args
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import polar_oracle as po
    import common
    ref = po.load_reference()
    t0 = time.perf_counter()
    if ref is not None:
        dec = ref.SCLLUTDecoder(**common.ref_kwargs(kw_small))
        x = sym.astype(np.int32)
        t0 = time.perf_counter()
        for i in range(x.shape[0]):
            dec.decode(x[i])
        kind = "reference"
    else:
        dec = po.OracleDecoder("SCLLUTDecoder", **kw_small)
        t0 = time.perf_counter()
        dec.decode(sym.astype(np.int32))
        kind = "port"
    return time.perf_counter() - t0, kind


def cpu_reference_throughput(kw, sym, frames_per_core, cores=None):
    """The reference's own CPU decode() (oracle/_ref) on `cores` pinned processes, disjoint shards of the same
    pre-generated inputs; frames/s = total frames / slowest process wall time."""
    import multiprocessing as mp
    avail = sorted(os.sched_getaffinity(0))
    cores = cores or len(avail)
    jobs = []
    for c in range(cores):
        shard = sym[(c * frames_per_core) % sym.shape[0]:][:frames_per_core]
        if shard.shape[0] < frames_per_core:
            shard = np.tile(sym, (-(-frames_per_core // sym.shape[0]), 1))[:frames_per_core]
        jobs.append((kw, shard, avail[c % len(avail)]))
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_ref_worker, jobs)
    wall = max(r[0] for r in res)
    return cores * frames_per_core / wall, cores, res[0][1], wall


# ---------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    kw, sym, _ = make_workload(256, seed=0)
    fpc = args.ref_frames
    times = []
    for it in range(args.warmup + args.steps):
        fps, cores, kind, wall = cpu_reference_throughput(kw, sym, fpc)
        if it >= args.warmup:
            times.append((fps, wall))
    fps = float(np.mean([t[0] for t in times]))
    ms = float(np.mean([t[1] for t in times]) * 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": cores * fpc, "info_gbit_s": fps * K / 1e9},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                         "sample": f"{fpc} frames per core per step, one pinned process per core, per-frame decode() calls"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    import quantized_decoder_polar_codes_b200 as q
    from quantized_decoder_polar_codes_b200 import capi
    from quantized_decoder_polar_codes_b200 import distributed as D

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    D.init("nccl", dev)
    kw, sym0, msg0 = make_workload(8192, seed=rank)
    dec = q.SCLLUTDecoder(device=local_rank, **kw)
    lib = capi.lib()
    # frames per step: --batch, or by default the multiple of the kernel's wave (SMs x resident warps x frames per warp)
    # nearest to 131072 -- a persistent kernel then has no tail
    F = args.batch
    wave = capi.wave_frames(dec, capi.PD_U8)
    if F <= 0:
        F = max(1, round(131072 / wave)) * wave if wave > 0 else 131072
    sym, msg = tile_frames(sym0, msg0, F)
    stream = torch.cuda.current_stream().cuda_stream

    d_in = torch.from_numpy(sym).to(dev)
    d_truth = torch.from_numpy(msg).to(dev)
    d_out = torch.empty((F, K), dtype=torch.uint8, device=dev)
    counters = torch.zeros(2, dtype=torch.int64, device=dev)   # run totals (identical on every rank)
    step_cnt = torch.zeros(2, dtype=torch.int64, device=dev)   # this step's local counts -> all-reduced -> added to the totals
    assert d_in.numel() >= 120 * 2 ** 20 or 0 < args.batch < 131072, "inputs + outputs of a step must exceed L2"

    def count_and_reduce():
        step_cnt.zero_()
        capi.check(lib.pd_count_errors(d_out.data_ptr(), d_truth.data_ptr(), F, K, step_cnt.data_ptr(), stream))
        D.allreduce_counters(step_cnt)   # the path's only exchange: 2 x int64 over NCCL/NVLink
        counters.add_(step_cnt)

    def step_device():
        capi.decode_device(dec, d_in.data_ptr(), capi.PD_U8, F, d_out.data_ptr(), stream)
        count_and_reduce()

    def barrier():
        D.barrier()
        torch.cuda.synchronize()

    launches0 = lib.pd_launch_count()
    for _ in range(args.warmup):
        counters.zero_()
        step_device()
    barrier()
    # --- device-resident timing: whole step and the decode kernel alone (roofline) ---
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    counters.zero_()
    with ClockSampler(local_rank) as clk:
        barrier()
        l_before = lib.pd_launch_count()
        ev[0].record()
        for i in range(args.steps):
            kev[i][0].record()
            capi.decode_device(dec, d_in.data_ptr(), capi.PD_U8, F, d_out.data_ptr(), stream)
            kev[i][1].record()
            count_and_reduce()
        ev[1].record()
        barrier()
        l_after = lib.pd_launch_count()
    capi.sync_check(dec, stream)
    elapsed_ms = ev[0].elapsed_time(ev[1])
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    elapsed_ms = D.max_over_ranks(elapsed_ms, dev)
    kernel_ms = D.max_over_ranks(kernel_ms, dev)
    value = world * F * args.steps / (elapsed_ms / 1e3)
    cnt = counters.cpu().tolist()
    total_frames = world * F * args.steps

    # --- end to end through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside) ---
    h_in_p = lib.pd_host_alloc(F * N)
    h_out_p = lib.pd_host_alloc(F * K)
    h_in = np.ctypeslib.as_array(ctypes.cast(h_in_p, ctypes.POINTER(ctypes.c_uint8)), (F, N))
    h_out = np.ctypeslib.as_array(ctypes.cast(h_out_p, ctypes.POINTER(ctypes.c_uint8)), (F, K))
    h_in[:] = sym
    for _ in range(max(1, args.warmup // 2)):
        capi.decode_host(dec, h_in_p, capi.PD_U8, F, h_out_p)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        capi.decode_host(dec, h_in_p, capi.PD_U8, F, h_out_p)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_value = world * F * args.steps / D.max_over_ranks(e2e_s, dev)
    same = bool((h_out == d_out.cpu().numpy()).all())
    lib.pd_host_free(h_in_p)
    lib.pd_host_free(h_out_p)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = BYTES_PER_FRAME * F / (kernel_ms / 1e3) / 1e9
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel, per frame
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("kernel") == dec.kernel:
                traffic = tj["dram_bytes_per_frame"] * F
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": F, "wave_frames": wave, "info_gbit_s": value * K / 1e9, "kernel": dec.kernel,
                       "l2": "inputs+outputs per step exceed L2 (no flush needed)" if F * BYTES_PER_FRAME > 126 * 2 ** 20 else "batch below L2 size",
                       "bit_errors": cnt[0], "block_errors": cnt[1], "frames_counted": total_frames,
                       "e2e_equals_device_output": same},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                         "kernel_ms": kernel_ms, "bytes_per_frame": BYTES_PER_FRAME,
                         "per": "decode call of one step = 4 scl_lut_warp launches overlapped on two internal streams; algorithmic bytes, kernel_ms and traffic all refer to that call",
                         "note": "HBM-nominal codec path; the kernel is SM-issue bound, not HBM bound (see DESIGN.md 4.1, profiles/)"},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": F * N, "d2h_bytes_per_step": F * K},
            "gpu_launches": int(l_after - l_before),
            "clocks": clk.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            try:
                fps, cores, kind, wall = cpu_reference_throughput(kw, sym0[:256], args.ref_frames)
                line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind,
                                        "sample": f"{args.ref_frames} frames per core ({wall:.1f} s), one pinned process per core, per-frame decode() calls of the compiled reference"}
            except Exception as e:  # pragma: no cover
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on stdout when
    NCCL_DEBUG is set): keep a private handle on the real stdout for the result and point fd 1 at stderr for everyone else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(text):
    out = _REAL_STDOUT or sys.stdout
    out.write(text + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="frames per step per GPU (default: the whole number of kernel waves nearest to 131072)")
    ap.add_argument("--ref-frames", type=int, default=400, help="CPU reference: frames per core per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and args.gpus > 1 and "RANK" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    protect_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
