"""Parity at the sizes SURVEY.md 8(d) asks for: every BASELINE.json configuration decoded on the GPU and by the
COMPILED REFERENCE (oracle/_ref, one pinned process per host core; the C port if it was not built) on the same frames
-- real encode -> AWGN -> channel quantizer inputs at several Eb/N0 -- and compared bit for bit.

The default sizes keep the whole file around a minute on a 16-core box.  POLAR_B200_SWEEP=full runs the full counts
(1e5 frames per SNR point for the N=128 shapes, 2e4 for N=1024, 256 for N=2048/L=32); POLAR_B200_SWEEP_LOG=<file>
appends one JSON line per point (the committed run is profiles/r1/parity_sweep.jsonl)."""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np
import pytest

import common
import real_lut
from quantized_decoder_polar_codes_b200 import simulation as sim

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FULL = os.environ.get("POLAR_B200_SWEEP", "") == "full"
LOG = os.environ.get("POLAR_B200_SWEEP_LOG", "")


@pytest.fixture(scope="module")
def q():
    import quantized_decoder_polar_codes_b200 as q
    return q


def _worker(job):
    kind, kw, x, core = job
    try:
        os.sched_setaffinity(0, {core})
    except Exception:
        pass
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import polar_oracle as po
    import common as cm
    ref = po.load_reference()
    if ref is not None:
        return cm.ref_decode(ref, kind, kw, x), "reference"
    xx = x.astype(np.int32) if "LUT" in kind else x.astype(np.float64)
    return po.OracleDecoder(kind, **kw).decode(xx), "port"


def reference_outputs(kind, kw, x):
    """decode() of the reference on every frame of x, sharded over the host cores."""
    cores = sorted(os.sched_getaffinity(0))
    n = min(len(cores), max(1, x.shape[0] // 8))
    shards = np.array_split(np.arange(x.shape[0]), n)
    jobs = [(kind, kw, x[s], cores[i]) for i, s in enumerate(shards)]
    if n == 1:
        res = [_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(n) as pool:
            res = pool.map(_worker, jobs)
    return np.concatenate([r[0] for r in res]), res[0][1], n


def channel_frames(N, K, A, crc, ebn0_db, frames, seed, edges=None, clut=None, Q=16, construction="nr"):
    """msg -> (CRC) -> polar encode -> BPSK/AWGN -> LLR (mainFPDecoder.py:91-113) -> the LLR-domain driver's channel
    quantizer when edges/clut are given (mainQuantizedDecoder_LLRDomain.py:167-176)."""
    rng = np.random.default_rng(seed)
    fm, _ = sim.frozen_mask(N, K, construction)
    msg = rng.integers(0, 2, (frames, A), dtype=np.uint8)
    word = sim.crc_attach(msg) if crc else msg
    llr = sim.awgn_llr(sim.polar_encode(word, fm), sim.awgn_sigma(ebn0_db, A / N), rng)
    if edges is None:
        return llr, msg
    idx = np.clip(np.searchsorted(edges[:-1], llr, side="left") - 1, 0, clut.size - 1)
    sym = np.where(llr <= edges[0], 0, np.where(llr >= edges[-1], Q - 1, clut[idx])).astype(np.uint8)
    return sym, msg


def _check(q, label, kind, kw, x, msg, ebn0):
    t0 = time.perf_counter()
    dec = getattr(q, kind)(**kw)
    got = dec.decode(x)
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    want, how, procs = reference_outputs(kind, kw, x)
    t_ref = time.perf_counter() - t0
    bad = int((got != want).any(axis=1).sum())
    rec = dict(config=label, cls=kind, kernel=dec.kernel, ebn0_db=ebn0, frames=int(x.shape[0]), mismatching_frames=bad,
               bler=float((got != msg).any(axis=1).mean()), bler_reference=float((want != msg).any(axis=1).mean()),
               ber=float((got != msg).mean()), checker=how, checker_processes=procs,
               gpu_decode_call_s=round(t_gpu, 3), checker_s=round(t_ref, 3))
    if LOG:
        with open(LOG, "a") as f:
            f.write(json.dumps(rec) + "\n")
    assert bad == 0, rec
    assert rec["bler"] == rec["bler_reference"]


@pytest.mark.parametrize("ebn0", [0, 1, 2, 3, 4] if FULL else [1, 3])
def test_c1_float_sc_n128(q, ebn0):
    N, K = 128, 64
    x, msg = channel_frames(N, K, K, False, ebn0, 100_000 if FULL else 20_000, seed=100 + ebn0)
    fm, mm = sim.frozen_mask(N, K)
    _check(q, "C1 float SC N=128 A=64", "SCDecoder", dict(N=N, K=K, frozen_bits=fm, message_bits=mm), x, msg, ebn0)


@pytest.mark.parametrize("ebn0", [1, 2, 3])
@pytest.mark.parametrize("kind", ["SCLUTDecoder", "SCLLUTDecoder"])
def test_c2_c3_lut_n128_real_tables(q, kind, ebn0):
    """C2 / C3: N=128, A=K=32, Q=16 real MinDistortion tables (design 3 dB), L=8 for the list decoder.  The channel
    quantizers shipped in the fixture cover Eb/N0 = 1, 2, 3 dB."""
    z = real_lut.load()
    tag = f"A32_eb{ebn0}"
    kw, _, _, _ = real_lut.build_kwargs(z, tag, kind)
    frames = (100_000 if FULL else (20_000 if kind == "SCLUTDecoder" else 6_000))
    x, msg = channel_frames(128, 32, 32, False, ebn0, frames, seed=200 + ebn0, edges=z[tag + "/chan_edges"], clut=z[tag + "/chan_lut"])
    _check(q, "C2 SC-LUT N=128 A=32" if kind == "SCLUTDecoder" else "C3 SCL-LUT N=128 A=32 L=8", kind, kw, x, msg, ebn0)


def _n1024_tables():
    z = np.load(os.path.join(ROOT, "quantized_decoder_polar_codes_b200", "data", "mindistortion_n1024_q16_3dB.npz"))
    f = [z["lut_f"][p].astype(np.int32)[None] for p in range(1023)]
    g = [z["lut_g"][p].astype(np.int32)[None] for p in range(1023)]
    return z, f, g


@pytest.mark.parametrize("ebn0", [1, 2, 3, 4] if FULL else [2])
def test_north_star_scl_lut_n1024(q, ebn0):
    N, K = 1024, 512
    z, f, g = _n1024_tables()
    fm, mm = sim.frozen_mask(N, K)
    kw = dict(N=N, K=K, L=8, frozen_bits=fm, message_bits=mm, LUT_f=f, LUT_g=g, virtual_channel_llr=z["llr_quanta"])
    x, msg = channel_frames(N, K, K, False, ebn0, 20_000 if FULL else 1_200, seed=300 + ebn0,
                            edges=z[f"chan_A512_eb{ebn0}/edges"], clut=z[f"chan_A512_eb{ebn0}/lut"])
    _check(q, "NS SCL-LUT N=1024 A=512 L=8", "SCLLUTDecoder", kw, x, msg, ebn0)


def mmi_channel_frames(N, K, A, ebn0_db, frames, seed, edges, Q=16):
    """probability-domain driver (mainQuantizedDecoder_ProbabilityDomain.py:153-177): y = BPSK + noise is cut at the MMI channel
    quantizer's edges interval_x[channel_lut] with utils.continous2discret (<= first edge -> 0, >= last -> Q-1, else
    bisect_left - 1); the symbols are handed over as float64."""
    rng = np.random.default_rng(seed)
    fm, _ = sim.frozen_mask(N, K)
    msg = rng.integers(0, 2, (frames, A), dtype=np.uint8)
    cw = sim.polar_encode(sim.crc_attach(msg) if K > A else msg, fm)
    sigma = sim.awgn_sigma(ebn0_db, A / N)
    y = (1.0 - 2.0 * cw) + rng.normal(0.0, sigma, cw.shape)
    idx = np.searchsorted(edges, y, side="left") - 1
    sym = np.where(y <= edges[0], 0, np.where(y >= edges[-1], Q - 1, idx))
    return sym.astype(np.float64), msg


def _n1024_mmi_tables():
    z = np.load(os.path.join(ROOT, "quantized_decoder_polar_codes_b200", "data", "mmi_n1024_q16_3dB.npz"))
    f = [z["lut_f"][p].astype(np.int32)[None] for p in range(1023)]
    g = [z["lut_g"][p].astype(np.int32)[None] for p in range(1023)]
    return z, f, g


@pytest.mark.parametrize("ebn0", [1, 2, 3] if FULL else [2])
def test_c4_ca_fast_scl_lut_n1024_mmi_tables(q, ebn0):
    """C4 as BASELINE.json names it: MMI tables (probability domain: virtual_channel_llr has n levels, symbols arrive as
    float64), CAFastSCLLUTDecoder N=1024, A=512 + CRC-24 (K=536), L=8.  The tables come from this package's MMI generator
    (tools/make_mmi_n1024.py; bit-identical to the reference generator on the golden sizes, tests/test_mmi_lutgen.py)."""
    N, A, K = 1024, 512, 536
    z, f, g = _n1024_mmi_tables()
    fm, mm = sim.frozen_mask(N, K)
    kw = dict(N=N, K=K, A=A, L=8, frozen_bits=fm, message_bits=mm, node_type=sim.identify_nodes(N, fm), LUT_f=f, LUT_g=g,
              virtual_channel_llr=z["llrs"])
    x, msg = mmi_channel_frames(N, K, A, ebn0, 20_000 if FULL else 1_200, seed=400 + ebn0, edges=z[f"chan_A512_eb{ebn0}/edges"])
    _check(q, "C4 MMI CAFastSCL-LUT N=1024 A=512 K=536 L=8", "CAFastSCLLUTDecoder", kw, x, msg, ebn0)


@pytest.mark.parametrize("kind", ["FastSCLLUTDecoder", "SCLLUTDecoder"])
def test_mmi_tables_n1024_other_list_decoders(q, kind):
    """the probability-domain driver's own list decoders (it has no CRC-aided Fast class) on the same tables"""
    N, K = 1024, 512
    z, f, g = _n1024_mmi_tables()
    fm, mm = sim.frozen_mask(N, K)
    kw = dict(N=N, K=K, L=8, frozen_bits=fm, message_bits=mm, LUT_f=f, LUT_g=g, virtual_channel_llr=z["llrs"])
    if "Fast" in kind:
        kw["node_type"] = sim.identify_nodes(N, fm)
    x, msg = mmi_channel_frames(N, K, K, 2, 2_000 if FULL else 600, seed=450, edges=z["chan_A512_eb2/edges"])
    _check(q, "MMI " + kind + " N=1024 A=512 L=8", kind, kw, x, msg, 2)


@pytest.mark.parametrize("ebn0", [1, 2, 3] if FULL else [2])
def test_c4_ca_fast_scl_lut_n1024(q, ebn0):
    """the same class on the LLR-domain (MinDistortion) tables of the north-star benchmark"""
    N, A, K = 1024, 512, 536
    z, f, g = _n1024_tables()
    fm, mm = sim.frozen_mask(N, K)
    kw = dict(N=N, K=K, A=A, L=8, frozen_bits=fm, message_bits=mm, node_type=sim.identify_nodes(N, fm), LUT_f=f, LUT_g=g,
              virtual_channel_llr=z["llr_quanta"])
    x, msg = channel_frames(N, K, A, True, ebn0, 20_000 if FULL else 1_200, seed=400 + ebn0,
                            edges=z[f"chan_A512_eb{ebn0}/edges"], clut=z[f"chan_A512_eb{ebn0}/lut"])
    _check(q, "C4 CAFastSCL-LUT N=1024 A=512 K=536 L=8", "CAFastSCLLUTDecoder", kw, x, msg, ebn0)


def c5_case(frames, ebn0=2.0, seed=500):
    """C5 the way the continuous-domain driver sets it up (mainQuantizedDecoder_ContinuousDomain.py:95-104,186-192): frozen set
    from PolarCodeConstructor.GA(sigma) -- the NR table stops at N=1024 --, step sizes from
    LLRLSUniformQuantizer.generate_uniform_quantizers(sigma), channel LLRs quantized by QUniform (:29-32) with the root
    step decoder_r_f[0]."""
    N, K, v = 2048, 1024, 16
    sigma = sim.awgn_sigma(ebn0, K / N)
    fm, mm = sim.frozen_mask_ga(N, K, sigma)
    r_f, r_g = sim.uniform_quantizer_steps(N, v, sigma)
    rng = np.random.default_rng(seed)
    msg = rng.integers(0, 2, (frames, K), dtype=np.uint8)
    llr = sim.awgn_llr(sim.polar_encode(msg, fm), sigma, rng)
    r = r_f[0]
    M = (v // 2 - 0.5) * r
    x = np.where(np.abs(llr) > M, np.sign(llr) * (M - 0.5 * r), (np.floor(llr / r) + 0.5) * r)
    kw = dict(N=N, K=K, L=32, frozen_bits=fm, message_bits=mm, decoder_r_f=r_f, decoder_r_g=r_g, v=v)
    return kw, x, msg


def test_c5_scl_uniform_n2048_l32(q):
    """C5: SCLUniformQuantizedDecoder N=2048, A=K=1024, L=32, v=16 with the reference's own construction and step sizes."""
    kw, x, msg = c5_case(256 if FULL else 32)
    _check(q, "C5 SCL-Uniform N=2048 A=1024 L=32 v=16 (GA construction, optimal uniform steps)", "SCLUniformQuantizedDecoder", kw, x, msg, 2.0)
