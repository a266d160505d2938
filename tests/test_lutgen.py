"""Lookup-table generation: the minimum-distortion quantizer of the reference's LLR-domain generator
and the density-evolution driver around it.  Golden vectors (tests/golden/lutgen_golden.npz) were produced by the
REFERENCE's own code (make_lutgen_golden.py).  CPU: oracle restatement vs golden / vs the reference module when its
sources are present.  GPU: pd_optls_quantize and lutgen.MinDistortionLUTGenerator vs golden and oracle, bit for bit."""
import os
import sys

import numpy as np
import pytest

from oracle import polar_oracle as po
from quantized_decoder_polar_codes_b200 import simulation as sim

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lutgen_golden.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_numpy_sum_order_is_the_pairwise_routine():
    """The restatement leans on np.sum's exact summation order; if a numpy release ever changes it this fails first."""
    import ctypes as C
    rng = np.random.default_rng(0)
    for n in list(range(1, 200)) + [255, 256, 257, 511, 600, 1000]:
        x = np.ascontiguousarray(rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, n))
        assert po.lib().po_np_sum(x.ctypes.data, n) == float(np.sum(x)), n


def test_oracle_quantizer_matches_golden(gold):
    for i in range(int(gold["nq"])):
        od, oq, lut = po.optls_quantize(gold[f"q{i}/d"], gold[f"q{i}/q"], int(gold[f"q{i}/K"]))
        assert (od == gold[f"q{i}/od"]).all() and (oq == gold[f"q{i}/oq"]).all() and (lut == gold[f"q{i}/lut"]).all(), i


@pytest.mark.skipif(not os.path.isdir("/root/reference/QuantizeDensityEvolution"), reason="reference sources not present")
def test_oracle_quantizer_matches_reference_module():
    sys.path.insert(0, "/root/reference")
    from QuantizeDensityEvolution import MinDistortionQuantizer as mdq
    rng = np.random.default_rng(5)
    for M, K in [(18, 16), (50, 16), (90, 4), (140, 16), (64, 2)]:
        q = np.unique(np.round(rng.standard_normal(M) * 16) / 4)
        if q.size <= K:
            continue
        d = rng.random(q.size)
        rd, rq, rl = mdq.find_OptLS_quantizer(d.copy(), q.copy(), K)
        od, oq, ol = po.optls_quantize(d, q, K)
        assert (rd == od).all() and (rq == oq).all() and (rl == ol).all()


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_quantizer_matches_golden_and_oracle(gold):
    from quantized_decoder_polar_codes_b200.lutgen import optls_quantize_batch
    for K in (16, 4, 8, 2):
        idx = [i for i in range(int(gold["nq"])) if int(gold[f"q{i}/K"]) == K]
        if not idx:
            continue
        od, oq, luts = optls_quantize_batch([gold[f"q{i}/d"] for i in idx], [gold[f"q{i}/q"] for i in idx], K)   # mixed sizes, one launch
        for k, i in enumerate(idx):
            assert (od[k] == gold[f"q{i}/od"]).all() and (oq[k] == gold[f"q{i}/oq"]).all() and (luts[k] == gold[f"q{i}/lut"]).all(), i
    rng = np.random.default_rng(11)
    ds, qs = [], []
    for M in [17, 33, 100, 129, 257, 400, 512, 700, 1024]:
        q = np.unique(rng.standard_normal(M) * 5 if M % 2 else np.round(rng.standard_normal(2 * M) * 64) / 8)[:M]
        d = rng.random(q.size) ** 3
        ds.append(d / d.sum())
        qs.append(q)
    od, oq, luts = optls_quantize_batch(ds, qs, 16)
    for k in range(len(ds)):
        wd, wq, wl = po.optls_quantize(ds[k], qs[k], 16)
        assert (od[k] == wd).all() and (oq[k] == wq).all() and (luts[k] == wl).all(), len(ds[k])


@pytest.mark.gpu
@pytest.mark.parametrize("tag,N,v", [("n16v4", 16, 4), ("n32v8", 32, 8), ("n64v16", 64, 16)])
def test_cuda_generator_reproduces_the_reference_tables(gold, tag, N, v):
    from quantized_decoder_polar_codes_b200.lutgen import MinDistortionLUTGenerator, optls_quantize_batch
    # the channel quantizer the driver designs first (GenerateLookUpTable_LLRDomain.py:40-47) is the same routine
    cells = gold[tag + "/chan_cells"]
    assert np.all(np.diff(cells) > 0)
    od, oq, luts = optls_quantize_batch([gold[tag + "/chan_pyx"]], [cells], v)
    assert (od[0] == gold[tag + "/chan_density"]).all() and (oq[0] == gold[tag + "/chan_quanta"]).all() and (luts[0] == gold[tag + "/chan_lut"]).all()
    dens, quan, lut_f, lut_g = MinDistortionLUTGenerator(N, v).run(od[0], oq[0])
    assert (lut_f == gold[tag + "/lut_f"]).all() and (lut_g == gold[tag + "/lut_g"]).all()
    assert (quan == gold[tag + "/llr_quanta"]).all() and (dens == gold[tag + "/llr_density"]).all()


@pytest.mark.gpu
def test_cuda_generator_reproduces_the_n128_fixture(gold):
    """BASELINE configs 2/3 (N=128, Q=16, design 3 dB): the tables the reference generator took ~4 minutes to make
    (tests/golden/real_lut_n128.npz) come out of the GPU generator bit for bit."""
    import real_lut
    from quantized_decoder_polar_codes_b200.lutgen import MinDistortionLUTGenerator, optls_quantize_batch
    z = real_lut.load()
    od, oq, _ = optls_quantize_batch([gold["n128v16/chan_pyx"]], [gold["n128v16/chan_cells"]], 16)
    dens, quan, lut_f, lut_g = MinDistortionLUTGenerator(128, 16).run(od[0], oq[0])
    assert (lut_f == z["lut_f"]).all() and (lut_g == z["lut_g"]).all() and (quan == z["llr_quanta"]).all()


@pytest.mark.gpu
def test_cuda_generator_reproduces_the_n1024_benchmark_tables(gold):
    """The north-star tables (N=1024, Q=16, design 3 dB; 30-40 minutes of the reference's Python) bit for bit."""
    from quantized_decoder_polar_codes_b200.lutgen import MinDistortionLUTGenerator, optls_quantize_batch
    z = np.load(os.path.join(os.path.dirname(GOLD), "..", "..", "quantized_decoder_polar_codes_b200", "data", "mindistortion_n1024_q16_3dB.npz"))
    od, oq, _ = optls_quantize_batch([gold["n128v16/chan_pyx"]], [gold["n128v16/chan_cells"]], 16)
    dens, quan, lut_f, lut_g = MinDistortionLUTGenerator(1024, 16).run(od[0], oq[0])
    assert (lut_f == z["lut_f"]).all() and (lut_g == z["lut_g"]).all() and (quan == z["llr_quanta"]).all()


@pytest.mark.gpu
def test_generated_tables_drive_the_decoders():
    """Tables for N=256, Q=16 designed on the GPU at 3 dB, then used by the LUT decoders on AWGN frames at 3 dB: the list
    decoder must beat plain SC and both must decode most frames (an end-to-end sanity check of the whole design chain)."""
    import quantized_decoder_polar_codes_b200 as q
    from quantized_decoder_polar_codes_b200.lutgen import MinDistortionLUTGenerator, optls_quantize_batch
    N, K, v = 256, 128, 16
    sigma = np.sqrt(1 / 10 ** 0.3)
    E = 2 / sigma ** 2
    D = np.sqrt(2 * E)
    edges = np.linspace(-E - 3 * D, E + 3 * D, 129)                 # 128 uniform LLR cells over +-3 sigma around +-E
    mid = 0.5 * (edges[:-1] + edges[1:])
    pdf = 0.5 * (np.exp(-(mid - E) ** 2 / (2 * D * D)) + np.exp(-(mid + E) ** 2 / (2 * D * D)))
    od, oq, luts = optls_quantize_batch([pdf / pdf.sum()], [mid], v)
    gen = MinDistortionLUTGenerator(N, v)
    dens, quan, lut_f, lut_g = gen.run(od[0], oq[0])
    fm, mm = sim.frozen_mask(N, K)
    tabs = gen.decoder_tables(lut_f, lut_g, quan)
    rng = np.random.default_rng(3)
    msg = rng.integers(0, 2, (4000, K), dtype=np.uint8)
    llr = sim.awgn_llr(sim.polar_encode(msg, fm), sim.awgn_sigma(3.0, K / N), rng)
    cell = np.clip(np.searchsorted(edges, llr) - 1, 0, 127)
    sym = luts[0][cell].astype(np.uint8)
    sc = q.SCLUTDecoder(N, K, fm, mm, tabs["LUT_f"], tabs["LUT_g"], tabs["virtual_channel_llr"])
    scl = q.SCLLUTDecoder(N, K, 8, fm, mm, tabs["LUT_f"], tabs["LUT_g"], tabs["virtual_channel_llr"])
    bler_sc = (sc.decode(sym) != msg).any(axis=1).mean()
    bler_scl = (scl.decode(sym) != msg).any(axis=1).mean()
    assert bler_scl <= bler_sc < 0.5
