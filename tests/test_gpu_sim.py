"""GPU tests of the on-device simulation front-end (pd_sim_*): exact encoder / CRC / quantizer-rule checks against
numpy, noise statistics, seed determinism across batch splits, and an end-to-end BLER on the real MinDistortion
tables that must agree with the reference decoder's BLER on numpy-generated frames."""
from bisect import bisect_left

import numpy as np
import pytest

import common
import real_lut
from quantized_decoder_polar_codes_b200 import simulation as sim

pytestmark = pytest.mark.gpu


def _mk(q, N, K, A, crc, quant):
    from quantized_decoder_polar_codes_b200.simulate import Simulator
    fm, mm = sim.frozen_mask(N, K)
    dec = q.SCDecoder(N=N, K=K, frozen_bits=fm, message_bits=mm)
    return Simulator(dec, fm, A=A, crc=crc, channel_quantizer=quant), fm


@pytest.fixture(scope="module")
def q():
    import quantized_decoder_polar_codes_b200 as q
    return q


@pytest.mark.parametrize("N,K,A,crc", [(128, 32, 32, False), (128, 56, 32, True), (1024, 536, 512, True), (32, 16, 16, False), (4096, 2048, 2048, False)])
def test_generator_encodes_like_numpy(q, N, K, A, crc):
    s, fm = _mk(q, N, K, A, crc, None)
    msg, llr = s.generate(sigma=1e-6, frames=300, seed=5)
    msg, llr = msg.cpu().numpy(), llr.cpu().numpy()
    word = sim.crc_attach(msg) if crc else msg
    cw = sim.polar_encode(word, fm)
    assert ((llr < 0).astype(np.uint8) == cw).all()
    assert 0.4 < msg.mean() < 0.6


def test_generator_noise_and_quantizer_rule(q):
    N, K = 1024, 512
    sigma = sim.awgn_sigma(2.0, 0.5)
    s_f, fm = _mk(q, N, K, K, False, None)
    edges = np.linspace(-12.3, 12.3, 129)
    lut = np.minimum(np.arange(128) // 8, 15).astype(np.uint8)
    s_q, _ = _mk(q, N, K, K, False, (edges, lut, 16))
    msg1, llr = s_f.generate(sigma, 2000, seed=9, first_frame=100)
    msg2, sym = s_q.generate(sigma, 2000, seed=9, first_frame=100)
    assert (msg1 == msg2).all()
    llr, sym, msg = llr.cpu().numpy(), sym.cpu().numpy(), msg1.cpu().numpy()
    # the driver's rule, mainQuantizedDecoder_LLRDomain.py:167-176
    want = np.empty_like(sym)
    xs = list(edges[:-1])
    flat = llr.ravel()
    w = want.ravel()
    for i in range(0, flat.size, 97):   # a strided sample keeps the pure-Python check fast
        v = flat[i]
        w[i] = 0 if v <= edges[0] else 15 if v >= edges[-1] else lut[bisect_left(xs, v) - 1]
        assert w[i] == sym.ravel()[i]
    idx = np.clip(np.searchsorted(edges[:-1], llr, side="left") - 1, 0, 127)
    vec = np.where(llr <= edges[0], 0, np.where(llr >= edges[-1], 15, lut[idx]))
    assert (vec == sym).all()
    # noise: y - (1-2x) ~ N(0, sigma^2)
    cw = sim.polar_encode(msg, fm)
    nz = llr * sigma ** 2 / 2 - (1.0 - 2.0 * cw)
    assert abs(nz.mean()) < 4 * sigma / np.sqrt(nz.size)
    assert abs(nz.std() / sigma - 1) < 0.01
    assert abs(np.mean(nz ** 4) / sigma ** 4 - 3) < 0.05
    # same frames whatever the split
    _, a = s_q.generate(sigma, 500, seed=9, first_frame=100)
    _, b = s_q.generate(sigma, 1500, seed=9, first_frame=600)
    assert (a.cpu().numpy() == sym[:500]).all() and (b.cpu().numpy() == sym[500:]).all()


def test_end_to_end_bler_on_real_luts(q):
    """SCL-LUT L=8, N=128 A=32, real MinDistortion tables at 3 dB: the fixture's 400 reference-decoded frames
    give BLER 0.013; 2e5 GPU-generated frames through the same channel quantizer must land in the same place."""
    from quantized_decoder_polar_codes_b200.simulate import Simulator
    z = real_lut.load()
    kw, x, want, msg = real_lut.build_kwargs(z, "A32_eb3", "SCLLUTDecoder")
    dec = q.SCLLUTDecoder(**kw)
    edges, lut = z["A32_eb3/chan_edges"], z["A32_eb3/chan_lut"]   # the driver's MinDistortion channel quantizer at 3 dB
    ref_bler = (want != msg).any(axis=1).mean()
    s = Simulator(dec, kw["frozen_bits"], A=32, channel_quantizer=(edges, lut, 16))
    res = s.run(3.0, 200000, batch=50000, seed=1)
    assert res["frames"] == 200000
    assert res["ber"] < res["bler"]
    # reference decoder on 400 numpy-generated frames vs 2e5 GPU-generated frames: same BLER within the binomial noise of the 400
    band = 5 * np.sqrt(max(res["bler"], 1e-3) * (1 - res["bler"]) / want.shape[0])
    assert abs(res["bler"] - ref_bler) < band, (res["bler"], ref_bler)
    # determinism + early stop
    res2 = s.run(3.0, 200000, batch=50000, seed=1)
    assert res2 == res
    res3 = s.run(0.0, 200000, batch=20000, seed=1, max_block_errors=1000)
    assert res3["frames"] < 200000 and res3["block_errors"] > 1000
