"""Blind-detection helpers of the reference's PolarEncoder/PolarBD package: DMetricCalculator
(Fast-SSC walk that returns the D-metric) and the CA-SCL decoder with an RNTI-scrambled CRC that returns
(bits, PM, isPass).  CPU: oracle restatement vs the compiled reference module (oracle/_ref/libPolarBD).  GPU: CUDA vs
the oracle, on both schedule interpreters."""
import numpy as np
import pytest

from oracle import polar_oracle as po
from quantized_decoder_polar_codes_b200 import simulation as sim

CRC_P = list(sim.CRC24_LOC)


@pytest.fixture(scope="module")
def bdref():
    m = po.load_bd_reference()
    if m is None:
        pytest.skip("oracle/_ref/libPolarBD (compiled reference) not present")
    return m


def cascl_case(N, A, L, rnti_len, ebn0, B, seed, grid=None, crc_n=24, crc_p=CRC_P):
    rng = np.random.default_rng(seed)
    K = A + crc_n
    fm, mm = sim.frozen_mask(N, K)
    pos = np.where(fm == 0)[0]
    rnti = rng.integers(0, 2, rnti_len).astype(np.int32)
    msg = rng.integers(0, 2, (B, A), dtype=np.uint8)
    word = po.crc_attach(msg, crc_n, crc_p)
    if rnti_len:
        word[:, K - rnti_len:] ^= rnti.astype(np.uint8)            # the transmitter scrambles the last CRC bits
    llr = sim.awgn_llr(po.polar_encode(word, pos, N), sim.awgn_sigma(ebn0, A / N), rng)
    if grid:
        llr = np.round(llr * grid) / grid                          # coarse grid: equal path metrics
    kw = dict(N=N, K=K, A=A, L=L, frozen_bits=fm, message_bits=mm, crc_n=crc_n, crc_p=crc_p)
    return kw, llr, rnti, msg


def dmetric_case(N, K, B, seed, grid=None, ebn0=1.0):
    rng = np.random.default_rng(seed)
    fm, mm = sim.frozen_mask(N, K)
    nt = sim.identify_nodes(N, fm)
    msg = rng.integers(0, 2, (B, K), dtype=np.uint8)
    llr = sim.awgn_llr(sim.polar_encode(msg, fm), sim.awgn_sigma(ebn0, K / N), rng)
    llr[B // 2:] = rng.standard_normal((B - B // 2, N)) * 3        # half of the candidates carry no codeword
    if grid:
        llr = np.round(llr * grid) / grid
    return dict(N=N, K=K, frozen_bits=fm, message_bits=mm, node_type=nt), llr


CASCL_CASES = [(32, 4, 1, 0), (64, 16, 2, 16), (128, 40, 8, 16), (128, 40, 8, 24), (256, 100, 4, 7), (128, 30, 16, 16), (64, 10, 32, 3)]


@pytest.mark.parametrize("N,A,L,rl", CASCL_CASES)
@pytest.mark.parametrize("grid", [None, 2])
def test_oracle_cascl_rnti_vs_compiled_reference(bdref, N, A, L, rl, grid):
    kw, llr, rnti, _ = cascl_case(N, A, L, rl, 1.0, 24, seed=N + L, grid=grid)
    ref = bdref.CASCLDecoder(kw["N"], kw["K"], A, L, kw["frozen_bits"].tolist(), kw["message_bits"].tolist(), 24, CRC_P)
    bits, pm, ok = po.OracleDecoder("BDCASCLDecoder", **kw).decode_bd(llr, rnti)
    for i in range(llr.shape[0]):
        rb, rpm, rok = ref.decode(llr[i:i + 1], rnti)
        assert (np.asarray(rb) == bits[i]).all() and rpm == pm[i] and bool(rok) == bool(ok[i])
    wrong = (np.arange(rl) % 2).astype(np.int32) ^ rnti[:rl] ^ 1 if rl else rnti
    if rl:                                                        # another user's RNTI: nothing may pass by construction ...
        assert po.OracleDecoder("BDCASCLDecoder", **kw).decode_bd(llr, wrong)[2].mean() <= ok.mean()


@pytest.mark.parametrize("N,K", [(32, 16), (64, 20), (128, 64), (256, 100), (1024, 512)])
@pytest.mark.parametrize("grid", [None, 2])
def test_oracle_dmetric_vs_compiled_reference(bdref, N, K, grid):
    kw, llr = dmetric_case(N, K, 16, seed=N + K, grid=grid)
    ref = bdref.DMetricCalculator(N, K, kw["frozen_bits"].tolist(), kw["message_bits"].tolist(), kw["node_type"].tolist())
    got = po.OracleDecoder("BDDMetricCalculator", **kw).decode_bd(llr)
    want = np.array([ref.calculate(llr[i:i + 1]) for i in range(llr.shape[0])])
    assert (got == want).all()
    assert got[:8].mean() > got[8:].mean()                       # the metric separates codewords from noise


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("force", ["", "1", "2"])
@pytest.mark.parametrize("N,A,L,rl", CASCL_CASES + [(1024, 512, 8, 16), (512, 100, 5, 16)])
def test_cuda_cascl_rnti_matches_oracle(N, A, L, rl, force, monkeypatch):
    import quantized_decoder_polar_codes_b200 as q
    if force:
        monkeypatch.setenv("POLAR_B200_FORCE_GENERIC", force)
    B = 64 if N >= 512 else 300
    for grid in (None, 2):
        kw, llr, rnti, msg = cascl_case(N, A, L, rl, 2.0, B, seed=7 * N + L, grid=grid)
        dec = q.BDCASCLDecoder(**kw)
        bits, pm, ok = dec.decode(llr, rnti)
        wb, wpm, wok = po.OracleDecoder("BDCASCLDecoder", **kw).decode_bd(llr, rnti)
        assert (bits == wb).all() and (pm == wpm).all() and (ok == wok).all(), dec.kernel
        one = dec.decode(llr[3], rnti)                             # the reference's call: one frame -> (bits, PM, isPass)
        assert (one[0] == wb[3]).all() and one[1] == wpm[3] and one[2] == bool(wok[3])
    if N == 1024:
        assert ok.mean() > 0.9 and (bits[ok] == msg[ok]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("force", ["", "1", "2"])
@pytest.mark.parametrize("N,K", [(32, 16), (64, 20), (128, 64), (256, 100), (1024, 512), (2048, 700)])
def test_cuda_dmetric_matches_oracle(N, K, force, monkeypatch):
    import quantized_decoder_polar_codes_b200 as q
    if force:
        monkeypatch.setenv("POLAR_B200_FORCE_GENERIC", force)
    if N > 1024:
        fm, mm = sim.frozen_mask(N, K, "pw")
        kw = dict(N=N, K=K, frozen_bits=fm, message_bits=mm, node_type=sim.identify_nodes(N, fm))
        llr = np.round(np.random.default_rng(1).standard_normal((50, N)) * 4) / 2
    else:
        kw, llr = dmetric_case(N, K, 400, seed=N, grid=2)
    dec = q.BDDMetricCalculator(**kw)
    got = dec.calculate(llr)
    want = po.OracleDecoder("BDDMetricCalculator", **kw).decode_bd(llr)
    assert got.shape == want.shape and (got == want).all(), dec.kernel
    assert dec.calculate(llr[5]) == want[5]


@pytest.mark.gpu
def test_reference_import_paths_for_polarbd():
    import quantized_decoder_polar_codes_b200 as q
    q.install_reference_import_paths()
    from PolarBD.PolarBD.CASCLWithRNTI import CASCLDecoder
    from PolarBD.PolarBD.DMetricCalculator import DMetricCalculator
    kw, llr, rnti, msg = cascl_case(128, 40, 8, 16, 4.0, 8, seed=3)
    dec = CASCLDecoder(kw["N"], kw["K"], kw["A"], kw["L"], kw["frozen_bits"], kw["message_bits"], 24, CRC_P)
    bits, pm, ok = dec.decode(llr[0:1], rnti)
    assert ok and (bits == msg[0]).all() and isinstance(pm, float)
    kw2, llr2 = dmetric_case(128, 64, 4, seed=1)
    m = DMetricCalculator(128, 64, kw2["frozen_bits"], kw2["message_bits"], kw2["node_type"]).calculate(llr2[0:1])
    assert isinstance(m, float)
    with pytest.raises(ValueError):
        dec.decode(llr[0], np.zeros(25, np.int32))                 # RNTI longer than the CRC
