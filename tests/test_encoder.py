"""Encoder side: PolarEnc / CRCEnc of the PolarBDEnc package the drivers import.
CPU: the oracle's encoder restatement is pinned against the REFERENCE's own code -- its decoders invert it on a
noiseless channel, and its CA decoder (which recomputes the CRC with CRC::encoding) accepts the attached CRC.
GPU: the CUDA encoder is bit-exact against the oracle."""
import numpy as np
import pytest

import common
from oracle import polar_oracle as po
from quantized_decoder_polar_codes_b200 import simulation as sim

CRC_P = list(sim.CRC24_LOC)


def _msgbits(N, K):
    fm, mm = sim.frozen_mask(N, K)
    return fm, mm, np.where(fm == 0)[0]


@pytest.mark.parametrize("N,K", [(32, 16), (128, 64), (1024, 512)])
def test_oracle_encoder_is_inverted_by_the_reference_decoder(refmod, N, K):
    fm, mm, pos = _msgbits(N, K)
    rng = np.random.default_rng(N)
    msg = rng.integers(0, 2, (20, K), dtype=np.uint8)
    x = po.polar_encode(msg, pos, N)
    llr = (1.0 - 2.0 * x.astype(np.float64)) * 4.0           # noiseless BPSK
    got = common.ref_decode(refmod, "SCDecoder", dict(N=N, K=K, frozen_bits=fm, message_bits=mm), llr)
    assert (got == msg).all()
    assert (x == sim.polar_encode(msg, fm)).all()            # the numpy helper the fixtures were made with


def test_oracle_crc_is_accepted_by_the_reference_ca_decoder(refmod):
    """CASCLDecoder re-computes CRC::encoding over the first A decoded bits and only then prefers a path: flip the
    least reliable positions so that the CRC has to pick the right candidate."""
    N, A = 128, 40
    K = A + 24
    fm, mm, pos = _msgbits(N, K)
    rng = np.random.default_rng(7)
    msg = rng.integers(0, 2, (30, A), dtype=np.uint8)
    word = po.crc_attach(msg, 24, CRC_P)
    assert word.shape == (30, K) and (word[:, :A] == msg).all()
    assert (word == sim.crc_attach(msg)).all()
    x = po.polar_encode(word, pos, N)
    llr = (1.0 - 2.0 * x.astype(np.float64)) * 2.0 + rng.standard_normal(x.shape) * 0.9
    kw = dict(N=N, K=K, A=A, L=8, frozen_bits=fm, message_bits=mm, crc_n=24, crc_p=CRC_P)
    got = common.ref_decode(refmod, "CASCLDecoder", kw, llr)
    assert (got == msg).all(axis=1).mean() > 0.9             # the CRC-aided list decoder recovers (nearly) all frames


def test_crc_is_linear_and_zero_for_zero():
    rng = np.random.default_rng(3)
    a = rng.integers(0, 2, (8, 57), dtype=np.uint8)
    b = rng.integers(0, 2, (8, 57), dtype=np.uint8)
    ca, cb, cab = (po.crc_attach(v, 24, CRC_P)[:, 57:] for v in (a, b, a ^ b))
    assert (ca ^ cb == cab).all()
    assert not po.crc_attach(np.zeros((1, 57), np.uint8), 24, CRC_P).any()


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("N,K,B", [(32, 7, 5), (64, 33, 100), (128, 64, 1000), (512, 300, 333), (1024, 512, 4096), (2048, 1111, 64), (4096, 2048, 17)])
def test_cuda_polar_encoder_matches_oracle(N, K, B):
    from quantized_decoder_polar_codes_b200.encoder import PolarEnc
    fm, mm, pos = _msgbits(N, K) if N <= 1024 else (None, None, np.sort(np.random.default_rng(1).choice(N, K, replace=False)))
    enc = PolarEnc(N, K, np.setdiff1d(np.arange(N), pos), pos)
    msg = np.random.default_rng(B).integers(0, 2, (B, K), dtype=np.uint8)
    got = enc.encode(msg)
    assert got.dtype == np.uint8 and got.shape == (B, N)
    assert (got == po.polar_encode(msg, pos, N)).all()
    one = enc.encode(msg[0])                                  # the drivers' call: one frame, 1-D
    assert one.shape == (N,) and (one == got[0]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("A,crc_n,crc_p", [(32, 24, CRC_P), (512, 24, CRC_P), (57, 24, CRC_P), (100, 16, [16, 12, 5, 0]), (10, 8, [8, 2, 1, 0]), (64, 32, [32, 26, 23, 22, 16, 12, 11, 10, 8, 7, 5, 4, 2, 1, 0])])
def test_cuda_crc_matches_oracle(A, crc_n, crc_p):
    from quantized_decoder_polar_codes_b200.encoder import CRCEnc
    enc = CRCEnc(crc_n, crc_p)
    msg = np.random.default_rng(A).integers(0, 2, (257, A), dtype=np.uint8)
    got = enc.encode(msg)
    assert (got == po.crc_attach(msg, crc_n, crc_p)).all()
    assert (enc.encode(msg[3]) == got[3]).all()


@pytest.mark.gpu
def test_reference_import_paths_and_driver_loop():
    """The frame loop of mainFPDecoder.py:95-113 with the reference's import lines, on this build."""
    import quantized_decoder_polar_codes_b200 as q
    q.install_reference_import_paths()
    from PolarBDEnc.Encoder.CRCEnc import CRCEnc
    from PolarBDEnc.Encoder.PolarEnc import PolarEnc
    from PolarDecoder.Decoder.CASCLDecoder import CASCLDecoder
    N, A, crc_n = 256, 100, 24
    K = A + crc_n
    fm, mm, pos = _msgbits(N, K)
    polar_encoder = PolarEnc(N, K, np.where(fm == 1)[0], pos)
    crc_encoder = CRCEnc(crc_n, CRC_P)
    dec = CASCLDecoder(N, K, A, 8, fm, mm, crc_n, CRC_P)
    rng = np.random.default_rng(0)
    sigma = sim.awgn_sigma(3.0, A / N)
    ok = 0
    for _ in range(40):
        msg = rng.integers(0, 2, A)
        cword = polar_encoder.encode(crc_encoder.encode(msg)).astype(int)
        rx = (1 - 2 * cword) + rng.normal(0, sigma, (1, N))
        ok += int((dec.decode(rx * (2 / sigma ** 2)) == msg).all())
    assert ok >= 36


@pytest.mark.gpu
def test_fused_crc_polar_and_device_entry():
    import ctypes as C
    import torch
    from quantized_decoder_polar_codes_b200 import capi
    from quantized_decoder_polar_codes_b200.encoder import PD_ENC_CRC_POLAR, _Handle
    N, A = 1024, 512
    K = A + 24
    fm, mm, pos = _msgbits(N, K)
    h = _Handle(N, K, A, fm, 24, CRC_P)
    msg = np.random.default_rng(5).integers(0, 2, (3000, A), dtype=np.uint8)
    want = po.polar_encode(po.crc_attach(msg, 24, CRC_P), pos, N)
    assert (h.run(PD_ENC_CRC_POLAR, msg, A, N) == want).all()
    d_in = torch.from_numpy(msg).cuda()
    d_out = torch.empty((3000, N), dtype=torch.uint8, device="cuda")
    h.run_device(PD_ENC_CRC_POLAR, d_in.data_ptr(), 3000, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (d_out.cpu().numpy() == want).all()
    # unaligned device buffers take the byte path
    buf = torch.empty(3000 * N + 1, dtype=torch.uint8, device="cuda")
    h.run_device(PD_ENC_CRC_POLAR, d_in.data_ptr(), 3000, buf.data_ptr() + 1, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (buf[1:].view(3000, N).cpu().numpy() == want).all()
    with pytest.raises(ValueError):
        h.run(7, msg, A, N)
