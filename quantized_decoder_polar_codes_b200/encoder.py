"""Batched polar encoder + CRC on the GPU behind the two classes every reference driver imports
from the (missing) PolarBDEnc package -- `PolarEnc(N, K, frozenbits, msgbits).encode(msg)` and
`CRCEnc(crc_n, crc_p).encode(msg)` (mainFPDecoder.py:12-13,56-57,102-105).  Conventions: word placed on the
ascending `msgbits`, x = u F^{(x)n} in natural order; CRC = msg || CRC::encoding(msg) (PD/src/utils.cpp:77-93).

`encode` keeps the drivers' one-frame call -- (K,) -> (N,) -- and also takes a batch, (B,K) -> (B,N); bits are uint8.
The work is done by csrc/pb_enc.cuh through pd_sim_encode (include/polar_b200.h); there is no CPU fallback."""
import ctypes as C

import numpy as np

from . import capi

PD_ENC_POLAR, PD_ENC_CRC, PD_ENC_CRC_POLAR = 0, 1, 2


class _Handle:
    def __init__(self, N, K, A, frozen_indicator, crc_n=0, crc_p=None, device=0):
        self.lib = capi.lib()
        fb = np.ascontiguousarray(frozen_indicator, dtype=np.int32)
        cfg = capi.SimConfig()
        cfg.N, cfg.K, cfg.A, cfg.device = int(N), int(K), int(A), int(device)
        cfg.frozen_bits = fb.ctypes.data
        if crc_n:
            loc = np.ascontiguousarray(crc_p, dtype=np.int32)
            cfg.crc_n, cfg.crc_loc, cfg.crc_loc_len = int(crc_n), loc.ctypes.data, loc.size
        h = C.c_void_p()
        capi.check(self.lib.pd_sim_create(C.byref(cfg), C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            self.lib.pd_sim_destroy(self.h)
        except Exception:
            pass

    def run(self, mode, bits, in_len, out_len):
        x = np.asarray(bits)
        single = x.ndim == 1
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.uint8)
        if x.shape[1] != in_len:
            raise ValueError(f"expected {in_len} bits per frame, got {x.shape[1]}")
        out = np.empty((x.shape[0], out_len), np.uint8)
        capi.check(self.lib.pd_sim_encode(self.h, mode, x.ctypes.data, x.shape[0], out.ctypes.data))
        return out[0] if single else out

    def run_device(self, mode, dev_in_ptr, B, dev_out_ptr, stream=0):
        capi.check(self.lib.pd_sim_encode_device(self.h, mode, dev_in_ptr, B, dev_out_ptr, stream))


class PolarEnc:
    """PolarEnc(N, K, frozenbits, msgbits): `frozenbits` / `msgbits` are the index arrays PolarCodeConstructor.PW()
    returns (PolarCodesUtils/CodeConstruction.py:71-84), msgbits ascending."""

    def __init__(self, N, K, frozenbits, msgbits, device=0):
        msgbits = np.asarray(msgbits, dtype=np.int64)
        if msgbits.size != K or np.any(np.diff(msgbits) <= 0):
            raise ValueError("msgbits must hold K ascending positions")
        if N < 32 or N & (N - 1):
            raise ValueError("N must be a power of two >= 32")
        ind = np.ones(N, np.int32)
        ind[msgbits] = 0
        self.N, self.K = int(N), int(K)
        self._h = _Handle(N, K, K, ind, device=device)

    def encode(self, msg):
        return self._h.run(PD_ENC_POLAR, msg, self.K, self.N)

    def encode_device(self, dev_in_ptr, B, dev_out_ptr, stream=0):
        """[B][K] uint8 device bits -> [B][N] uint8 device code bits, asynchronous on `stream`."""
        self._h.run_device(PD_ENC_POLAR, dev_in_ptr, B, dev_out_ptr, stream)


class CRCEnc:
    """CRCEnc(crc_n, crc_p): crc_p = the generator's exponents, e.g. [24, 23, 21, 20, 17, 15, 13, 12, 8, 4, 2, 1, 0]
    (mainFPDecoder.py:33-34); crc_n <= 32.  encode(msg) -> msg || crc."""

    def __init__(self, crc_n, crc_p, device=0):
        if not 1 <= int(crc_n) <= 32:
            raise ValueError("crc_n must be in [1,32]")
        self.crc_n, self.crc_p, self.device = int(crc_n), [int(v) for v in crc_p], device
        self._by_len = {}

    def _handle(self, A):
        h = self._by_len.get(A)
        if h is None:   # the C ABI wants a code around the word: the smallest one that holds A + crc_n bits
            K = A + self.crc_n
            N = 32
            while N < K:
                N *= 2
            ind = np.ones(N, np.int32)
            ind[:K] = 0
            h = self._by_len[A] = _Handle(N, K, A, ind, self.crc_n, self.crc_p, self.device)
        return h

    def encode(self, msg):
        A = int(np.asarray(msg).shape[-1])
        return self._handle(A).run(PD_ENC_CRC, msg, A, A + self.crc_n)
