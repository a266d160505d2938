"""Batched Monte-Carlo front-end on the GPU: what the reference drivers' frame loop does
(mainQuantizedDecoder_LLRDomain.py:130-192) -- random message, [CRC], polar encode, BPSK + AWGN, LLR, channel
quantizer, decode, BER/BLER counters -- without ever leaving the device.  One process per GPU; with
torch.distributed initialised the two counters are all-reduced (NCCL) at the end of each Eb/N0 point.

    sim = Simulator(decoder, frozen_bits, A=512, crc=False, channel_quantizer=(edges, lut, 16))
    res = sim.run(ebn0_db=2.0, frames=1 << 22)        # {'ber':..., 'bler':..., 'frames':..., ...}
"""
import ctypes as C

import numpy as np
import torch

from . import capi
from . import distributed as D
from .simulation import CRC24_LOC, awgn_sigma


class Simulator:
    def __init__(self, decoder, frozen_bits, A, crc=False, channel_quantizer=None, device=None, crc_n=24, crc_loc=CRC24_LOC):
        self.dec = decoder
        self.lib = capi.lib()
        self.N = self.lib.pd_code_len(decoder._handle)
        self.kout = self.lib.pd_out_len(decoder._handle)
        self.A = int(A)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        fb = np.ascontiguousarray(frozen_bits, dtype=np.int32)
        K = int((fb == 0).sum())
        cfg = capi.SimConfig()
        cfg.N, cfg.K, cfg.A, cfg.device = self.N, K, self.A, self.device.index
        cfg.frozen_bits = fb.ctypes.data
        loc = np.ascontiguousarray(crc_loc, dtype=np.int32)
        if crc:
            cfg.crc_n, cfg.crc_loc, cfg.crc_loc_len = crc_n, loc.ctypes.data, loc.size
        self.quantized = channel_quantizer is not None
        if self.quantized:
            edges, lut, qc = channel_quantizer
            edges = np.ascontiguousarray(edges, dtype=np.float64)
            lut = np.ascontiguousarray(lut, dtype=np.uint8)
            assert lut.size == edges.size - 1
            cfg.edges, cfg.n_edges, cfg.chan_lut, cfg.q_channel = edges.ctypes.data, edges.size, lut.ctypes.data, int(qc)
        h = C.c_void_p()
        capi.check(self.lib.pd_sim_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.rate = self.A / self.N

    def __del__(self):
        try:
            self.lib.pd_sim_destroy(self._h)
        except Exception:
            pass

    def generate(self, sigma, frames, seed=0, first_frame=0):
        """-> (msg [B,A] uint8, channel output [B,N] uint8 symbols or fp64 LLRs) as device tensors."""
        msg = torch.empty((frames, self.A), dtype=torch.uint8, device=self.device)
        out = torch.empty((frames, self.N), dtype=torch.uint8 if self.quantized else torch.float64, device=self.device)
        s = torch.cuda.current_stream(self.device).cuda_stream
        capi.check(self.lib.pd_sim_generate(self._h, float(sigma), frames, seed, first_frame, msg.data_ptr(), out.data_ptr(), s))
        return msg, out

    def run(self, ebn0_db, frames, batch=1 << 16, seed=0, max_block_errors=None, rank=0, world=1):
        """Decodes `frames` frames (this rank's contiguous share of them) at one Eb/N0 point.  `max_block_errors`
        reproduces the drivers' early stop (> 1000 block errors, :184), checked between batches."""
        if self.kout != self.A:
            raise ValueError(f"the decoder returns {self.kout} bits per frame but the messages have A={self.A}: with a CRC use a "
                             "CRC-aided class (it returns the A message bits); pd_count_errors compares [B][A] rows")
        sigma = awgn_sigma(ebn0_db, self.rate)
        lo, hi = D.shard_range(frames, rank, world)
        s = torch.cuda.current_stream(self.device).cuda_stream
        counters = torch.zeros(3, dtype=torch.int64, device=self.device)   # bit errors, block errors, frames
        dec_out = torch.empty((min(batch, max(hi - lo, 1)), self.kout), dtype=torch.uint8, device=self.device)
        dt = capi.PD_U8 if self.quantized else capi.PD_F64
        f0 = lo
        while f0 < hi:
            nb = min(batch, hi - f0)
            msg, x = self.generate(sigma, nb, seed, f0)
            capi.decode_device(self.dec, x.data_ptr(), dt, nb, dec_out.data_ptr(), s)
            capi.check(self.lib.pd_count_errors(dec_out.data_ptr(), msg.data_ptr(), nb, self.A, counters.data_ptr(), s))
            counters[2] += nb
            f0 += nb
            if max_block_errors is not None and int(counters[1].item()) > max_block_errors:
                break
        capi.sync_check(self.dec, s)
        D.allreduce_counters(counters)
        be, ble, n = (int(v) for v in counters.cpu().tolist())
        return {"ebn0_db": ebn0_db, "frames": n, "bit_errors": be, "block_errors": ble,
                "ber": be / max(1, n * self.A), "bler": ble / max(1, n)}
