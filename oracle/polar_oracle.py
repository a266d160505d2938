"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of oracle/polar_oracle.c and loader of oracle/_ref.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  The product package (quantized_decoder_polar_codes_b200) never does.

`OracleDecoder(kind, **ctor_kwargs)` takes exactly the constructor keywords of the reference pybind
classes (PolarDecoder/PolarDecoder/_cpp/py_interface/py_*.cpp) and `decode(x)` accepts (N,), (1,N) or
(B,N) inputs.  `load_reference()` imports the compiled, unmodified reference module from oracle/_ref.
"""
import ctypes as C
import importlib.util
import os
import subprocess
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

KINDS = {
    "SCDecoder": 0, "FastSCDecoder": 1, "SCLDecoder": 2, "FastSCLDecoder": 3, "CASCLDecoder": 4,
    "SCLUTDecoder": 5, "FastSCLUTDecoder": 6, "SCLLUTDecoder": 7, "FastSCLLUTDecoder": 8,
    "CASCLLUTDecoder": 9, "CAFastSCLLUTDecoder": 10,
    "SCUniformQuantizedDecoder": 11, "SCLUniformQuantizedDecoder": 12,
    "SCLloydQuantizedDecoder": 13, "SCLLloydQuantizedDecoder": 14,
    # blind-detection helpers of PolarEncoder/PolarBD (module libPolarBD: DMetricCalculator, CASCLDecoder)
    "BDDMetricCalculator": 15, "BDCASCLDecoder": 16,
}
LUT_KINDS = {5, 6, 7, 8, 9, 10}
CA_KINDS = {4, 9, 10, 16}


class _Cfg(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("A", C.c_int32), ("L", C.c_int32),
        ("frozen", C.c_void_p), ("node_type", C.c_void_p),
        ("crc_n", C.c_int32), ("crc_p", C.c_void_p),
        ("lut_pool", C.c_void_p), ("f_off", C.c_void_p), ("g_off", C.c_void_p),
        ("f_pos", C.c_void_p), ("g_pos", C.c_void_p),
        ("fqa", C.c_void_p), ("fqb", C.c_void_p), ("gqa", C.c_void_p), ("gqb", C.c_void_p),
        ("llr_pool", C.c_void_p), ("llr_off", C.c_void_p), ("llr_levels", C.c_int32),
        ("r_f", C.c_void_p), ("r_g", C.c_void_p), ("v", C.c_int32),
        ("bnd_f", C.c_void_p), ("bnd_g", C.c_void_p), ("rec_f", C.c_void_p), ("rec_g", C.c_void_p),
        ("nb", C.c_int32), ("nr", C.c_int32),
    ]


_lib = None


def build_port(force=False):
    so = os.path.join(HERE, "libpolar_oracle.so")
    src = os.path.join(HERE, "polar_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_port())
        _lib.po_decode.restype = C.c_int
        _lib.po_decode.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.po_crc_attach.restype = None
        _lib.po_crc_attach.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        _lib.po_polar_encode.restype = None
        _lib.po_polar_encode.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        _lib.po_decode_bd.restype = C.c_int
        _lib.po_decode_bd.argtypes = [C.POINTER(_Cfg), C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.po_np_sum.restype = C.c_double
        _lib.po_np_sum.argtypes = [C.c_void_p, C.c_int]
        _lib.po_optls_quantize.restype = C.c_int
        _lib.po_optls_quantize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.po_std_sort_idx.restype = None
        _lib.po_std_sort_idx.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    return _lib


def std_sort_idx(keys):
    """argsort with the tie order of libstdc++ std::sort (GCC 13)."""
    keys = np.ascontiguousarray(keys, dtype=np.float64)
    idx = np.arange(keys.size, dtype=np.int32)
    lib().po_std_sort_idx(idx.ctypes.data, keys.size, keys.ctypes.data)
    return idx


def _ptr(a):
    return a.ctypes.data if a is not None else None


def flatten_lut(N, lut, is_g):
    """nested [N-1][npos][(2)][Qa][Qb] -> (pool int32, off int64[N-1], npos[N-1], qa[N-1], qb[N-1])."""
    chunks, off, npos, qa, qb = [], [], [], [], []
    cur = 0
    for p in range(N - 1):
        t = np.asarray(lut[p], dtype=np.int32)
        want = 4 if is_g else 3
        if t.ndim == want - 1:          # a single table for the node (compact form)
            t = t[None]
        assert t.ndim == want, f"node {p}: bad LUT rank {t.ndim}"
        if t.shape[0] > 1 and np.all(t == t[:1]):
            t = t[:1]
        off.append(cur); npos.append(t.shape[0]); qa.append(t.shape[-2]); qb.append(t.shape[-1])
        chunks.append(t.ravel()); cur += t.size
    return (np.concatenate(chunks).astype(np.int32), np.asarray(off, np.int64), np.asarray(npos, np.int32),
            np.asarray(qa, np.int32), np.asarray(qb, np.int32))


def flatten_llr(N, llr):
    levels = len(llr)
    chunks, off = [], []
    cur = 0
    for lv in range(levels):
        assert len(llr[lv]) >= N
        for pos in range(N):
            row = np.asarray(llr[lv][pos], dtype=np.float64).ravel()
            off.append(cur); chunks.append(row); cur += row.size
    return np.concatenate(chunks), np.asarray(off, np.int64), levels


class OracleDecoder:
    def __init__(self, kind, N, K, frozen_bits, message_bits=None, L=1, A=0, node_type=None, crc_n=0, crc_p=None,
                 LUT_f=None, LUT_g=None, LUT_Fs=None, LUT_Gs=None, virtual_channel_llr=None,
                 decoder_r_f=None, decoder_r_g=None, v=0,
                 boundaries_f=None, boundaries_g=None, reconstruction_f=None, reconstruction_g=None):
        self.kind = KINDS[kind] if isinstance(kind, str) else int(kind)
        self.N, self.K, self.A, self.L = int(N), int(K), int(A), int(L)
        self._keep = []
        cfg = _Cfg()
        cfg.kind, cfg.N, cfg.K, cfg.A, cfg.L = self.kind, self.N, self.K, self.A, self.L

        def keep(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            self._keep.append(a)
            return a

        cfg.frozen = _ptr(keep(frozen_bits, np.int32))
        if node_type is not None:
            cfg.node_type = _ptr(keep(node_type, np.int32))
        if crc_p is not None:
            poly = np.zeros(int(crc_n) + 1, np.int32)
            poly[np.asarray(crc_p, dtype=np.int64)] = 1      # PD/src/CASCLDecoder.cpp:49-53
            cfg.crc_n = int(crc_n)
            cfg.crc_p = _ptr(keep(poly, np.int32))
        LUT_f = LUT_f if LUT_f is not None else LUT_Fs
        LUT_g = LUT_g if LUT_g is not None else LUT_Gs
        if self.kind in LUT_KINDS:
            fp, fo, fn, fa, fb = flatten_lut(self.N, LUT_f, False)
            gp, go, gn, ga, gb = flatten_lut(self.N, LUT_g, True)
            pool = keep(np.concatenate([fp, gp]), np.int32)
            cfg.lut_pool = _ptr(pool)
            cfg.f_off = _ptr(keep(fo, np.int64)); cfg.g_off = _ptr(keep(go + fp.size, np.int64))
            cfg.f_pos = _ptr(keep(fn, np.int32)); cfg.g_pos = _ptr(keep(gn, np.int32))
            cfg.fqa = _ptr(keep(fa, np.int32)); cfg.fqb = _ptr(keep(fb, np.int32))
            cfg.gqa = _ptr(keep(ga, np.int32)); cfg.gqb = _ptr(keep(gb, np.int32))
            lp, lo, lv = flatten_llr(self.N, virtual_channel_llr)
            cfg.llr_pool = _ptr(keep(lp, np.float64)); cfg.llr_off = _ptr(keep(lo, np.int64)); cfg.llr_levels = lv
        if decoder_r_f is not None:
            cfg.r_f = _ptr(keep(decoder_r_f, np.float64)); cfg.r_g = _ptr(keep(decoder_r_g, np.float64))
        cfg.v = int(v)
        if boundaries_f is not None:
            bf = keep(boundaries_f, np.float64); rf = keep(reconstruction_f, np.float64)
            cfg.bnd_f = _ptr(bf); cfg.bnd_g = _ptr(keep(boundaries_g, np.float64))
            cfg.rec_f = _ptr(rf); cfg.rec_g = _ptr(keep(reconstruction_g, np.float64))
            cfg.nb, cfg.nr = bf.shape[1], rf.shape[1]
        self.cfg = cfg
        self.kout = self.A if self.kind in CA_KINDS else self.K
        self.list = self.kind in (2, 3, 4, 7, 8, 9, 10, 12, 14, 16)

    def decode(self, x, return_pm=False):
        x = np.asarray(x)
        single = x.ndim == 1
        xb = x.reshape(-1, self.N) if x.ndim <= 1 else x
        xb = np.ascontiguousarray(xb, dtype=np.int32 if self.kind in LUT_KINDS else np.float64)
        B = xb.shape[0]
        out = np.empty((B, self.kout), np.uint8)
        Lr = self.L if self.list else 1
        pm = np.empty((B, Lr), np.float64)
        win = np.empty(B, np.int32)
        rc = lib().po_decode(C.byref(self.cfg), xb.ctypes.data, B, out.ctypes.data, pm.ctypes.data, win.ctypes.data)
        assert rc == 0
        res = out[0] if single else out
        return (res, pm, win) if return_pm else res


    def decode_bd(self, x, rnti=None):
        """BDDMetricCalculator: -> metric (B,) float64.  BDCASCLDecoder: -> (bits (B,A), PM (B,), isPass (B,) bool)."""
        xb = np.ascontiguousarray(np.atleast_2d(np.asarray(x, dtype=np.float64)))
        B = xb.shape[0]
        metric = np.empty(B, np.float64)
        if self.kind == KINDS["BDDMetricCalculator"]:
            rc = lib().po_decode_bd(C.byref(self.cfg), xb.ctypes.data, B, None, 0, None, metric.ctypes.data, None)
            assert rc == 0
            return metric
        r = np.ascontiguousarray(rnti if rnti is not None else [], dtype=np.int32)
        bits = np.empty((B, self.kout), np.uint8)
        ok = np.empty(B, np.uint8)
        rc = lib().po_decode_bd(C.byref(self.cfg), xb.ctypes.data, B, r.ctypes.data if r.size else None, r.size,
                                bits.ctypes.data, metric.ctypes.data, ok.ctypes.data)
        assert rc == 0
        return bits, metric, ok.astype(bool)


# ---------------------------------------------------------------------------------------------------
def bd_ref_path():
    return os.path.join(HERE, "_ref", "libPolarBD" + sysconfig.get_config_var("EXT_SUFFIX"))


_bd_mod = None


def load_bd_reference(reference_root="/root/reference"):
    """oracle/_ref/libPolarBD*.so = the reference's PolarEncoder/PolarBD module compiled unmodified (built on demand where
    the reference sources exist); None otherwise."""
    global _bd_mod
    if _bd_mod is None:
        p = bd_ref_path()
        if not os.path.exists(p) and os.path.isdir(reference_root):
            subprocess.check_call(["make", "-C", HERE, "ref_bd", f"REFERENCE={reference_root}"], stdout=subprocess.DEVNULL)
        if not os.path.exists(p):
            return None
        spec = importlib.util.spec_from_file_location("libPolarBD", p)
        _bd_mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_bd_mod)
    return _bd_mod


def ref_path():
    return os.path.join(HERE, "_ref", "_libPolarDecoder" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_reference(reference_root="/root/reference"):
    """Compile the unmodified reference decoders into oracle/_ref (needs the reference sources)."""
    if os.path.exists(ref_path()):
        return ref_path()
    if not os.path.isdir(reference_root):
        return None
    subprocess.check_call(["make", "-C", HERE, f"-j{os.cpu_count() or 4}", "ref", f"REFERENCE={reference_root}"],
                          stdout=subprocess.DEVNULL)
    return ref_path()


_ref_mod = None


def load_reference():
    """Import oracle/_ref/_libPolarDecoder*.so (the compiled reference); None if it has not been built."""
    global _ref_mod
    if _ref_mod is None:
        p = ref_path()
        if not os.path.exists(p):
            return None
        spec = importlib.util.spec_from_file_location("_libPolarDecoder", p)
        _ref_mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_ref_mod)
    return _ref_mod


def crc_attach(msg, crc_n, crc_p):
    """msg || CRC::encoding(msg) (PD/src/utils.cpp:77-93); crc_p = generator exponents as the drivers pass them."""
    m = np.ascontiguousarray(np.atleast_2d(msg), dtype=np.uint8)
    poly = np.zeros(int(crc_n) + 1, np.int32)
    poly[np.asarray(crc_p, dtype=np.int64)] = 1
    out = np.empty((m.shape[0], m.shape[1] + int(crc_n)), np.uint8)
    lib().po_crc_attach(m.ctypes.data, m.shape[0], m.shape[1], poly.ctypes.data, int(crc_n), out.ctypes.data)
    return out


def polar_encode(word, msg_positions, N):
    """u[msg_positions] = word, x = u F^{(x)n} (natural order, the decoders' re-encode butterfly, FastSCDecoder.cpp:153-164)."""
    w = np.ascontiguousarray(np.atleast_2d(word), dtype=np.uint8)
    pos = np.ascontiguousarray(msg_positions, dtype=np.int32)
    out = np.empty((w.shape[0], int(N)), np.uint8)
    lib().po_polar_encode(w.ctypes.data, w.shape[0], w.shape[1], pos.ctypes.data, int(N), out.ctypes.data)
    return out


def optls_quantize(density, quanta, K):
    """Minimum-distortion quantizer of the LLR-domain table generator (MinDistortionQuantizer.py:28-99) for inputs sorted by
    ascending quanta, M > K.  -> (density[K], quanta[K], lut[M] int32)."""
    d = np.ascontiguousarray(density, dtype=np.float64)
    q = np.ascontiguousarray(quanta, dtype=np.float64)
    assert d.ndim == 1 and d.shape == q.shape and np.all(np.diff(q) > 0)
    od, oq, lut = np.empty(K), np.empty(K), np.empty(d.size, np.int32)
    rc = lib().po_optls_quantize(d.ctypes.data, q.ctypes.data, d.size, int(K), od.ctypes.data, oq.ctypes.data, lut.ctypes.data)
    assert rc == 0
    return od, oq, lut
