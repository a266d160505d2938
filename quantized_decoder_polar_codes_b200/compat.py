"""Compatibility layer that lets the reference's UNMODIFIED scripts -- the Monte-Carlo drivers main*Decoder_*.py and the
table generators GenerateLookUpTable_*.py -- run against this package in today's environment (SURVEY.md Appendix C lists
what stops them otherwise).  `install()` is idempotent and only adds what is missing:

  import paths   PolarDecoder.Decoder.<X> (this package's pybind classes), PolarBDEnc.Encoder.{PolarEnc,CRCEnc} (the GPU
                 encoder; the package is imported by all four drivers but absent from the reference tree), PolarBD
  quantizers     quantizers.quantizer.LLROptLSQuantizer.LLRQuantizer and quantizers.quantizer.{MMI,MMIQuantizer}.MMIQuantizer
                 -- the reference's C++/OpenCV design package (cannot be built without OpenCV) -- served by lutgen's GPU
                 quantizers with the C++ call signatures (LLRQuantizer.cpp:67,164; MMIQuantizer.cpp:73,264)
  numpy          np.int (removed in 1.24; mainFPDecoder.py:103, CodeConstruction.py:68,79-80), np.loadtxt(delimiter="\\n")
                 (CodeConstruction.py:68), ragged np.array([...]) of per-node tables (mainQuantizedDecoder_LLRDomain.py:88-89)
                 -> object arrays as numpy < 1.24 built them
  stubs          torchtracer (Tracer/Config: results are written as text files next to where the drivers put them) and
                 matplotlib.pyplot when the real packages are not installed
  frame cap      optional: tqdm(range(MaxBlock)) in the drivers' frame loops is cut to `max_frames` iterations (the
                 drivers hard-code MaxBlock = 1e5 frames per Eb/N0 point, one Python call per frame)

`python -m quantized_decoder_polar_codes_b200.run_driver <script.py> [args]` installs it and runs a script."""
import importlib
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_installed = {}


def _module(name, **attrs):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        sys.modules[name] = m
        parent, _, child = name.rpartition(".")
        if parent:
            setattr(_module(parent), child, m)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


# ---------------------------------------------------------------------------------------------------- numpy
def _patch_numpy():
    if _installed.get("numpy"):
        return
    _installed["numpy"] = True
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    if not hasattr(np, "bool"):
        np.bool = bool
    _loadtxt, _array = np.loadtxt, np.array

    def loadtxt(fname, *a, **kw):
        if kw.get("delimiter") == "\n":      # one value per line: the default whitespace splitting reads the same file
            kw = dict(kw)
            kw.pop("delimiter")
        return _loadtxt(fname, *a, **kw)

    def array(obj, *a, **kw):
        try:
            return _array(obj, *a, **kw)
        except ValueError as e:              # ragged list of per-node tables: an object array, as numpy < 1.24 made
            if "inhomogeneous" not in str(e) or a or kw:
                raise
            out = np.empty(len(obj), dtype=object)
            for i, v in enumerate(obj):
                out[i] = v
            return out

    np.loadtxt, np.array = loadtxt, array


# ---------------------------------------------------------------------------------------------------- stubs
class _Tracer:
    """torchtracer.Tracer as the drivers use it: Tracer(root).attach(name); .store(Config|figure[, file]); .log(msg, file=)"""

    def __init__(self, root="."):
        self.root, self.dir = root, root

    def attach(self, name):
        self.dir = os.path.join(self.root, name)
        os.makedirs(self.dir, exist_ok=True)
        return self

    def store(self, obj, file=None):
        if isinstance(obj, _Config):
            with open(os.path.join(self.dir, "config.json"), "w") as f:
                import json
                json.dump(obj.conf, f, default=str)
        elif file is not None and hasattr(obj, "savefig"):
            obj.savefig(os.path.join(self.dir, file))

    def log(self, msg, file="log"):
        with open(os.path.join(self.dir, file + ".log"), "a") as f:
            f.write(str(msg) + "\n")


class _Config:
    def __init__(self, conf):
        self.conf = dict(conf)


class _Figure:
    def savefig(self, path, *a, **kw):
        open(path, "wb").close()


def _install_stubs():
    try:
        importlib.import_module("torchtracer")
    except ImportError:
        _module("torchtracer", Tracer=_Tracer)
        _module("torchtracer.data", Config=_Config)
    try:
        importlib.import_module("matplotlib.pyplot")
    except ImportError:
        fig = _Figure()
        noop = lambda *a, **kw: None   # noqa: E731
        _module("matplotlib")
        _module("matplotlib.pyplot", figure=lambda *a, **kw: fig, gcf=lambda: fig, **{
            k: noop for k in ("semilogy", "plot", "legend", "xlabel", "ylabel", "grid", "show", "savefig", "title", "close", "ylim", "xlim")})


# ---------------------------------------------------------------------------------------------------- quantizers
class LLRQuantizer:
    """quantizers.quantizer.LLROptLSQuantizer.LLRQuantizer (LLRQuantizer.cpp:67-164) on the GPU quantizer of lutgen."""

    def __init__(self, device=0):
        self.device = device

    def find_OptLS_quantizer(self, density, quanta, M, K):
        from . import lutgen
        density = np.asarray(density, dtype=np.float64).ravel()
        quanta = np.asarray(quanta, dtype=np.float64).ravel()
        if density.shape[0] <= K:            # nothing to merge: identity on the sorted symbols, zero padded
            order = np.argsort(quanta, kind="stable")
            lut = np.zeros(density.shape[0], np.int32)
            lut[order] = np.arange(density.shape[0])
            d, q = np.zeros(K), np.zeros(K)
            d[:density.shape[0]], q[:density.shape[0]] = density[order], quanta[order]
            return d[None], q[None], lut[None], 0.0
        od, oq, luts = lutgen.optls_quantize_batch([density], [quanta], int(K), self.device)
        return od[0][None], oq[0][None], luts[0][None], 0.0


class MMIQuantizer:
    """quantizers.quantizer.MMI.MMIQuantizer (MMIQuantizer.cpp:73-165, 264-325) on the GPU passes of lutgen."""

    def __init__(self, px1=0.5, px_minus1=0.5, device=0):
        self.px1, self.pxm, self.device = float(px1), float(px_minus1), device

    def find_opt_quantizer(self, joint_prob, K):
        from . import lutgen
        Q, pzx, Az, perm = lutgen.mmi_quantize_batch(np.asarray(joint_prob, dtype=np.float64)[None], int(K), self.device, self.px1, self.pxm)
        return Q[0], pzx[0], Az[0], perm[0]

    def find_opt_quantizer_AWGN(self, joint_prob, K):
        from . import lutgen
        _, _, Az, _ = lutgen.mmi_quantize_batch(np.asarray(joint_prob, dtype=np.float64)[None], int(K), self.device, self.px1, self.pxm, sort=False)
        return Az[0]


def _install_quantizers():
    _module("quantizers")
    _module("quantizers.quantizer")
    _module("quantizers.quantizer.LLROptLSQuantizer", LLRQuantizer=LLRQuantizer)
    _module("quantizers.quantizer.MMI", MMIQuantizer=MMIQuantizer)            # the name the scripts import
    _module("quantizers.quantizer.MMIQuantizer", MMIQuantizer=MMIQuantizer)   # the name of the reference's file


# ---------------------------------------------------------------------------------------------------- frame cap
def _cap_tqdm(max_frames):
    import itertools
    import tqdm as _tq

    class capped:
        def __init__(self, iterable=None, *a, **kw):
            self.it = iterable

        def __iter__(self):
            return itertools.islice(iter(self.it), max_frames)

        def set_description(self, *a, **kw):
            pass

        def update(self, *a, **kw):
            pass

        def close(self):
            pass

    _tq.tqdm = capped


def install(decoders=True, encoder=True, quantizers=True, max_frames=None, seed=None):
    """decoders / encoder / quantizers = False leaves those import paths alone (a caller that provides its own)."""
    _patch_numpy()
    _install_stubs()
    if (decoders or encoder) and _HERE not in sys.path:
        sys.path.insert(0, _HERE)            # PolarDecoder/, PolarBDEnc/, PolarBD/ live next to this file
    if decoders:
        importlib.import_module("PolarDecoder")
    if encoder:
        importlib.import_module("PolarBDEnc")
    if quantizers:
        _install_quantizers()
    if max_frames:
        _cap_tqdm(int(max_frames))
    if seed is not None:
        np.random.seed(int(seed))
