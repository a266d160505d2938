"""CPU tests of the oracle itself: the C restatement (oracle/polar_oracle.c) against
  (a) the committed golden vectors produced by the compiled reference, and
  (b) the compiled reference directly (oracle/_ref), when it is present."""
import hashlib

import numpy as np
import pytest

import common
from golden.cases import CASES
from golden.make_golden import case_hash
from oracle import polar_oracle as po

FAST_CASES = [c for c in CASES if not c[0].startswith("C5-")]


def _run_case(cid, ckw):
    ckw = dict(ckw)
    kind = ckw.pop("kind")
    kw, x, truth = common.make_case(kind, **ckw)
    return kind, kw, x, truth


@pytest.mark.parametrize("cid,ckw", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_golden(golden, cid, ckw):
    kind, kw, x, _ = _run_case(cid, ckw)
    sha = bytes(golden[cid + "/sha"]).hex()
    assert sha == case_hash(kw, x), "case generator drifted from the committed fixture"
    shape = tuple(golden[cid + "/shape"])
    want = np.unpackbits(golden[cid + "/out"], axis=1)[:, : shape[1]]
    got = po.OracleDecoder(kind, **kw).decode(x)
    assert got.shape == shape
    bad = int((got != want).any(axis=1).sum())
    assert bad == 0, f"{bad}/{shape[0]} frames differ from the reference"


@pytest.mark.parametrize("kind", common.ALL_KINDS)
@pytest.mark.parametrize("seed", [101, 102])
def test_oracle_matches_compiled_reference(refmod, kind, seed):
    kw, x, _ = common.make_case(kind, N=128, K=48, L=8, A=24, B=80, seed=seed)
    want = common.ref_decode(refmod, kind, kw, x)
    got = po.OracleDecoder(kind, **kw).decode(x)
    assert (got == want).all()


_STD_SORT_SRC = r"""
#include <algorithm>
extern "C" void real_std_sort(int *idx, int n, const double *key) {
    for (int i = 0; i < n; ++i) idx[i] = i;
    std::sort(idx, idx + n, [key](int a, int b) { return key[a] < key[b]; });
}
"""


@pytest.fixture(scope="module")
def real_std_sort(tmp_path_factory):
    """The real libstdc++ std::sort of THIS toolchain with the reference's comparator (PD/src/SCLLUTDecoder.cpp:16)."""
    import ctypes
    import subprocess
    d = tmp_path_factory.mktemp("stdsort")
    src = d / "s.cpp"
    src.write_text(_STD_SORT_SRC)
    so = d / "s.so"
    subprocess.check_call(["g++", "-O3", "-std=c++14", "-fPIC", "-shared", str(src), "-o", str(so)])
    lib = ctypes.CDLL(str(so))
    lib.real_std_sort.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]

    def run(keys):
        keys = np.ascontiguousarray(keys, np.float64)
        idx = np.zeros(keys.size, np.int32)
        lib.real_std_sort(idx.ctypes.data, keys.size, keys.ctypes.data)
        return idx
    return run


@pytest.mark.parametrize("n", [1, 2, 5, 16, 17, 31, 32, 33, 48, 64, 100, 257, 1024])
def test_std_sort_emulation_matches_libstdcpp(real_std_sort, n):
    """Tie order of std::sort is algorithm-defined above 16 elements: the emulation must
    reproduce the real libstdc++ permutation exactly, on tie-heavy, sorted, reversed and heap-fallback-prone keys."""
    rng = np.random.default_rng(n)
    trials = []
    for _ in range(30):
        trials.append(rng.choice([0.0, 0.5, 1.0, 2.0, np.inf], size=n))
        trials.append(rng.standard_normal(n))
        trials.append(np.round(rng.standard_normal(n) * 2) / 2)
    trials.append(np.arange(n, dtype=np.float64))
    trials.append(np.arange(n, dtype=np.float64)[::-1].copy())
    trials.append(np.zeros(n))
    trials.append(np.where(np.arange(n) % 2 == 0, np.arange(n), -np.arange(n)).astype(np.float64))  # organ pipe-ish
    for keys in trials:
        idx = po.std_sort_idx(keys)
        assert (idx == real_std_sort(keys)).all()
        if n <= 16:
            assert (idx == np.argsort(keys, kind="stable")).all()


import real_lut


@pytest.mark.parametrize("tag,kind", real_lut.CASES, ids=[f"{t}-{k}" for t, k in real_lut.CASES])
def test_oracle_matches_reference_on_real_mindistortion_luts(tag, kind):
    """BASELINE.json configs 2/3: tables from the reference's own MinDistortion generator code (N=128, Q=16,
    design SNR 3 dB), channel quantizer and frame loop of mainQuantizedDecoder_LLRDomain.py; outputs of the
    compiled reference decoders are the golden vectors (tests/golden/make_real_luts.py)."""
    kw, x, want, msg = real_lut.build_kwargs(real_lut.load(), tag, kind)
    got = po.OracleDecoder(kind, **kw).decode(x.astype(np.int32))
    assert (got == want).all()
    if tag.endswith("eb3") and "crc" not in tag:
        assert (got != msg).any(axis=1).mean() < 0.06   # a working decoder: BLER at 3 dB


def test_oracle_matches_compiled_reference_on_benchmark_workload(refmod):
    """The benchmark's real N=1024 MinDistortion tables: many duplicated / zero LLR quanta, i.e. constant ties."""
    import bench
    kw, sym, msg = bench.make_workload(bench.CONFIGS["NS"], 12, seed=4)
    want = common.ref_decode(refmod, "SCLLUTDecoder", kw, sym.astype(np.int32))
    got = po.OracleDecoder("SCLLUTDecoder", **kw).decode(sym.astype(np.int32))
    assert (got == want).all()


def test_truthful_decoding_on_clean_channel():
    """Sanity of the whole chain (encoder conventions, frozen mask, CRC): high SNR => message recovered."""
    for kind in ["SCDecoder", "FastSCDecoder", "SCLDecoder", "FastSCLDecoder", "CASCLDecoder"]:
        K = 64 + (24 if kind in common.CA_KINDS else 0)
        kw, x, truth = common.make_case(kind, N=128, K=K, L=4, A=64, B=20, seed=5, tables="channel", ebn0_db=8.0)
        got = po.OracleDecoder(kind, **kw).decode(x)
        assert (got == truth).all(), kind
    for kind in ["SCLUTDecoder", "FastSCLUTDecoder", "SCLLUTDecoder", "FastSCLLUTDecoder", "CASCLLUTDecoder", "CAFastSCLLUTDecoder"]:
        K = 32 + (24 if kind in common.CA_KINDS else 0)
        kw, x, truth = common.make_case(kind, N=128, K=K, L=4, A=32, B=20, seed=6, tables="minsum", ebn0_db=9.0)
        got = po.OracleDecoder(kind, **kw).decode(x)
        assert (got == truth).all(), kind
