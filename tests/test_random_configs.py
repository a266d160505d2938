"""Randomised configuration sweeps (hypothesis): code length, rate, list size (including non powers of two), table
alphabets, per-position tables, LLR alphabets engineered for ties.  CPU: oracle vs the compiled reference.
GPU: CUDA (whatever kernel the shape is routed to) vs the oracle."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import common
from oracle import polar_oracle as po

KINDS = common.ALL_KINDS


@st.composite
def configs(draw, max_n_log=8, max_l=32):
    kind = draw(st.sampled_from(KINDS))
    n = draw(st.integers(3, max_n_log))
    N = 1 << n
    ca = kind in common.CA_KINDS
    lo = 25 if ca else 1
    if ca and N - 1 < lo:
        n, N = 5, 32
    K = draw(st.integers(lo, N - 1))
    A = K - draw(st.integers(0, 24)) if ca else None
    if ca and kind == "CASCLDecoder":
        A = K - 24
    if ca and A < 1:
        A = 1 if kind != "CASCLDecoder" else None
    L = draw(st.sampled_from([1, 2, 3, 4, 5, 8, 12, 16, 32])) if kind in common.LIST_KINDS else 1
    L = min(L, max_l)
    Q = draw(st.sampled_from([2, 4, 8, 16, 32])) if "LUT" in kind else 16
    Qc = draw(st.sampled_from([Q, Q, 2 * Q])) if "LUT" in kind else None
    alphabet = draw(st.sampled_from([common.TIE_ALPHABET, (-1.0, 1.0), (-3.0, -1.0, 0.0, 1.0, 3.0), None]))
    share = draw(st.booleans())
    seed = draw(st.integers(0, 10 ** 6))
    return dict(kind=kind, N=N, K=K, A=A, L=L, Q=Q, Qc=Qc, alphabet=alphabet, share=share, seed=seed)


def _case(c, B):
    kw = dict(N=c["N"], K=c["K"], L=c["L"], B=B, Q=c["Q"], Qc=c["Qc"], seed=c["seed"], alphabet=c["alphabet"],
              share=c["share"], per_position=not c["share"], v=8 if c["Q"] < 16 else 16)
    if c["A"] is not None:
        kw["A"] = c["A"]
    return common.make_case(c["kind"], **kw)


def _degenerate(c, kw):
    """Fast kinds whose root is a special node are rejected by pd_create (the reference reads level -1 there)."""
    if "Fast" not in c["kind"]:
        return False
    t = kw["node_type"][0]
    return 0 <= t <= (2 if c["kind"] in common.LIST_KINDS else 3)


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(configs(max_n_log=7))
def test_oracle_vs_compiled_reference_random(refmod, c):
    if c["kind"] == "CASCLDecoder" and c["K"] < 25:
        return
    kw, x, _ = _case(c, B=12)
    if _degenerate(c, kw):
        return
    want = common.ref_decode(refmod, c["kind"], kw, x)
    got = po.OracleDecoder(c["kind"], **kw).decode(x)
    assert (got == want).all(), c


@pytest.mark.gpu
@settings(max_examples=120, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(configs(max_n_log=8))
def test_cuda_vs_oracle_random(c):
    import quantized_decoder_polar_codes_b200 as q
    if c["kind"] == "CASCLDecoder" and c["K"] < 25:
        return
    kw, x, _ = _case(c, B=70)
    if _degenerate(c, kw):
        with pytest.raises(ValueError):
            getattr(q, c["kind"])(**kw)
        return
    dec = getattr(q, c["kind"])(**kw)
    got = dec.decode(x)
    want = po.OracleDecoder(c["kind"], **kw).decode(x)
    assert (got == want).all(), (c, dec.kernel)
