#!/usr/bin/env python
"""Per-call latency of decode() for single frames (how the reference drivers call it) vs the compiled reference."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
import quantized_decoder_polar_codes_b200 as q
from oracle import polar_oracle as po
ref = po.load_reference()
for kind, ckw in [("SCLUTDecoder", dict(N=128, K=32)), ("SCLLUTDecoder", dict(N=128, K=32, L=8)), ("SCLLUTDecoder", dict(N=1024, K=512, L=8)), ("SCDecoder", dict(N=128, K=64))]:
    kw, x, _ = common.make_case(kind, B=200, seed=1, **ckw)
    dec = getattr(q, kind)(**kw)
    xs = [np.ascontiguousarray(x[i]) for i in range(200)]
    for i in range(20): dec.decode(xs[i])
    t = time.perf_counter()
    for i in range(200): dec.decode(xs[i])
    ours = (time.perf_counter() - t) / 200 * 1e6
    r = None
    if ref is not None:
        rd = getattr(ref, kind)(**common.ref_kwargs(kw))
        n = 200 if ckw["N"] == 128 else 20
        t = time.perf_counter()
        for i in range(n): rd.decode(xs[i])
        r = (time.perf_counter() - t) / n * 1e6
    print(f"{kind:16s} N={ckw['N']:5d} L={ckw.get('L',1)}  ours {ours:8.1f} us/call   reference {r if r is None else round(r,1)} us/call")
