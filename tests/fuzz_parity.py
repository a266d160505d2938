"""Open-ended random-configuration parity run (needs a GPU): python tests/fuzz_parity.py [seed]
400 random decoders per seed -- class, N, K, L (also non powers of two), alphabets, table sharing -- CUDA vs the oracle;
stops at the first mismatch and prints the configuration.  (The bounded version is tests/test_random_configs.py.)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, common
import quantized_decoder_polar_codes_b200 as q
from oracle import polar_oracle as po
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
BIG = len(sys.argv) > 2 and sys.argv[2] == "big"     # "big": the LUT Fast-SSC list kinds at N = 512 / 1024 (60 decoders)
FAST_LUT = ["FastSCLLUTDecoder", "CAFastSCLLUTDecoder", "FastSCLUTDecoder", "SCLLUTDecoder"]
for it in range(60 if BIG else 400):
    kind = FAST_LUT[rng.integers(0, 4)] if BIG else common.ALL_KINDS[rng.integers(0, 15)]
    n = int(rng.integers(9, 11)) if BIG else int(rng.integers(3, 9)); N = 1 << n
    ca = kind in common.CA_KINDS
    if ca and N < 32: N = 32
    K = int(rng.integers(25 if ca else 1, N))
    A = None
    if ca: A = K - 24 if kind == "CASCLDecoder" else max(1, K - int(rng.integers(0, 25)))
    L = int(rng.choice([2, 4, 8] if BIG else [1, 2, 3, 4, 5, 8, 12, 16, 32])) if kind in common.LIST_KINDS else 1
    Q = int(rng.choice([2, 4, 8, 16, 32])) if "LUT" in kind else 16
    Qc = int(rng.choice([Q, 2 * Q])) if "LUT" in kind else None
    share = bool(rng.integers(0, 2))
    kwc = dict(N=N, K=K, L=L, B=24 if BIG else 70, Q=Q, Qc=Qc, seed=int(rng.integers(0, 1 << 20)), share=share, per_position=not share, v=8 if Q < 16 else 16)
    if A is not None: kwc["A"] = A
    kw, x, _ = common.make_case(kind, **kwc)
    if "Fast" in kind and 0 <= kw["node_type"][0] <= (2 if kind in common.LIST_KINDS else 3):
        continue
    print(it, kind, {k: v for k, v in kwc.items() if k != "B"}, end=" ", flush=True)
    dec = getattr(q, kind)(**kw)
    print(dec.kernel, end=" ", flush=True)
    got = dec.decode(x)
    want = po.OracleDecoder(kind, **kw).decode(x)
    ok = (got == want).all()
    print("OK" if ok else "MISMATCH", flush=True)
    if not ok: sys.exit(1)
