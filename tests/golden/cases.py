"""The golden-vector case list: (case id, make_case keyword arguments).  Inputs are regenerated from the
seed by tests/common.py:make_case (numpy's default_rng streams are stable across versions); the fixture
stores a SHA-256 of the regenerated inputs/tables next to the reference's outputs, so a drift in the
generator is detected rather than silently compared."""

CASES = []


def _add(cid, **kw):
    CASES.append((cid, kw))


_ALL = ["SCDecoder", "FastSCDecoder", "SCLDecoder", "FastSCLDecoder", "CASCLDecoder",
        "SCLUTDecoder", "FastSCLUTDecoder", "SCLLUTDecoder", "FastSCLLUTDecoder", "CASCLLUTDecoder",
        "CAFastSCLLUTDecoder", "SCUniformQuantizedDecoder", "SCLUniformQuantizedDecoder",
        "SCLloydQuantizedDecoder", "SCLLloydQuantizedDecoder"]
# every class, tie-heavy random tables / coarse-grid LLRs
for _k in _ALL:
    _add(f"{_k}-N128-K64-L8", kind=_k, N=128, K=64, L=8, A=40, B=200, seed=11)
    _add(f"{_k}-N64-K20-L4", kind=_k, N=64, K=30, L=4, A=6, B=100, seed=12)
# 2L > 16: libstdc++ introsort tie order becomes observable
for _k in ["SCLLUTDecoder", "FastSCLLUTDecoder", "CASCLLUTDecoder", "CAFastSCLLUTDecoder", "SCLDecoder",
           "FastSCLDecoder", "SCLUniformQuantizedDecoder", "SCLLloydQuantizedDecoder"]:
    _add(f"{_k}-N256-K128-L16", kind=_k, N=256, K=128, L=16, A=104, B=40, seed=13)
    _add(f"{_k}-N128-K64-L32", kind=_k, N=128, K=64, L=32, A=40, B=40, seed=14)
# per-position tables that really differ (the API allows it: PD/src/SCLUTDecoder.cpp:57)
for _k in ["SCLUTDecoder", "SCLLUTDecoder", "FastSCLLUTDecoder"]:
    _add(f"{_k}-N64-perpos", kind=_k, N=64, K=32, L=8, B=60, seed=15, share=False, per_position=True)
# QChannel != QDecoder, LLR-domain style n+1 level table
_add("SCLLUTDecoder-N128-Qc32", kind="SCLLUTDecoder", N=128, K=32, L=8, B=100, seed=16, Q=16, Qc=32, llr_levels=8)
_add("FastSCLUTDecoder-N256-Q8", kind="FastSCLUTDecoder", N=256, K=128, B=100, seed=17, Q=8)
# BASELINE.json configs 1-3 (shape), working decoders on a real encode+AWGN chain
_add("C1-SC-N128-A64-awgn", kind="SCDecoder", N=128, K=64, B=300, seed=21, tables="channel", ebn0_db=2.0)
_add("C2-SCLUT-N128-A32-minsum", kind="SCLUTDecoder", N=128, K=32, B=300, seed=22, tables="minsum", ebn0_db=1.0)
_add("C3-SCLLUT-N128-A32-L8-minsum", kind="SCLLUTDecoder", N=128, K=32, L=8, B=300, seed=23, tables="minsum", ebn0_db=1.0)
_add("FastSCL-N256-awgn", kind="FastSCLDecoder", N=256, K=128, L=8, B=100, seed=24, tables="channel", ebn0_db=1.5)
_add("CASCL-N256-awgn", kind="CASCLDecoder", N=256, K=128 + 24, A=128, L=8, B=100, seed=25, tables="channel", ebn0_db=1.5)
# north-star shape and config 4 shape (N=1024, L=8): few frames, the reference needs ~20 ms each
_add("NS-SCLLUT-N1024-K512-L8", kind="SCLLUTDecoder", N=1024, K=512, L=8, B=24, seed=31)
_add("NS-SCLLUT-N1024-K512-L8-minsum", kind="SCLLUTDecoder", N=1024, K=512, L=8, B=24, seed=32, tables="minsum", ebn0_db=1.5)
_add("C4-CAFastSCLLUT-N1024-A512-L8", kind="CAFastSCLLUTDecoder", N=1024, K=536, A=512, L=8, B=24, seed=33)
_add("C4-CAFastSCLLUT-N1024-A512-L8-minsum", kind="CAFastSCLLUTDecoder", N=1024, K=536, A=512, L=8, B=24, seed=34, tables="minsum", ebn0_db=1.5)
_add("FastSCLUT-N1024-K512", kind="FastSCLUTDecoder", N=1024, K=512, B=100, seed=35)
# config 5 shape (N=2048, L=32, uniform quantizer; PW construction since the NR table stops at 1024)
_add("C5-SCLUniform-N2048-K1024-L32", kind="SCLUniformQuantizedDecoder", N=2048, K=1024, L=32, B=3, seed=41, construction="pw")
