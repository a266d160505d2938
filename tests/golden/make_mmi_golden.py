"""Generates tests/golden/mmi_golden.npz: inputs and outputs of the REFERENCE's probability-domain table generator
(QuantizeDensityEvolution/QDensityEvolution_MMI.py: QDensityEvolutionMMI.run, driven like
GenerateLookUpTable_ProbabilityDomain.py:40-62) for small codes, plus stand-alone quantizer problems, as golden vectors for
quantized_decoder_polar_codes_b200/lutgen.py (mmi_*).

The generator's `MMIQuantizer` is C++ on OpenCV (Quantizers/quantizers/_cpp/MMIQuantizer, cannot be built here); it is served
by the reference's own numpy restatement QuantizeDensityEvolution/MMIQuantizer.py (class MMIQunatizer) behind an adapter
with the C++ call signature (MMIQuantizer.cpp:73-165: returns Q, P(z|x), Az, permutation).  Two implementation-defined
points of that restatement are pinned so that the vectors do not depend on the machine that made them:
  * its `np.argsort(llr)` (introsort / AVX-512 sort, order of equal keys unspecified) is run as a STABLE sort;
  * the joint distribution is handed over as float64 (what the C++ binding's py::array_t<double> does to the float32
    arrays the generator passes) -- numpy would otherwise do the whole design in float32.
Run in the build container only:  python tests/golden/make_mmi_golden.py   (~10 minutes)
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, REF)

from QuantizeDensityEvolution import MMIQuantizer as _mq  # noqa: E402  (the numpy restatement)


class _StableNumpy:
    """numpy with a stable argsort, handed to the restatement's module namespace only"""
    def __getattr__(self, name):
        return getattr(np, name)

    @staticmethod
    def argsort(a, *args, **kw):
        return np.argsort(a, kind="stable")


_mq.np = _StableNumpy()


class MMIQuantizer:
    """quantizers.quantizer.MMI.MMIQuantizer (C++) served by the numpy restatement"""
    def __init__(self, px1=0.5, px_minus1=0.5):
        self.q = _mq.MMIQunatizer(px1, px_minus1)

    def find_opt_quantizer(self, joint_prob, K):
        joint = np.ascontiguousarray(joint_prob, dtype=np.float64)
        Q = self.q.find_opt_quantizer(joint, K)
        llr = np.log2(joint[0] / joint[1])
        perm = np.argsort(llr, kind="stable")
        Az = np.zeros(K + 1, dtype=np.int64)
        pzx = np.zeros((2, K))
        for i in range(K):
            Az[i + 1] = Az[i] + int(Q[i].sum())
            for j in range(Az[i], Az[i + 1]):            # MMIQuantizer.cpp:151-160: sequential accumulation in sorted order
                assert Q[i, perm[j]] == 1
                pzx[0, i] += joint[0, perm[j]]
                pzx[1, i] += joint[1, perm[j]]
        return Q.astype(np.int32), pzx, Az, perm


for name in ["quantizers", "quantizers.quantizer", "quantizers.quantizer.MMI"]:
    sys.modules[name] = types.ModuleType(name)
sys.modules["quantizers.quantizer.MMI"].MMIQuantizer = MMIQuantizer

from QuantizeDensityEvolution.QDensityEvolution_MMI import QDensityEvolutionMMI  # noqa: E402


def channel_probs(qc, design_db, rng):
    """P(z|x) of a Qc-level symmetric channel: a plain uniform quantizer of y = x + n on [-1-3s, 1+3s] (the MMI channel
    quantizer of the generator script only moves the cell edges; any valid P(z|x) exercises the density evolution)."""
    from scipy.stats import norm
    s = np.sqrt(1 / 10 ** (design_db / 10))
    edges = np.linspace(-1 - 3 * s, 1 + 3 * s, qc + 1)
    edges[0], edges[-1] = -np.inf, np.inf
    p = np.zeros((2, qc))
    p[0] = np.diff(norm.cdf(edges, loc=1, scale=s))
    p[1] = np.diff(norm.cdf(edges, loc=-1, scale=s))
    return p


def main():
    out = {}
    rng = np.random.default_rng(11)
    # stand-alone quantizer problems: random joint distributions, some with exactly equal likelihood ratios
    nq = 0
    for M, K, ties in [(16, 4, False), (64, 8, True), (64, 16, False), (128, 16, True), (256, 16, False), (512, 16, True), (40, 16, False)]:
        j = rng.random((2, M)) + 1e-3
        if ties:
            j[:, M // 2:] = j[::-1, : M - M // 2]        # mirrored columns: llr(i) = -llr(mirror), plus exact duplicates
            j[:, 1] = j[:, 0]
        j /= j.sum(axis=1, keepdims=True)
        j = j.astype(np.float32).astype(np.float64)
        Q, pzx, Az, perm = MMIQuantizer().find_opt_quantizer(j, K)
        out[f"q{nq}/joint"], out[f"q{nq}/K"] = j, np.int32(K)
        out[f"q{nq}/Q"], out[f"q{nq}/pzx"], out[f"q{nq}/Az"], out[f"q{nq}/perm"] = Q, pzx, Az, perm
        nq += 1
        print("quantizer problem", M, K, "done", flush=True)
    out["nq"] = np.int32(nq)
    # whole generator runs (level 0 tables are Qc x Qc, deeper ones Qd x Qd)
    for tag, N, qd, qc, db in [("n16q4", 16, 4, 4, 2.0), ("n32q8c6", 32, 8, 6, 3.0), ("n64q16", 64, 16, 16, 3.0)]:
        pzx = channel_probs(qc, db, rng)
        lut_fs, lut_gs, llrs, probs = QDensityEvolutionMMI(N, qd).run(pzx.copy())
        out[tag + "/pzx"] = pzx
        out[tag + "/llrs"], out[tag + "/probs"] = llrs, probs
        out[tag + "/f_root"] = np.asarray(lut_fs[0][0], np.int32)
        out[tag + "/g_root"] = np.asarray(lut_gs[0][0], np.int32)
        out[tag + "/lut_f"] = np.stack([np.asarray(lut_fs[p][0], np.int32) for p in range(1, N - 1)])
        out[tag + "/lut_g"] = np.stack([np.asarray(lut_gs[p][0], np.int32) for p in range(1, N - 1)])
        print(tag, "done", flush=True)
    np.savez_compressed(os.path.join(HERE, "mmi_golden.npz"), **out)


if __name__ == "__main__":
    main()
