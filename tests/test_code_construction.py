"""The inputs every Fast test and the benchmark take from simulation.py -- the NR-sequence frozen set and the Fast-SSC node
types -- pinned against the reference's own code (PolarCodesUtils/CodeConstruction.py:71-84 `PW`, :86-115 `GA`;
PolarCodesUtils/IdentifyNodes.py:13-150 `NodeIdentifier.run`, use_new_node=False).  Both are numpy and importable in the
build container; on a box without the reference tree the tests are skipped (the committed fixture below still runs)."""
import os
import sys

import numpy as np
import pytest

from quantized_decoder_polar_codes_b200 import compat, simulation as sim

REF = "/root/reference"
have_ref = os.path.isdir(os.path.join(REF, "PolarCodesUtils"))
FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "code_construction.npz")


def _ref_modules():
    compat.install(decoders=False, encoder=False, quantizers=False)      # np.int / np.loadtxt shims the reference code needs
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from PolarCodesUtils.CodeConstruction import PolarCodeConstructor
    from PolarCodesUtils.IdentifyNodes import NodeIdentifier
    return PolarCodeConstructor, NodeIdentifier


def visited(t, N):
    """heap ids of the internal nodes the decoders' walk reaches: node_type is read only there (a special node's descendants
    are never visited, leaves are decided by frozen_bits)"""
    out, stack = [], [(0, 0)]
    n = int(np.log2(N))
    while stack:
        d, node = stack.pop()
        p = (1 << d) + node - 1
        if d == n:
            continue
        out.append(p)
        if not (0 <= t[p] <= 3):
            stack += [(d + 1, 2 * node), (d + 1, 2 * node + 1)]
    return out


CASES = [(32, 16), (64, 20), (64, 57), (128, 32), (128, 64), (128, 100), (256, 128), (256, 152), (512, 100), (512, 256),
         (512, 400), (1024, 512), (1024, 536), (1024, 256), (1024, 900)]


@pytest.mark.skipif(not have_ref, reason="reference sources not present")
@pytest.mark.parametrize("N,K", CASES)
def test_frozen_set_and_node_types_match_the_reference(N, K):
    PolarCodeConstructor, NodeIdentifier = _ref_modules()
    frozenbits, msgbits, fmask, mmask = PolarCodeConstructor(N, K, os.path.join(REF, "reliable sequence.txt")).PW()
    fm, mm = sim.frozen_mask(N, K)
    assert (fm == fmask).all() and (mm == mmask).all()
    want = NodeIdentifier(N, K, frozenbits, msgbits, use_new_node=False).run().astype(np.int32)
    got = sim.identify_nodes(N, fm)
    vis = visited(want, N)
    assert vis == visited(got, N)
    assert (got[vis] == want[vis]).all()


@pytest.mark.skipif(not have_ref, reason="reference sources not present")
@pytest.mark.parametrize("N,K,ebn0", [(2048, 1024, 2.0), (1024, 512, 1.0), (256, 100, 3.0)])
def test_ga_construction_matches_the_reference(N, K, ebn0):
    PolarCodeConstructor, _ = _ref_modules()
    sigma = sim.awgn_sigma(ebn0, K / N)
    c = PolarCodeConstructor(N, K, os.path.join(REF, "reliable sequence.txt"))
    _, _, fmask, mmask = c.GA(sigma)
    fm, mm = sim.frozen_mask_ga(N, K, sigma)
    assert (fm == fmask).all() and (mm == mmask).all()


@pytest.mark.skipif(not have_ref, reason="reference sources not present")
@pytest.mark.parametrize("N,v,ebn0", [(64, 16, 2.0), (256, 8, 3.0), (512, 16, 2.0)])
def test_uniform_step_sizes_match_the_reference(N, v, ebn0):
    _ref_modules()
    from QuantizeDensityEvolution.QLLRDensityEvolution_OptUniform import LLRLSUniformQuantizer
    sigma = sim.awgn_sigma(ebn0, 0.5)
    rf, rg = LLRLSUniformQuantizer(N, v).generate_uniform_quantizers(sigma)
    mf, mg = sim.uniform_quantizer_steps(N, v, sigma)
    assert (rf == mf).all() and (rg == mg).all()


def test_committed_fixture():
    """reference outputs stored for the boxes without the reference tree (made by tests/golden/make_code_construction.py)"""
    z = np.load(FIX)
    for N, K in [(128, 64), (1024, 512), (1024, 536)]:
        fm, _ = sim.frozen_mask(N, K)
        assert (fm == z[f"pw_{N}_{K}/frozen"]).all()
        nt, want = sim.identify_nodes(N, fm), z[f"pw_{N}_{K}/node_type"]
        vis = visited(want, N)
        assert vis == visited(nt, N) and (nt[vis] == want[vis]).all()
    fm, _ = sim.frozen_mask_ga(2048, 1024, float(z["ga_2048_1024/sigma"]))
    assert (fm == z["ga_2048_1024/frozen"]).all()
    rf, rg = sim.uniform_quantizer_steps(2048, 16, float(z["ga_2048_1024/sigma"]))
    assert (rf == z["uq_2048_16/r_f"]).all() and (rg == z["uq_2048_16/r_g"]).all()
