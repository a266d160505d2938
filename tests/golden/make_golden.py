"""Regenerates tests/golden/golden.npz from the compiled, unmodified reference (oracle/_ref).
Run in the build container (needs /root/reference for `make -C oracle ref`):  python tests/golden/make_golden.py
Also refreshes the NR reliability-sequence data file of the package from the reference's table."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import polar_oracle as po  # noqa: E402
import common  # noqa: E402
from golden.cases import CASES  # noqa: E402


def case_hash(kw, x):
    h = hashlib.sha256()
    for k in sorted(kw):
        v = kw[k]
        if isinstance(v, (list, tuple)) and len(v) and isinstance(v[0], np.ndarray):
            for t in v:
                h.update(np.ascontiguousarray(t).tobytes())
        else:
            h.update(np.ascontiguousarray(np.asarray(v)).tobytes())
    h.update(np.ascontiguousarray(x).tobytes())
    return h.hexdigest()


def main():
    po.build_reference()
    ref = po.load_reference()
    assert ref is not None, "oracle/_ref is not built"
    out = {}
    for cid, ckw in CASES:
        ckw = dict(ckw)
        kind = ckw.pop("kind")
        kw, x, _ = common.make_case(kind, **ckw)
        y = common.ref_decode(ref, kind, kw, x)
        out[cid + "/out"] = np.packbits(y, axis=1)
        out[cid + "/shape"] = np.asarray(y.shape, np.int64)
        out[cid + "/sha"] = np.frombuffer(bytes.fromhex(case_hash(kw, x)), np.uint8)
        print(f"{cid:45s} frames {y.shape[0]:4d}  bits/frame {y.shape[1]:4d}  ones {y.mean():.3f}")
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    seq_src = "/root/reference/reliable sequence.txt"
    if os.path.exists(seq_src):
        seq = np.array([int(float(t)) for t in open(seq_src).read().split()], dtype=np.int16)
        np.save(os.path.join(ROOT, "quantized_decoder_polar_codes_b200", "data", "nr_reliability_sequence_1024.npy"), seq)


if __name__ == "__main__":
    main()
