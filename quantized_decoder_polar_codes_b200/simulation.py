"""Host-side helpers that sit either side of the decode hot path: code construction, node typing, the
encoder/CRC the reference drivers import from the (absent) `PolarBDEnc` package, the AWGN channel and
synthetic lookup tables.  Plain numpy; nothing here is on the GPU path.

Reference pointers (relative to the reference root):
  * frozen set from the 5G-NR reliability sequence: PolarCodesUtils/CodeConstruction.py:65-84 (`PW`)
  * node typing R0/R1/REP/SPC:                      PolarCodesUtils/IdentifyNodes.py:13-150
  * encoder / CRC conventions:                       SURVEY.md 8(c) (natural order, no bit reversal;
                                                     CRC long division of PD/src/utils.cpp:77-93)
  * channel model:                                   mainFPDecoder.py:91-110
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
CRC24_LOC = (24, 23, 21, 20, 17, 15, 13, 12, 8, 4, 2, 1, 0)  # mainQuantizedDecoder_LLRDomain.py:41-42


def nr_reliability_sequence():
    """3GPP TS 38.212 Table 5.3.1.2-1 (N_max = 1024), least reliable first."""
    return np.load(os.path.join(_DATA, "nr_reliability_sequence_1024.npy")).astype(np.int64)


def polarization_weight_sequence(N, beta=2 ** 0.25):
    """beta-expansion reliability order (least reliable first); used where the NR table (<=1024) stops."""
    n = int(np.log2(N))
    idx = np.arange(N)
    w = np.zeros(N)
    for j in range(n):
        w += ((idx >> j) & 1) * beta ** j
    return np.argsort(w, kind="stable")


def frozen_mask(N, K, construction="nr"):
    """-> (frozen_mask int32[N] (1 = frozen), message_mask int32[N]) ; CodeConstruction.py:71-84."""
    if construction == "nr" and N <= 1024:
        seq = nr_reliability_sequence()
        seq = seq[seq < N]
    else:
        seq = polarization_weight_sequence(N)
    fm = np.zeros(N, np.int32)
    fm[np.sort(seq[: N - K])] = 1
    return fm, (1 - fm).astype(np.int32)


def identify_nodes(N, frozen, spc=True):
    """Fast-SSC node classification (IdentifyNodes.py:13-150 with use_new_node=False).
    -> int32[2N-1] indexed by heap id (1<<depth)+node-1: -1 ordinary, 0 R0, 1 R1, 2 REP, 3 SPC.
    Children of a special node stay -1; leaves reached by the walk are typed 0/1."""
    n = int(np.log2(N))
    info = 1 - np.asarray(frozen, dtype=np.int64)
    t = -np.ones(2 * N - 1, np.int32)

    def walk(d, node):
        size = N >> d
        p = (1 << d) + node - 1
        seg = info[size * node: size * (node + 1)]
        s = int(seg.sum())
        if d == n:
            t[p] = 1 if s else 0
            return
        if s == 0:
            t[p] = 0
        elif s == size:
            t[p] = 1
        elif s == 1 and seg[-1] == 1:
            t[p] = 2
        elif s == size - 1 and seg[0] == 0:
            t[p] = 3
        else:
            walk(d + 1, 2 * node)
            walk(d + 1, 2 * node + 1)

    walk(0, 0)
    return t


def polar_transform(u):
    """x = u F^{(x)n} over GF(2), natural order (butterfly x[i+j]^=x[i+m+j], m=1,2,..N/2).  (B,N) or (N,)."""
    x = np.array(u, dtype=np.uint8, copy=True)
    flat = x.reshape(-1, x.shape[-1])
    N = flat.shape[1]
    m = 1
    while m < N:
        v = flat.reshape(flat.shape[0], N // (2 * m), 2, m)
        v[:, :, 0, :] ^= v[:, :, 1, :]
        m *= 2
    return x


def polar_encode(msg, frozen):
    """msg (B,K) uint8 -> codeword (B,N): u[non-frozen] = msg, x = polar_transform(u)."""
    msg = np.atleast_2d(np.asarray(msg, dtype=np.uint8))
    frozen = np.asarray(frozen)
    u = np.zeros((msg.shape[0], frozen.size), np.uint8)
    u[:, frozen == 0] = msg
    return polar_transform(u)


def crc_poly(crc_n, loc):
    p = np.zeros(crc_n + 1, np.uint8)
    p[np.asarray(loc, dtype=np.int64)] = 1
    return p


def crc_remainder(msg, crc_n=24, loc=CRC24_LOC):
    """MSB-first long division (PD/src/utils.cpp:77-93); msg (B,A) -> (B,crc_n)."""
    msg = np.atleast_2d(np.asarray(msg, dtype=np.uint8))
    B, A = msg.shape
    p = crc_poly(crc_n, loc)
    u = np.zeros((B, A + crc_n), np.uint8)
    u[:, :A] = msg
    for i in range(A):
        sel = u[:, i] == 1
        u[sel, i: i + crc_n + 1] ^= p
    return u[:, A:]


def crc_attach(msg, crc_n=24, loc=CRC24_LOC):
    msg = np.atleast_2d(np.asarray(msg, dtype=np.uint8))
    return np.concatenate([msg, crc_remainder(msg, crc_n, loc)], axis=1)


def awgn_sigma(ebn0_db, rate):
    return float(np.sqrt(1.0 / (2.0 * rate * 10 ** (ebn0_db / 10.0))))  # mainFPDecoder.py:91-92


def awgn_llr(codeword, sigma, rng):
    """BPSK 1-2x, AWGN, llr = 2y/sigma^2 (mainFPDecoder.py:106-109)."""
    y = (1.0 - 2.0 * codeword.astype(np.float64)) + sigma * rng.standard_normal(codeword.shape)
    return 2.0 * y / sigma ** 2


# ----------------------------------------------------------------------------------------------------
# synthetic lookup tables in the reference's pickle layout:
#   LUT_f[node][pos][a][b], LUT_g[node][pos][u][a][b], virtual_channel_llr[level][position][symbol]
def quantize_uniform(llr, Q, delta):
    """symbol s <-> value (s - Q/2 + 0.5)*delta, saturating."""
    s = np.floor(np.asarray(llr) / delta).astype(np.int64) + Q // 2
    return np.clip(s, 0, Q - 1)


def minsum_lut_tables(N, Q=16, Qc=None, delta=1.0, levels=None, per_position=True):
    """A working quantized decoder expressed as lookup tables: uniform-grid min-sum f and saturating g.
    Heavily tie-prone (few distinct LLR magnitudes), which is what parity tests want."""
    n = int(np.log2(N))
    Qc = Qc or Q
    levels = n if levels is None else levels

    def grid(q):
        return (np.arange(q) - q // 2 + 0.5) * delta

    LUT_f, LUT_g = [], []
    for d in range(n):
        qi = Qc if d == 0 else Q
        va = grid(qi)
        a, b = np.meshgrid(va, va, indexing="ij")
        f = np.sign(a) * np.sign(b) * np.minimum(np.abs(a), np.abs(b))
        tf = quantize_uniform(f, Q, delta).astype(np.int32)
        tg = np.stack([quantize_uniform(a + b, Q, delta), quantize_uniform(b - a, Q, delta)]).astype(np.int32)
        npos = (N >> (d + 1)) if per_position else 1
        for _ in range(1 << d):
            LUT_f.append(np.broadcast_to(tf, (npos,) + tf.shape))
            LUT_g.append(np.broadcast_to(tg, (npos,) + tg.shape))
    llr = np.broadcast_to(grid(Q), (levels, N, Q)).copy()
    return LUT_f, LUT_g, llr


def random_lut_tables(N, Q=16, Qc=None, rng=None, llr_alphabet=None, levels=None, per_position=False, share=True):
    """Random tables: exercise every code path; bit-exactness does not need meaningful tables.
    `llr_alphabet`: draw LLR entries from this small set to force path-metric / |LLR| ties and zeros."""
    rng = rng or np.random.default_rng(0)
    n = int(np.log2(N))
    Qc = Qc or Q
    levels = n if levels is None else levels
    LUT_f, LUT_g = [], []
    for d in range(n):
        qi = Qc if d == 0 else Q
        npos_full = N >> (d + 1)
        for _ in range(1 << d):
            if share:
                tf = rng.integers(0, Q, (1, qi, qi), dtype=np.int32)
                tg = rng.integers(0, Q, (1, 2, qi, qi), dtype=np.int32)
                if per_position:
                    tf = np.broadcast_to(tf, (npos_full, qi, qi))
                    tg = np.broadcast_to(tg, (npos_full, 2, qi, qi))
            else:
                tf = rng.integers(0, Q, (npos_full, qi, qi), dtype=np.int32)
                tg = rng.integers(0, Q, (npos_full, 2, qi, qi), dtype=np.int32)
            LUT_f.append(tf)
            LUT_g.append(tg)
    if llr_alphabet is None:
        llr = rng.standard_normal((levels, N, Q)) * 4.0
    else:
        llr = rng.choice(np.asarray(llr_alphabet, dtype=np.float64), size=(levels, N, Q))
    return LUT_f, LUT_g, llr


def to_reference_lists(LUT_f, LUT_g, llr):
    """What the reference drivers hand to the pybind constructors: plain nested Python lists."""
    return ([np.asarray(t).tolist() for t in LUT_f], [np.asarray(t).tolist() for t in LUT_g],
            np.asarray(llr).tolist())
