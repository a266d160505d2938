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


def _ga_phi(x):
    """CodeConstruction.py:38-49 (Chung's approximation; note the exponent 0.859 here and 0.86 in the inverse -- kept)."""
    if 0 <= x <= 10:
        return np.exp(-0.4527 * np.power(x, 0.859) + 0.0218)
    return np.sqrt(np.pi / x) * np.exp(-x / 4) * (1 - 10 / 7 / x)


def _ga_dphi(x):
    """CodeConstruction.py:3-13"""
    if 0 <= x <= 10:
        return -0.4527 * 0.86 * np.power(x, -0.14) * _ga_phi(x)
    return np.exp(-x / 4) * np.sqrt(np.pi / x) * (-1 / 2 / x * (1 - 10 / 7 / x) - 1 / 4 * (1 - 10 / 7 / x) + 10 / 7 / x / x)


def _ga_phi_inverse(x):
    """CodeConstruction.py:15-36: closed form on [0.0388, 1.0221], Newton iteration elsewhere (same start, same stop rule)."""
    if 0.0388 <= x <= 1.0221:
        return np.power((0.0218 - np.log(x)) / 0.4527, 1 / 0.86)
    x0 = 0.0388
    x1 = x0 - (_ga_phi(x0) - x) / _ga_dphi(x0)
    delta, eps = abs(x1 - x0), 1e-3
    while delta >= eps:
        x0 = x1
        x1 = x1 - (_ga_phi(x1) - x) / _ga_dphi(x1)
        if x1 > 1e2:
            eps = 10
        delta = abs(x1 - x0)
    return x1


def _bitreverse(x):
    """CodeConstruction.py:51-61"""
    if len(x) == 1:
        return x
    return np.concatenate([_bitreverse(x[0::2]), _bitreverse(x[1::2])])


def frozen_mask_ga(N, K, sigma):
    """Gaussian-approximation construction, PolarCodeConstructor.GA(sigma) (CodeConstruction.py:86-115): mean LLR of every bit
    channel by the f / g recursions, the K largest (after the reference's bit reversal of the leaf order) carry information.
    Used for N = 2048 (BASELINE config 5): the NR reliability table stops at 1024.
    -> (frozen_mask int32[N], message_mask int32[N])."""
    n = int(np.log2(N))
    u = np.zeros((n + 1, N))
    u[0, :] = 2 / sigma ** 2
    for level in range(1, n + 1):
        nb_parent = 2 ** (n - level + 1)
        nb = nb_parent // 2
        for node in range(2 ** (level - 1)):
            tmp = u[level - 1, nb_parent * node]
            f = _ga_phi_inverse(1 - (1 - _ga_phi(tmp)) ** 2)
            g = 2 * tmp
            u[level, 2 * node * nb:(2 * node + 1) * nb] = f
            u[level, (2 * node + 1) * nb:(2 * node + 2) * nb] = g
    order = np.argsort(_bitreverse(u[-1, :]))
    fm = np.zeros(N, np.int32)
    fm[order[: N - K]] = 1
    return fm, (1 - fm).astype(np.int32)


def _uq_dr(mu, sigma2, v, r):
    """derivative of the uniform quantizer's distortion w.r.t. its step r for the mixture N(mu,s2)/2 + N(-mu,s2)/2
    (OptUniformQuantizerGaussian.py:76-101: dr_middle / dr_last / dr_bimodal_Gaussian, same order of operations)"""
    from scipy.special import erf

    def gauss(m, x):
        return 1 / np.sqrt(2 * np.pi * sigma2) * np.exp(-(x - m) ** 2 / (2 * sigma2))

    def middle(a, b, rv):
        p1 = -sigma2 / 2 * (gauss(mu, b) - gauss(mu, a))
        p2 = 0.25 * (mu - rv) * (erf((b - mu) / (np.sqrt(2 * sigma2))) - erf((a - mu) / (np.sqrt(2 * sigma2))))
        p3 = -sigma2 / 2 * (gauss(-mu, b) - gauss(-mu, a))
        p4 = 0.25 * (-mu - rv) * (erf((b - -mu) / (np.sqrt(2 * sigma2))) - erf((a - -mu) / (np.sqrt(2 * sigma2))))
        return 2 * (p1 + p2 + p3 + p4)

    def last(a, rv):
        p1 = sigma2 / 2 * gauss(mu, a) + 0.25 * (mu - rv) * (1 - erf((a - mu) / (np.sqrt(2 * sigma2))))
        p2 = sigma2 / 2 * gauss(-mu, a) + 0.25 * (-mu - rv) * (1 - erf((a - -mu) / (np.sqrt(2 * sigma2))))
        return 2 * (p1 + p2)

    result = 0
    half = v // 2
    for k in range(1, half):
        result -= (2 * k - 1) * middle((k - 1) * r, k * r, (k - 1 / 2) * r)
    result -= (2 * half - 1) * last((half - 1) * r, (half - 1 / 2) * r)
    return result


def _uq_opt_step(mu, sigma2, v, max_iter=30):
    """OptUniformQuantizerGaussian.find_optimal_interval_bimodal_Gaussian (:134-149): Newton iteration on dr, second
    derivative by central differences (delta 1e-6), stop when dr moves by < 1e-6."""
    r = 2 * (mu + 3 * np.sqrt(sigma2) - (-mu - 3 * np.sqrt(sigma2))) / (v - 2)
    for _ in range(max_iter):
        d2r = (_uq_dr(mu, sigma2, v, r + 1e-6) - _uq_dr(mu, sigma2, v, r - 1e-6)) / (2 * 1e-6)
        dr = _uq_dr(mu, sigma2, v, r)
        r -= dr / d2r
        if np.abs(_uq_dr(mu, sigma2, v, r) - dr) < 1e-6:
            break
    return r


def uniform_quantizer_steps(N, v, sigma):
    """LLRLSUniformQuantizer(N, v).generate_uniform_quantizers(sigma) (QLLRDensityEvolution_OptUniform.py:11-36): the step sizes
    decoder_r_f / decoder_r_g [N-1] (heap order) the continuous-domain driver hands to the uniformly quantized decoders
    (mainQuantizedDecoder_ContinuousDomain.py:99-104) -- Gaussian approximation of the LLR mean per node, optimal uniform step
    for the resulting +-mu mixture with variance 2 mu."""
    n = int(np.log2(N))
    mu_llr = np.zeros((n + 1, N))
    mu_llr[0, :] = 2 / sigma ** 2
    r_f, r_g = np.zeros(N - 1), np.zeros(N - 1)
    for level in range(1, n + 1):
        nb_parent = 2 ** (n - level + 1)
        nb = nb_parent // 2
        for node in range(2 ** (level - 1)):
            p = 2 ** (level - 1) - 1 + node
            mu = mu_llr[level - 1, nb_parent * node]
            mu_f = _ga_phi_inverse(1 - (1 - _ga_phi(mu)) ** 2)
            r_f[p] = _uq_opt_step(mu_f, 2 * mu_f, v)
            mu_g = 2 * mu
            r_g[p] = _uq_opt_step(mu_g, 2 * mu_g, v)
            mu_llr[level, 2 * node * nb:(2 * node + 1) * nb] = mu_f
            mu_llr[level, (2 * node + 1) * nb:(2 * node + 2) * nb] = mu_g
    return r_f, r_g


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
