"""LLR-domain lookup-table generation on the GPU: what the reference's
GenerateLookUpTable_LLRDomain.py does through QuantizeDensityEvolution/QLLRDensityEvolution_MinDistortion.py
(`LLRQuantizerSC.run`) -- quantized density evolution over the polar tree with a minimum-distortion quantizer after
every f and g step -- with all quantizer problems of a tree level solved in ONE batched launch (pd_optls_quantize,
csrc/pb_lutgen.cuh) instead of one slow dynamic programme per node.

    gen = MinDistortionLUTGenerator(N=1024, v=16)
    llr_density, llr_quanta, lut_f, lut_g = gen.run(channel_density, channel_quanta)
    kwargs = gen.decoder_tables(lut_f, lut_g, llr_quanta)       # LUT_f / LUT_g / virtual_channel_llr for the decoders

Results are bit-identical to the reference generator driven with its own numpy quantizer
(MinDistortionQuantizer.find_OptLS_quantizer) -- tests/test_lutgen.py, golden vectors made by that code.  The cheap glue
(product distributions, merging equal values with numpy's own np.unique / np.sum) stays in numpy on the host; the
quantizer, which is > 99 % of the reference's run time, is the CUDA kernel.  There is no CPU fallback for it."""
import numpy as np

from . import capi


def optls_quantize_batch(densities, quantas, K, device=0):
    """Batched LLRQuantizer.find_OptLS_quantizer.  densities / quantas: lists of 1-D arrays, each sorted by strictly
    ascending quanta (np.unique output) with more than K entries.  -> (density [P,K], quanta [P,K], list of int32 luts)."""
    P = len(densities)
    M = np.array([len(d) for d in densities], np.int32)
    stride = int(M.max())
    D = np.zeros((P, stride))
    Q = np.zeros((P, stride))
    for p in range(P):
        D[p, :M[p]] = densities[p]
        Q[p, :M[p]] = quantas[p]
    od, oq = np.empty((P, K)), np.empty((P, K))
    lut = np.empty((P, stride), np.int32)
    capi.check(capi.lib().pd_optls_quantize(D.ctypes.data, Q.ctypes.data, M.ctypes.data, stride, P, int(K),
                                            od.ctypes.data, oq.ctypes.data, lut.ctypes.data, int(device)))
    return od, oq, [lut[p, :M[p]].copy() for p in range(P)]


def _merge_equal(density, quanta):
    """LLRQuantizerSC.get_unique_quanta (QLLRDensityEvolution_MinDistortion.py:44-51): symbols with the same value are
    merged first; the masses are added with np.sum over the selected entries, like the reference."""
    uq, inverse = np.unique(quanta, return_inverse=True)
    ud = np.array([np.sum(density[quanta == u]) for u in uq])
    return ud, uq, inverse.astype(np.int32)


class MinDistortionLUTGenerator:
    def __init__(self, N, v, device=0):
        self.N, self.v, self.device = int(N), int(v), int(device)
        self.n = int(np.log2(N))
        if 1 << self.n != self.N:
            raise ValueError("N must be a power of two")

    def _quantize_level(self, problems):
        """problems: list of (density, quanta) after merging equal values -> list of (density[v], quanta[v], lut_merge)."""
        v = self.v
        res = [None] * len(problems)
        big = [i for i, (d, q) in enumerate(problems) if len(d) > v]
        if big:
            od, oq, luts = optls_quantize_batch([problems[i][0] for i in big], [problems[i][1] for i in big], v, self.device)
            for k, i in enumerate(big):
                res[i] = (od[k], oq[k], luts[k])
        for i, (d, q) in enumerate(problems):
            if res[i] is None:      # at most v distinct values: nothing to merge (identity on the sorted symbols, zero padded)
                dd, qq = np.zeros(v), np.zeros(v)
                dd[:len(d)], qq[:len(q)] = d, q
                res[i] = (dd, qq, np.arange(len(d), dtype=np.int32))
        return res

    def run(self, channel_llr_density, channel_llr_quanta):
        """-> llr_density [n+1,N,v], llr_quanta [n+1,N,v] (= virtual_channel_llr), lut_f [N-1,v,v], lut_g [N-1,2,v,v]
        (heap order, one table per node: the reference replicates it for every position of the node)."""
        N, n, v = self.N, self.n, self.v
        dens = np.zeros((n + 1, N, v))
        quan = np.zeros((n + 1, N, v))
        dens[0, :, :] = channel_llr_density
        quan[0, :, :] = channel_llr_quanta
        lut_f = np.zeros((N - 1, v, v), np.int32)
        lut_g = np.zeros((N - 1, 2, v, v), np.int32)
        for level in range(n):
            nb = 1 << (n - level)
            half = nb // 2
            nodes = 1 << level
            problems, erasure = [], []
            for node in range(nodes):
                off = node * nb
                d1, d2 = dens[level, off], dens[level, off + half]
                q1, q2 = quan[level, off], quan[level, off + half]
                # f: min-sum of the two symbols, joint probability (QLLRDensityEvolution_MinDistortion.py:15-31)
                a, b = q1[:, None], q2[None, :]
                qf = (np.sign(a) * np.sign(b) * np.minimum(np.abs(a), np.abs(b))).reshape(-1)
                df = (d1[:, None] * d2[None, :]).reshape(-1)
                # g: (1-2u)a + b for u = 0, 1, each with half of the joint probability (:18-42)
                u = np.arange(2)[:, None, None]
                qg = ((1 - 2 * u) * a[None] + b[None]).reshape(-1)
                dg = np.broadcast_to((0.5 * d1[:, None] * d2[None, :])[None], (2, v, v)).reshape(-1)
                for d_, q_ in ((df, qf), (dg, qg)):
                    ud, uq, inv = _merge_equal(d_, q_)
                    problems.append((ud, uq))
                    erasure.append(inv)
            solved = self._quantize_level(problems)
            for node in range(nodes):
                off = node * nb
                p = (1 << level) + node - 1
                (cdf, cqf, mf), (cdg, cqg, mg) = solved[2 * node], solved[2 * node + 1]
                lut_f[p] = mf[erasure[2 * node]].reshape(v, v)
                lut_g[p] = mg[erasure[2 * node + 1]].reshape(2, v, v)
                dens[level + 1, off:off + half] = cdf
                quan[level + 1, off:off + half] = cqf
                dens[level + 1, off + half:off + nb] = cdg
                quan[level + 1, off + half:off + nb] = cqg
        return dens, quan, lut_f, lut_g

    @staticmethod
    def decoder_tables(lut_f, lut_g, llr_quanta):
        """Constructor arguments of the LUT decoders in the compact one-table-per-node form they accept."""
        return dict(LUT_f=[t[None] for t in lut_f], LUT_g=[t[None] for t in lut_g], virtual_channel_llr=llr_quanta)
