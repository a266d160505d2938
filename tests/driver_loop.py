"""The frame loop of the reference's LLR-domain driver, mainQuantizedDecoder_LLRDomain.py:64-192, restated as a function
with the same order of np.random calls, so that on one seed it sees the frames the unmodified script sees.  Test
infrastructure: test_driver_plugin.py runs the unmodified script (where the reference tree exists) and this loop on the same
pickles and seed, with the compiled reference's decoders and with this package's.  Everything that the script imports is
passed in: decoder classes, encoder classes, the channel quantizer and the two design helpers."""
import os
import pickle as pkl
from bisect import bisect_left

import numpy as np

CRC_N, CRC_P = 24, [24, 23, 21, 20, 17, 15, 13, 12, 8, 4, 2, 1, 0]


def write_llr_domain_pickles(root, N, qc, qd, design_db, lut_f, lut_g, llr_quanta, quantizer="MinDistortion"):
    """tables -> ./LUT/<Quantizer>/N.._ChannelQ.._DecoderQ../{LUT_F,LUT_G,LLRQuanta}_SNRdB=<d>.pkl in the layout
    GenerateLookUpTable_LLRDomain.py:31,57-68 writes: dict heap id -> list over the node's positions of the node's table."""
    d = os.path.join(root, "LUT", quantizer, "N{:d}_ChannelQ{:d}_DecoderQ{:d}".format(N, qc, qd))
    os.makedirs(d, exist_ok=True)
    n = int(np.log2(N))
    fs, gs = {}, {}
    for depth in range(n):
        for node in range(1 << depth):
            p = (1 << depth) + node - 1
            npos = N >> (depth + 1)
            fs[p] = [np.asarray(lut_f[p], dtype=np.float64) for _ in range(npos)]     # the generator stores float arrays
            gs[p] = [np.asarray(lut_g[p], dtype=np.float64) for _ in range(npos)]
    for name, obj in (("LUT_F", fs), ("LUT_G", gs), ("LLRQuanta", np.asarray(llr_quanta))):
        with open(os.path.join(d, "{:s}_SNRdB={:.0f}.pkl".format(name, design_db)), "wb") as f:
            pkl.dump(obj, f)
    return d


def run(workdir, N, A, L, decoder_type, is_crc, qcu, qd, qc, design_db, frames, seed, ebn0_list, decoders, PolarEnc, CRCEnc,
        LLRQuantizer, frozen_sets, node_types, channel_llr_density_table, quantizer="MinDistortion"):
    """-> list over Eb/N0 of (bit errors, block errors, blocks) and the decoded words [point][frame]."""
    K = A + CRC_N if is_crc else A
    rate = A / N
    load_dir = os.path.join(workdir, "LUT", quantizer, "N{:d}_ChannelQ{:d}_DecoderQ{:d}".format(N, qc, qd))
    frozenbits, msgbits, frozen_ind, message_ind = frozen_sets(N, K)
    node_type = node_types(N, K, frozenbits, msgbits).astype(np.int32)
    polar_encoder = PolarEnc(N, K, frozenbits, msgbits)
    crc_encoder = CRCEnc(CRC_N, CRC_P)
    with open(os.path.join(load_dir, "LUT_F_SNRdB={:.0f}.pkl".format(design_db)), "rb") as f:
        lut_fs = pkl.load(f)
    with open(os.path.join(load_dir, "LUT_G_SNRdB={:.0f}.pkl".format(design_db)), "rb") as f:
        lut_gs = pkl.load(f)
    with open(os.path.join(load_dir, "LLRQuanta_SNRdB={:.0f}.pkl".format(design_db)), "rb") as f:
        vllr = pkl.load(f).tolist()
    fs = [np.array(lut_fs[k]).astype(np.int32).tolist() for k in lut_fs.keys()]          # :88-95
    gs = [np.array(lut_gs[k]).astype(np.int32).tolist() for k in lut_gs.keys()]
    make = {
        "SC-LUT": lambda: decoders["SCLUTDecoder"](N, K, frozen_ind, message_ind, fs, gs, vllr),
        "SCL-LUT": lambda: decoders["SCLLUTDecoder"](N, K, L, frozen_ind, message_ind, fs, gs, vllr),
        "FastSC-LUT": lambda: decoders["FastSCLUTDecoder"](N, K, frozen_ind, message_ind, node_type, fs, gs, vllr),
        "FastSCL-LUT": lambda: decoders["FastSCLLUTDecoder"](N, K, L, frozen_ind, message_ind, node_type, fs, gs, vllr),
        "CASCL-LUT": lambda: decoders["CASCLLUTDecoder"](N, K, A, L, frozen_ind, message_ind, CRC_N, CRC_P, fs, gs, vllr),
    }
    polar_decoder = make[decoder_type]()
    np.random.seed(seed)
    stats, words = [], []
    for ebn0_db in ebn0_list:
        sigma = np.sqrt(1 / (2 * rate * 10 ** (ebn0_db / 10)))
        e_llr = 2 / (sigma ** 2)
        d_llr = np.sqrt(2 * e_llr)
        pyx, interval_x, quanta = channel_llr_density_table(qcu, -e_llr - 3 * d_llr, e_llr + 3 * d_llr, e_llr, -e_llr, d_llr)
        _, _, channel_lut, _ = LLRQuantizer().find_OptLS_quantizer(pyx, quanta, qcu, qc)
        channel_lut = np.asarray(channel_lut).squeeze()
        nbit = nblk = nblocks = 0
        out = []
        for _ in range(frames):
            msg = np.random.randint(low=0, high=2, size=A)
            cword = np.asarray(polar_encoder.encode(crc_encoder.encode(msg) if is_crc else msg)).astype(int)
            y = (1 - 2 * cword) + np.random.normal(loc=0, scale=sigma, size=(1, N))
            llr = y * 2 / (sigma ** 2)
            sym = np.zeros(N).astype(np.int32)
            for i in range(N):                                                         # :167-176
                if llr[0, i] <= interval_x[0]:
                    sym[i] = 0
                elif llr[0, i] >= interval_x[-1]:
                    sym[i] = qc - 1
                else:
                    sym[i] = channel_lut[bisect_left(interval_x[:-1], llr[0, i]) - 1]
            dec = np.asarray(polar_decoder.decode(sym))
            out.append(dec.copy())
            nbit += int(np.sum(msg != dec))
            nblk += int(np.any(msg != dec))
            if nblk > 1000:
                break
            nblocks += 1
        stats.append((nbit, nblk, nblocks))
        words.append(np.stack(out))
    return stats, words
