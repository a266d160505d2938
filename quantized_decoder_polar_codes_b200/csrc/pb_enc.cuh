// Batched polar encoder + CRC ("next" row f2: the PolarBDEnc package the reference drivers import --
// PolarEnc(N,K,frozenbits,msgbits).encode, CRCEnc(crc_n,crc_p).encode, mainFPDecoder.py:12-13,56-57,102-105 -- is not in
// the reference tree; the conventions are the ones of the code that is):
//   CRC   : word = msg || CRC::encoding(msg), MSB-first long division (PD/src/utils.cpp:77-93).  Linear, so the
//           remainder of a message is the XOR of the host-precomputed remainders of its set bits.
//   polar : u[msgbits[k]] = word[k], x = u F^{(x)n} in natural order (the butterfly of the decoders' own re-encode,
//           PD/src/FastSCDecoder.cpp:153-164).
// One warp per frame; pure streaming kernels: `in_len` bytes in, `out_len` bytes out per frame (one byte per bit, as
// the drivers' numpy arrays) -- the roofline is HBM bandwidth.
//   encode_words_kernel (N <= 1024): everything in registers, one 32-bit word of the code per lane.  The message bytes
//     become a compact bit string by warp ballots; lane w funnel-shifts its slice out of it and deposits it on the
//     non-frozen positions of word w with a 5-step mask expand (host-precomputed move masks, Hacker's Delight 7-5);
//     the butterfly is 5 in-word shift/mask stages + log2(N/32) shuffle stages; 16 bits -> 16 bytes per vector store.
//   encode_kernel (any N <= 4096): same steps through shared memory.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "pb_sim.cuh"

namespace pb {

enum { ENC_POLAR = 0, ENC_CRC = 1, ENC_CRC_POLAR = 2 };

// per 32-bit word of the code: where its message bits sit in the compact string and how to spread them
struct EncWord {
    uint32_t mask;      // non-frozen positions of the word
    uint32_t kstart;    // index of the word's first message bit in the compact string
    uint32_t mv[5];     // move masks of expand(), step i shifts by 1 << i
    uint32_t pad;
};

__device__ __forceinline__ uint4 bits16_to_bytes(uint32_t bits) {
    uint4 v;
    v.x = ((bits & 15u) * 0x00204081u) & 0x01010101u;
    v.y = (((bits >> 4) & 15u) * 0x00204081u) & 0x01010101u;
    v.z = (((bits >> 8) & 15u) * 0x00204081u) & 0x01010101u;
    v.w = (((bits >> 12) & 15u) * 0x00204081u) & 0x01010101u;
    return v;
}

template <int MODE>
__global__ void __launch_bounds__(256)
encode_words_kernel(const SimDev s, const EncWord *__restrict__ tab, const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                    long long B, int vec_ok) {
    constexpr int kMaxQ = 8;                                 // 8 x 128 message bytes per frame at most (in_len <= 1024)
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int N = s.N, NW = N >> 5;
    const int in_len = (MODE == ENC_POLAR) ? s.K : s.A;
    const bool wide = vec_ok && (in_len & 3) == 0;          // message rows are 4-byte aligned: 4 bits per load
    const int n4 = in_len >> 2, nq = (in_len + 127) >> 7, n_cw = (in_len + 31) >> 5;
    // this lane's word description (zeros for lanes beyond the code)
    const uint4 e0 = __ldg(reinterpret_cast<const uint4 *>(tab + lane)), e1 = __ldg(reinterpret_cast<const uint4 *>(tab + lane) + 1);
    const long long stride = (long long)gridDim.x * wpb;
    long long f = (long long)blockIdx.x * wpb + wib;
    // the message of frame f+stride is fetched while frame f is encoded (one frame is only ~1.5 KB of traffic: without
    // the prefetch every warp would sit out a full DRAM round trip per frame)
    uint32_t v[kMaxQ], vn[kMaxQ];
    auto fetch = [&](long long fr, uint32_t (&dst)[kMaxQ]) {
        const uint32_t *src32 = reinterpret_cast<const uint32_t *>(in + (size_t)fr * in_len);
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) {
            const int i = lane + 32 * q;
            dst[q] = (fr < B && i < n4) ? __ldcs(src32 + i) : 0u;
        }
    };
    if (wide) fetch(f, v);
    for (; f < B; f += stride) {
        uint32_t cw = 0, crc = 0;
        if (wide) {
            fetch(f + stride, vn);
#pragma unroll
            for (int q = 0; q < kMaxQ; ++q) {
                if (q < nq) {
                    const uint32_t t = __vsetne4(v[q], 0u);                 // 0x01 per non-zero byte
                    if (MODE != ENC_POLAR) {
                        const int i = lane + 32 * q;
                        if (i < n4) {
                            const uint4 r = __ldg(reinterpret_cast<const uint4 *>(s.crc_rem) + i);
                            crc ^= ((t & 1u) ? r.x : 0u) ^ ((t & 0x100u) ? r.y : 0u) ^ ((t & 0x10000u) ? r.z : 0u) ^ ((t & 0x1000000u) ? r.w : 0u);
                        }
                    }
                    // 4 bits of this lane -> the 32-bit compact word of its 8-lane group -> the lane that owns that word
                    uint32_t g = ((t * 0x01020408u) >> 24) << (4 * (lane & 7));
                    g |= __shfl_xor_sync(0xffffffffu, g, 1);
                    g |= __shfl_xor_sync(0xffffffffu, g, 2);
                    g |= __shfl_xor_sync(0xffffffffu, g, 4);
                    const uint32_t mine = __shfl_sync(0xffffffffu, g, 8 * (lane & 3));
                    if ((lane >> 2) == q) cw = mine;
                }
            }
#pragma unroll
            for (int q = 0; q < kMaxQ; ++q) v[q] = vn[q];
        } else {
            const uint8_t *src = in + (size_t)f * in_len;
            for (int j = 0; j < n_cw; ++j) {
                const int k = 32 * j + lane;
                const bool bit = k < in_len && src[k] != 0;
                if (MODE != ENC_POLAR && bit) crc ^= __ldg(s.crc_rem + k);
                const uint32_t b = __ballot_sync(0xffffffffu, bit);
                if (lane == j) cw = b;
            }
        }
        if (MODE != ENC_POLAR) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) crc ^= __shfl_xor_sync(0xffffffffu, crc, o);
            // check bit i (MSB first) goes to compact position A + i
            const uint32_t rev = __brev(crc) >> (32 - s.crc_n);
            const int wa = s.A >> 5, sa = s.A & 31;
            if (lane == wa) cw |= rev << sa;
            if (sa != 0 && lane == wa + 1) cw |= rev >> (32 - sa);
        }
        // deposit: the word's popc(mask) message bits start at compact position kstart
        const int ks = (int)e0.y;
        const uint32_t lo = __shfl_sync(0xffffffffu, cw, ks >> 5), hi = __shfl_sync(0xffffffffu, cw, min(31, (ks >> 5) + 1));
        uint32_t x = __funnelshift_r(lo, hi, ks & 31);
        {
            uint32_t t;
            t = x << 16; x = (x & ~e1.z) | (t & e1.z);
            t = x << 8;  x = (x & ~e1.y) | (t & e1.y);
            t = x << 4;  x = (x & ~e1.x) | (t & e1.x);
            t = x << 2;  x = (x & ~e0.w) | (t & e0.w);
            t = x << 1;  x = (x & ~e0.z) | (t & e0.z);
            x &= e0.x;
        }
        // x = u F^{(x)n}
        x ^= (x >> 1) & 0x55555555u;
        x ^= (x >> 2) & 0x33333333u;
        x ^= (x >> 4) & 0x0f0f0f0fu;
        x ^= (x >> 8) & 0x00ff00ffu;
        x ^= (x >> 16) & 0x0000ffffu;
        for (int m = 1; m < NW; m <<= 1) {
            const uint32_t t = __shfl_xor_sync(0xffffffffu, x, m);
            if (!(lane & m)) x ^= t;
        }
        uint8_t *dst = out + (size_t)f * N;
        if (vec_ok) {
            for (int j = 0; 512 * j < N; ++j) {                  // 16-byte chunk h = bits 16h .. 16h+15 of the code
                const int h = lane + 32 * j;
                const uint32_t w = __shfl_sync(0xffffffffu, x, (lane >> 1) + 16 * j);
                if (16 * h < N) __stcs(reinterpret_cast<uint4 *>(dst) + h, bits16_to_bytes((w >> ((lane & 1) * 16)) & 0xffffu));
            }
        } else {
            for (int i = 0; i < NW; ++i) {
                const uint32_t w = __shfl_sync(0xffffffffu, x, i);
                dst[32 * i + lane] = (uint8_t)((w >> lane) & 1u);
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(256)
encode_kernel(const SimDev s, const uint8_t *__restrict__ in, uint8_t *__restrict__ out, long long B, int vec_ok) {
    extern __shared__ uint32_t sm_u[];                       // [warps][N/32] code bits
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int N = s.N, NW = N >> 5;
    const int in_len = (MODE == ENC_POLAR) ? s.K : s.A;
    uint32_t *U = sm_u + wib * NW;
    for (long long f = (long long)blockIdx.x * wpb + wib; f < B; f += (long long)gridDim.x * wpb) {
        const uint8_t *src = in + (size_t)f * in_len;
        if (MODE != ENC_CRC) {
            for (int w = lane; w < NW; w += 32) U[w] = 0;
            __syncwarp();
        }
        uint32_t crc = 0;
        auto take = [&](int k, uint32_t bit) {
            if (MODE == ENC_CRC) out[(size_t)f * s.K + k] = (uint8_t)bit;
            if (bit) {
                if (MODE != ENC_POLAR) crc ^= __ldg(s.crc_rem + k);
                if (MODE != ENC_CRC) {
                    const int pos = __ldg(s.info_pos + k);
                    atomicOr(&U[pos >> 5], 1u << (pos & 31));
                }
            }
        };
        if (vec_ok && (in_len & 3) == 0) {   // 4 message bits per load
            for (int k4 = lane; k4 < (in_len >> 2); k4 += 32) {
                const uint32_t v = __ldcs(reinterpret_cast<const uint32_t *>(src) + k4);
#pragma unroll
                for (int b = 0; b < 4; ++b) take(4 * k4 + b, ((v >> (8 * b)) & 0xffu) != 0);
            }
        } else {
            for (int k = lane; k < in_len; k += 32) take(k, src[k] != 0);
        }
        if (MODE != ENC_POLAR) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) crc ^= __shfl_xor_sync(0xffffffffu, crc, o);
            for (int k = lane; k < s.crc_n; k += 32) {
                const uint32_t bit = (crc >> (s.crc_n - 1 - k)) & 1u;
                if (MODE == ENC_CRC) out[(size_t)f * s.K + s.A + k] = (uint8_t)bit;
                else if (bit) {
                    const int pos = __ldg(s.info_pos + s.A + k);
                    atomicOr(&U[pos >> 5], 1u << (pos & 31));
                }
            }
        }
        if (MODE == ENC_CRC) continue;
        __syncwarp();
        for (int w = lane; w < NW; w += 32) {
            uint32_t x = U[w];
            x ^= (x >> 1) & 0x55555555u;
            x ^= (x >> 2) & 0x33333333u;
            x ^= (x >> 4) & 0x0f0f0f0fu;
            x ^= (x >> 8) & 0x00ff00ffu;
            x ^= (x >> 16) & 0x0000ffffu;
            U[w] = x;
        }
        __syncwarp();
        for (int m = 1; m < NW; m <<= 1) {
            for (int t = lane; t < NW / 2; t += 32) {
                const int w = ((t & ~(m - 1)) << 1) | (t & (m - 1));
                U[w] ^= U[w + m];
            }
            __syncwarp();
        }
        // one byte per code bit; 16 bits -> 16 bytes per lane and store when the layout allows
        uint8_t *dst = out + (size_t)f * N;
        if (vec_ok && (N & 15) == 0) {
            for (int h = lane; h < (N >> 4); h += 32)
                __stcs(reinterpret_cast<uint4 *>(dst) + h, bits16_to_bytes((U[h >> 1] >> ((h & 1) * 16)) & 0xffffu));
        } else {
            for (int p = lane; p < N; p += 32) dst[p] = (uint8_t)((U[p >> 5] >> (p & 31)) & 1u);
        }
        __syncwarp();
    }
}

// host: move masks of expand(x, m) = "deposit the low popc(m) bits of x on the set positions of m"
inline void enc_expand_masks(uint32_t m, uint32_t mv_out[5]) {
    uint32_t mk = ~m << 1;
    for (int i = 0; i < 5; ++i) {
        uint32_t mp = mk ^ (mk << 1);
        mp ^= mp << 2;
        mp ^= mp << 4;
        mp ^= mp << 8;
        mp ^= mp << 16;
        const uint32_t mv = mp & m;
        mv_out[i] = mv;
        m = (m ^ mv) | (mv >> (1 << i));
        mk &= ~mp;
    }
}

}  // namespace pb
