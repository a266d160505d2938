// scl_lut_warp: the specialised kernel for the LUT SC / SCL / CRC-aided SCL decoders
// (SCLUTDecoder, SCLLUTDecoder, CASCLLUTDecoder -- SURVEY.md 8a rows a10, a14, a18; the north-star shape).
//
// Mapping: one warp decodes 32/L codewords at once, ONE LANE PER LIST PATH (L in {1,2,4,8}); the tree walk is
// the same for every frame, so the whole warp runs one uniform instruction stream and never diverges.
//   * symbols are 4-bit (table dims <= 16) and packed 8 per 32-bit word; level d of a path lives in shared
//     memory at V[(voff[d]+w)*32 + slot_lane]  (lane-interleaved: conflict-free for any slot permutation);
//   * partial sums are bits, stored IN PLACE in an N-bit array per slot: X[w*32 + slot_lane];
//   * list permutations copy nothing: two 4-bit-per-level pointer words per lane say which physical slot
//     holds level d of this logical path (values) / the left-child partial sums of level d; every f/g/combine
//     writes the lane's OWN slot and re-points that level to itself (lazy copy-on-write);
//   * f / g lookups: the 16x16 nibble table of the node is ONE 32-bit register per lane (128 B per warp) and a
//     lookup is a warp shuffle + shift (no bank conflicts, no shared memory);
//   * every table / LLR row is consumed exactly once per pass and in a fixed order, so the host lays them out
//     as one linear stream of 128-byte lines that each lane prefetches with cp.async into a private 16-slot
//     ring in shared memory (L2 latency fully hidden, no register cost);
//   * fork: 2L path metrics ranked with a stable (key,index) count -- identical to libstdc++'s insertion sort
//     for 2L <= 16 (PD/src/SCLLUTDecoder.cpp:8-22); keys are compared as the uint64 bit
//     patterns of the non-negative fp64 metrics.
// Output: the root partial sums x of the selected path are turned back into u = x F^{(x)n} and gathered.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <type_traits>
#include <vector>

#include "pb_internal.h"

namespace pb {

constexpr int kRing = 16;          // ring slots per lane; kRing-1 lines in flight
constexpr unsigned kFull = 0xffffffffu;

struct FastParams {
    const uint32_t *stream;        // table stream, 32 words per line
    int n_lines;
    const uint32_t *frozen_words;  // frozen mask, 32 leaves per word
    const uint32_t *crc_rem;       // [A] CRC remainder of each unit message bit (CA kinds)
    uint32_t crc_checkmask;
    int vwords, xwords, inwords;
    int voff[kMaxLog + 2];
};

struct FastPlan {
    bool ok = false;
    const char *name = "generic";
    int logL = 0;
    bool ca = false;
    FastParams p{};
    size_t smem = 0;
    int ctas_per_sm = 0;
    std::vector<void *> allocs;
};

__device__ __forceinline__ void cp_async4(uint32_t *smem_dst, const uint32_t *gsrc) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc));
    asm volatile("cp.async.commit_group;\n" ::);
}
template <int NPEND>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(NPEND) : "memory");
}

// 4 symbol bytes -> 4 nibbles (16 bits)
__device__ __forceinline__ uint32_t pack4(uint32_t x) {
    return (x & 0xfu) | ((x >> 4) & 0xf0u) | ((x >> 8) & 0xf00u) | ((x >> 12) & 0xf000u);
}

template <int LOGL, bool CA>
__global__ void __launch_bounds__(32)
scl_lut_warp_kernel(const Dev d, const FastParams fp, const void *__restrict__ in, int in_dtype,
                    uint8_t *__restrict__ out, long long B, int *err_flag, double *dbg_pm, int *dbg_win) {
    constexpr int L = 1 << LOGL;
    constexpr int FPW = 32 / L;
    extern __shared__ __align__(16) uint32_t sm[];
    const int lane = threadIdx.x;
    const int grp = lane >> LOGL, me = lane & (L - 1), gbase = lane & ~(L - 1);
    const int N = d.N, n = d.n;

    uint32_t *V = sm;
    uint32_t *X = V + fp.vwords * 32;
    uint32_t *IN = X + fp.xwords * 32;
    uint32_t *RING = IN + fp.inwords * FPW;
    unsigned long long *KS = reinterpret_cast<unsigned long long *>(RING + kRing * 32);  // [32][2]
    uint32_t *SEL = reinterpret_cast<uint32_t *>(KS + 64);                                // [32]

    // ---- table stream: private ring per lane (lane l only ever touches word l of a line) ----
    int fetch_pos = 0;      // stream position of the next line to prefetch
    unsigned line_no = 0;   // lines consumed so far (ring slot = line_no & 15)
    for (int i = 0; i < kRing - 1; ++i) {
        cp_async4(&RING[i * 32 + lane], fp.stream + (size_t)fetch_pos * 32 + lane);
        fetch_pos = (fetch_pos + 1 == fp.n_lines) ? 0 : fetch_pos + 1;
    }
    auto next_line = [&]() -> uint32_t {
        cp_async_wait<kRing - 2>();
        uint32_t v = RING[(line_no & (kRing - 1)) * 32 + lane];
        // refill the slot that was consumed one call ago
        cp_async4(&RING[((line_no + kRing - 1) & (kRing - 1)) * 32 + lane], fp.stream + (size_t)fetch_pos * 32 + lane);
        fetch_pos = (fetch_pos + 1 == fp.n_lines) ? 0 : fetch_pos + 1;
        ++line_no;
        return v;
    };
    auto lut16 = [&](uint32_t treg, uint32_t a, uint32_t b) -> uint32_t {
        uint32_t w = __shfl_sync(kFull, treg, a * 2 + (b >> 3));
        return (w >> ((b & 7) * 4)) & 15u;
    };

    const long long n_groups = (B + FPW - 1) / FPW;
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        // ---- stage the channel symbols of the FPW frames as nibbles: IN[w*FPW + frame_in_warp] ----
        for (int fi = 0; fi < FPW; ++fi) {
            long long frame = g * FPW + fi;
            if (frame >= B) frame = B - 1;
            for (int c = lane; c < N / 16; c += 32) {
                uint32_t x[4];
                bool bad = false;
                if (in_dtype == 0) {
                    uint4 v = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(in) + (size_t)frame * N) + c);
                    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
                } else {
                    const uint4 *p = reinterpret_cast<const uint4 *>(reinterpret_cast<const int32_t *>(in) + (size_t)frame * N) + c * 4;
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 v = __ldg(p + q);
                        bad |= ((v.x | v.y | v.z | v.w) & 0xffffff00u) != 0;
                        x[q] = (v.x & 0xff) | ((v.y & 0xff) << 8) | ((v.z & 0xff) << 16) | ((v.w & 0xff) << 24);
                    }
                }
                const uint32_t bound = (c * 16 < N / 2) ? (uint32_t)d.root_qa : (uint32_t)d.root_qb;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) bad |= ((x[q] >> (8 * bb)) & 0xffu) >= bound;
                }
                if (bad) { *err_flag = 1; x[0] = x[1] = x[2] = x[3] = 0; }
                IN[(2 * c) * FPW + fi] = pack4(x[0]) | (pack4(x[1]) << 16);
                IN[(2 * c + 1) * FPW + fi] = pack4(x[2]) | (pack4(x[3]) << 16);
            }
        }
        __syncwarp();

        double PM = (me == 0) ? 0.0 : d.pm_init;
        uint32_t pv_lo = 0x11111111u * me, pv_hi = 0x11111111u * me;   // value-level pointers, 4 bits per level
        uint32_t pu_lo = 0x11111111u * me, pu_hi = 0x11111111u * me;   // left-partial-sum pointers
        auto getp = [&](uint32_t lo, uint32_t hi, int lev) -> int {
            return (int)((lev <= 8 ? lo >> ((lev - 1) * 4) : hi >> ((lev - 9) * 4)) & 15u);
        };
        auto setown = [&](uint32_t &lo, uint32_t &hi, int lev) {
            if (lev <= 8) lo = (lo & ~(15u << ((lev - 1) * 4))) | ((uint32_t)me << ((lev - 1) * 4));
            else hi = (hi & ~(15u << ((lev - 9) * 4))) | ((uint32_t)me << ((lev - 9) * 4));
        };
        auto vslot = [&](int lev) -> int { return L == 1 ? lane : (gbase | getp(pv_lo, pv_hi, lev)); };
        auto uslot = [&](int lev) -> int { return L == 1 ? lane : (gbase | getp(pu_lo, pu_hi, lev)); };

        // f / g step at depth dd (< n-1): level dd -> level dd+1 of the child, written to the lane's own slot
        auto fg_step = [&](int dd, uint32_t node, auto isg_c) {
            constexpr bool ISG = decltype(isg_c)::value;
            const int ct = N >> (dd + 1);
            const uint32_t t0 = next_line();
            uint32_t t1 = 0;
            if (ISG) t1 = next_line();
            const uint32_t *src;
            int sstride;
            if (dd == 0) { src = IN + grp; sstride = FPW; }
            else { src = V + fp.voff[dd] * 32 + vslot(dd); sstride = 32; }
            uint32_t *dst = V + fp.voff[dd + 1] * 32 + lane;
            const uint32_t *xsrc = X + uslot(dd + 1);
            const uint32_t ub0 = (2u * node) * (uint32_t)ct;
            if (ct >= 8) {
                const int nw = ct >> 3;
                for (int w = 0; w < nw; ++w) {
                    const uint32_t A = src[w * sstride], Bv = src[(nw + w) * sstride];
                    uint32_t ub = 0;
                    if (ISG) { const uint32_t bit = ub0 + 8u * w; ub = xsrc[(bit >> 5) * 32] >> (bit & 31u); }
                    uint32_t o = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t a = (A >> (4 * k)) & 15u, b = (Bv >> (4 * k)) & 15u;
                        const uint32_t sl = a * 2 + (b >> 3);
                        uint32_t wv = __shfl_sync(kFull, t0, sl);
                        if (ISG) { const uint32_t wv1 = __shfl_sync(kFull, t1, sl); wv = ((ub >> k) & 1u) ? wv1 : wv; }
                        o |= ((wv >> ((b & 7u) * 4)) & 15u) << (4 * k);
                    }
                    dst[w * 32] = o;
                }
            } else {
                const uint32_t Wd = src[0];
                uint32_t ub = 0;
                if (ISG) ub = xsrc[(ub0 >> 5) * 32] >> (ub0 & 31u);
                uint32_t o = 0;
                for (int k = 0; k < ct; ++k) {
                    const uint32_t a = (Wd >> (4 * k)) & 15u, b = (Wd >> (4 * (k + ct))) & 15u;
                    const uint32_t sl = a * 2 + (b >> 3);
                    uint32_t wv = __shfl_sync(kFull, t0, sl);
                    if (ISG) { const uint32_t wv1 = __shfl_sync(kFull, t1, sl); wv = ((ub >> k) & 1u) ? wv1 : wv; }
                    o |= ((wv >> ((b & 7u) * 4)) & 15u) << (4 * k);
                }
                dst[0] = o;
            }
            if (L > 1) setown(pv_lo, pv_hi, dd + 1);
            __syncwarp();
        };

        // combine at depth dc: own[left range] = ptr-slot[left range] ^ own[right range]   (u(), utils.cpp:62-67)
        auto combine = [&](int dc, uint32_t node) {
            const int ct = N >> (dc + 1);
            const uint32_t lo = node * 2u * (uint32_t)ct;
            const uint32_t *xl = X + uslot(dc + 1);
            uint32_t *xo = X + lane;
            if (ct >= 32) {
                const int nw = ct >> 5, lw = (int)(lo >> 5);
                for (int w = 0; w < nw; ++w) {
                    const uint32_t v = xl[(lw + w) * 32] ^ xo[(lw + nw + w) * 32];
                    __syncwarp();
                    xo[(lw + w) * 32] = v;
                }
            } else {
                const int W = (int)(lo >> 5), sh = (int)(lo & 31u);
                const uint32_t lwv = xl[W * 32];
                uint32_t ow = xo[W * 32];
                const uint32_t mask = ((1u << ct) - 1u) << sh;
                ow = (ow & ~mask) | ((lwv ^ (ow >> ct)) & mask);
                __syncwarp();
                xo[W * 32] = ow;
            }
            if (L > 1 && dc >= 1 && (node & 1u) == 0) setown(pu_lo, pu_hi, dc);
            __syncwarp();
        };

        uint32_t frozen_w = 0;
        for (uint32_t phi = 0; phi < (uint32_t)N; ++phi) {
            if ((phi & 31u) == 0) frozen_w = __ldg(fp.frozen_words + (phi >> 5));
            const bool frozen = (frozen_w >> (phi & 31u)) & 1u;
            uint32_t sym;
            const uint32_t pnode = phi >> 1;
            if ((phi & 1u) == 0) {
                int dstart = 0;
                if (phi != 0) {
                    const int t = __ffs((int)phi) - 1;
                    const int dg = n - 1 - t;
                    fg_step(dg, phi >> (t + 1), std::true_type{});
                    dstart = dg + 1;
                }
                for (int dd = dstart; dd <= n - 2; ++dd) fg_step(dd, phi >> (n - dd), std::false_type{});
                const uint32_t Wd = V[fp.voff[n - 1] * 32 + vslot(n - 1)];
                const uint32_t t0 = next_line();
                sym = lut16(t0, Wd & 15u, (Wd >> 4) & 15u);
            } else {
                const uint32_t Wd = V[fp.voff[n - 1] * 32 + vslot(n - 1)];
                const uint32_t t0 = next_line(), t1 = next_line();
                const uint32_t u = (X[(phi >> 5) * 32 + uslot(n)] >> ((phi - 1u) & 31u)) & 1u;
                const uint32_t a = Wd & 15u, b = (Wd >> 4) & 15u;
                const uint32_t s0 = lut16(t0, a, b), s1 = lut16(t1, a, b);
                sym = u ? s1 : s0;
            }
            (void)pnode;
            uint32_t bit = 0;
            if (L == 1) {
                if (!frozen) {   // PD/src/SCLUTDecoder.cpp:59-67
                    const uint32_t lr = next_line();
                    const int lo_ = __shfl_sync(kFull, (int)lr, 2 * sym), hi_ = __shfl_sync(kFull, (int)lr, 2 * sym + 1);
                    const double DM = __hiloint2double(hi_, lo_);
                    bit = (DM <= 0) ? 1u : 0u;
                }
            } else {
                const uint32_t lr = next_line();
                const int lo_ = __shfl_sync(kFull, (int)lr, 2 * sym), hi_ = __shfl_sync(kFull, (int)lr, 2 * sym + 1);
                const double DM = __hiloint2double(hi_, lo_);
                if (frozen) {    // PD/src/SCLLUTDecoder.cpp:99-104
                    PM += fabs(DM) * (double)(DM < 0);
                } else {         // PD/src/SCLLUTDecoder.cpp:105-145
                    const uint32_t dec = (DM < 0) ? 1u : 0u;
                    const unsigned long long K0 = (unsigned long long)__double_as_longlong(PM);
                    const unsigned long long K1 = (unsigned long long)__double_as_longlong(PM + fabs(DM));
                    __syncwarp();
                    KS[lane * 2] = K0;
                    KS[lane * 2 + 1] = K1;
                    __syncwarp();
                    int r0 = 0, r1 = 0;
#pragma unroll
                    for (int j = 0; j < L; ++j) {
                        const ulonglong2 kf = *reinterpret_cast<const ulonglong2 *>(&KS[(gbase + j) * 2]);
                        r0 += (kf.x < K0) || (kf.x == K0 && j < me);
                        r0 += (kf.y < K0);
                        r1 += (kf.x <= K1);
                        r1 += (kf.y < K1) || (kf.y == K1 && j < me);
                    }
                    if (r0 < L) SEL[gbase + r0] = (uint32_t)me;
                    if (r1 < L) SEL[gbase + r1] = (uint32_t)me | 16u;
                    __syncwarp();
                    const uint32_t s = SEL[lane];
                    const int p = gbase | (int)(s & 15u);
                    const uint32_t fl = s >> 4;
                    PM = __longlong_as_double((long long)KS[p * 2 + fl]);
                    bit = __shfl_sync(kFull, dec, p) ^ fl;
                    pv_lo = __shfl_sync(kFull, pv_lo, p);
                    pu_lo = __shfl_sync(kFull, pu_lo, p);
                    if (n > 8) { pv_hi = __shfl_sync(kFull, pv_hi, p); pu_hi = __shfl_sync(kFull, pu_hi, p); }
                }
            }
            {   // leaf bit -> own slot (read-modify-write keeps the neighbouring ranges other paths may reference)
                uint32_t *xo = X + (phi >> 5) * 32 + lane;
                const uint32_t xw = *xo;
                *xo = (xw & ~(1u << (phi & 31u))) | (bit << (phi & 31u));
                if (L > 1 && (phi & 1u) == 0) setown(pu_lo, pu_hi, n);
                __syncwarp();
            }
            if (phi & 1u) {
                const int t1n = __ffs((int)~phi) - 1;   // trailing ones
                for (int k = 0; k < t1n; ++k) combine(n - 1 - k, phi >> (k + 1));
            }
        }

        // ---------------- epilogue: choose the path, u = x F^{(x)n}, gather the information bits ----------------
        const int NW = N >> 5;
        uint32_t *SCR = IN;   // level-0 buffer is dead now: SCR[w*FPW + grp]
        auto transform_from = [&](int src_lane) {
            __syncwarp();
            for (int w = me; w < NW; w += L) {
                uint32_t x = X[w * 32 + src_lane];
                x ^= (x >> 1) & 0x55555555u;
                x ^= (x >> 2) & 0x33333333u;
                x ^= (x >> 4) & 0x0f0f0f0fu;
                x ^= (x >> 8) & 0x00ff00ffu;
                x ^= (x >> 16) & 0x0000ffffu;
                SCR[w * FPW + grp] = x;
            }
            __syncwarp();
            for (int m = 1; m < NW; m <<= 1) {
                for (int t = me; t < NW / 2; t += L) {
                    const int w = ((t & ~(m - 1)) << 1) | (t & (m - 1));
                    SCR[w * FPW + grp] ^= SCR[(w + m) * FPW + grp];
                }
                __syncwarp();
            }
        };
        auto ubit = [&](int pos) -> uint32_t { return (SCR[(pos >> 5) * FPW + grp] >> (pos & 31)) & 1u; };

        int winner = lane;
        if (L > 1) {
            __syncwarp();
            KS[lane] = (unsigned long long)__double_as_longlong(PM);
            __syncwarp();
            const unsigned long long mine = KS[lane];
            int rank = 0, best = 0;
            unsigned long long bk = KS[gbase];
#pragma unroll
            for (int j = 0; j < L; ++j) {
                const unsigned long long kj = KS[gbase + j];
                rank += (kj < mine) || (kj == mine && j < me);
                if (kj < bk) { bk = kj; best = j; }   // std::min_element: first minimum
            }
            winner = gbase | best;
            if (CA) {   // PD/src/CASCLLUTDecoder.cpp:264-289: candidates in argsort(PML) order, first CRC pass wins
                SEL[gbase + rank] = (uint32_t)me;
                __syncwarp();
                winner = gbase | (int)SEL[gbase];
                bool decided = false;
                for (int t = 0; t < L; ++t) {
                    const int cand = gbase | (int)SEL[gbase + t];
                    transform_from(cand);
                    uint32_t acc = 0;
                    for (int k = me; k < d.A + d.crc_check; k += L) {
                        const uint32_t bitv = ubit(d.info_pos[k]);
                        const uint32_t contrib = (k < d.A) ? __ldg(fp.crc_rem + k) : (1u << (d.crc_n - 1 - (k - d.A)));
                        acc ^= bitv ? contrib : 0u;
                    }
#pragma unroll
                    for (int o = 1; o < L; o <<= 1) acc ^= __shfl_xor_sync(kFull, acc, o);
                    const bool pass = (acc & fp.crc_checkmask) == 0;
                    if (pass && !decided) { winner = cand; decided = true; }
                    if (__all_sync(kFull, decided)) break;
                }
            }
        }
        transform_from(winner);
        {
            long long frame = g * FPW + grp;
            if (frame < B) {
                uint8_t *o = out + (size_t)frame * d.Kout;
                for (int k = me; k < d.Kout; k += L) o[k] = (uint8_t)ubit(d.info_pos[k]);
                if (dbg_pm && L > 1) dbg_pm[(size_t)frame * L + me] = PM;
                if (dbg_pm && L == 1) dbg_pm[frame] = 0.0;
                if (dbg_win && me == 0) dbg_win[frame] = winner - gbase;
            }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// host side
inline void free_fast_plan(FastPlan *p) {
    for (void *q : p->allocs) cudaFree(q);
    p->allocs.clear();
    p->ok = false;
}

template <typename T>
inline bool fast_upload(FastPlan *pl, const std::vector<T> &h, const T **out) {
    void *p = nullptr;
    if (cudaMalloc(&p, std::max<size_t>(h.size() * sizeof(T), 16)) != cudaSuccess) return false;
    pl->allocs.push_back(p);
    if (!h.empty() && cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) return false;
    *out = reinterpret_cast<const T *>(p);
    return true;
}

inline const void *fast_kernel_fn(int logL, bool ca) {
    switch (logL) {
    case 0: return (const void *)scl_lut_warp_kernel<0, false>;
    case 1: return ca ? (const void *)scl_lut_warp_kernel<1, true> : (const void *)scl_lut_warp_kernel<1, false>;
    case 2: return ca ? (const void *)scl_lut_warp_kernel<2, true> : (const void *)scl_lut_warp_kernel<2, false>;
    default: return ca ? (const void *)scl_lut_warp_kernel<3, true> : (const void *)scl_lut_warp_kernel<3, false>;
    }
}

// Decide whether the specialised kernel applies and build its table stream (consumption order!).
// kind_* are the pd_kind values of the three eligible classes.
inline void plan_fast_lut(const Dev &d, int kind_sclut, int kind_scllut, int kind_cascllut,
                          const std::vector<NodeTab> &tabs, const std::vector<uint8_t> &pool,
                          const std::vector<double> &llr, const std::vector<uint32_t> &llr_off,
                          const int64_t *llr_off64, const int32_t *frozen, uint32_t crc_taps, FastPlan *pl) {
    pl->ok = false;
    if (d.kind != kind_sclut && d.kind != kind_scllut && d.kind != kind_cascllut) return;
    const int N = d.N, n = d.n, L = d.list ? d.L : 1;
    if (N < 32) return;
    int logL = 0;
    while ((1 << logL) < L) logL++;
    if ((1 << logL) != L || L > 8) return;
    for (int p = 0; p < N - 1; ++p) {
        const NodeTab &t = tabs[p];
        if (t.f_pstride || t.g_pstride) return;
        if (t.f_qb > 16 || t.g_qb > 16 || t.f_sz / t.f_qb > 16 || t.g_sz / t.g_qb > 16) return;
    }
    for (int leaf = 0; leaf < N; ++leaf) {
        int64_t r = (int64_t)(n - 1) * N + leaf;
        if (llr_off64[r + 1] - llr_off64[r] > 16) return;
    }
    const bool listk = L > 1;
    // ---- the stream, in the exact order scl_lut_warp_kernel consumes it ----
    std::vector<uint32_t> stream;
    auto push_table = [&](uint32_t off, int qa, int qb) {
        uint32_t line[32];
        memset(line, 0, sizeof line);
        for (int a = 0; a < qa; ++a)
            for (int b = 0; b < qb; ++b) {
                uint32_t v = pool[off + a * qb + b];
                line[a * 2 + (b >> 3)] |= (v & 15u) << ((b & 7) * 4);
            }
        stream.insert(stream.end(), line, line + 32);
    };
    auto push_f = [&](int p) { push_table(tabs[p].f_off, tabs[p].f_sz / tabs[p].f_qb, tabs[p].f_qb); };
    auto push_g = [&](int p) {
        push_table(tabs[p].g_off, tabs[p].g_sz / tabs[p].g_qb, tabs[p].g_qb);
        push_table(tabs[p].g_off + tabs[p].g_sz, tabs[p].g_sz / tabs[p].g_qb, tabs[p].g_qb);
    };
    auto push_llr = [&](int leaf) {
        uint32_t line[32];
        memset(line, 0, sizeof line);
        int64_t r = (int64_t)(n - 1) * N + leaf;
        int len = (int)(llr_off64[r + 1] - llr_off64[r]);
        memcpy(line, &llr[llr_off[r]], (size_t)len * sizeof(double));
        stream.insert(stream.end(), line, line + 32);
    };
    for (uint32_t phi = 0; phi < (uint32_t)N; ++phi) {
        if ((phi & 1u) == 0) {
            int dstart = 0;
            if (phi != 0) {
                int t = __builtin_ctz(phi);
                int dg = n - 1 - t;
                push_g((1 << dg) + (int)(phi >> (t + 1)) - 1);
                dstart = dg + 1;
            }
            for (int dd = dstart; dd <= n - 2; ++dd) push_f((1 << dd) + (int)(phi >> (n - dd)) - 1);
            push_f((1 << (n - 1)) + (int)(phi >> 1) - 1);
        } else {
            push_g((1 << (n - 1)) + (int)(phi >> 1) - 1);
        }
        if (listk || frozen[phi] != 1) push_llr((int)phi);
    }
    FastParams &P = pl->p;
    P = FastParams{};
    P.n_lines = (int)(stream.size() / 32);
    if (P.n_lines < kRing) return;
    std::vector<uint32_t> fw((N + 31) / 32, 0);
    for (int i = 0; i < N; ++i) if (frozen[i] == 1) fw[i >> 5] |= 1u << (i & 31);
    if (!fast_upload(pl, stream, &P.stream) || !fast_upload(pl, fw, &P.frozen_words)) { free_fast_plan(pl); return; }
    if (d.ca) {
        // remainder of the unit message e_k under the reference's long division (utils.cpp:77-93): linear, so the
        // CRC of a word is the XOR of the remainders of its set bits
        std::vector<uint32_t> rem(d.A);
        const uint32_t msb = 1u << (d.crc_n - 1), mask = d.crc_n >= 32 ? 0xffffffffu : ((1u << d.crc_n) - 1u);
        for (int k = 0; k < d.A; ++k) {
            uint32_t reg = 0;
            for (int i = k; i < d.A; ++i) {
                uint32_t top = ((reg & msb) ? 1u : 0u) ^ (i == k ? 1u : 0u);
                reg = (reg << 1) & mask;
                if (top) reg ^= crc_taps;
            }
            rem[k] = reg;
        }
        if (!fast_upload(pl, rem, &P.crc_rem)) { free_fast_plan(pl); return; }
        uint32_t cm = 0;
        for (int k = 0; k < d.crc_check; ++k) cm |= 1u << (d.crc_n - 1 - k);
        P.crc_checkmask = cm;
    }
    // level offsets (words) of the nibble-packed value levels 1..n-1
    int off = 0;
    for (int lev = 1; lev <= n - 1; ++lev) {
        P.voff[lev] = off;
        off += std::max(1, (N >> lev) / 8);
    }
    P.vwords = off;
    P.xwords = N / 32;
    P.inwords = N / 8;
    const int FPW = 32 / L;
    size_t words = (size_t)P.vwords * 32 + (size_t)P.xwords * 32 + (size_t)P.inwords * FPW + kRing * 32 + 128 + 32;
    pl->smem = words * 4;
    pl->logL = logL;
    pl->ca = d.ca != 0;
    const void *fn = fast_kernel_fn(logL, pl->ca);
    if (pl->smem > 200 * 1024) { free_fast_plan(pl); return; }
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem) != cudaSuccess) { cudaGetLastError(); free_fast_plan(pl); return; }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 32, pl->smem) != cudaSuccess || occ < 1) { cudaGetLastError(); free_fast_plan(pl); return; }
    pl->ctas_per_sm = occ;
    pl->name = "scl_lut_warp";
    pl->ok = true;
}

inline int launch_fast_lut(const Dev &d, const FastPlan &pl, const void *d_in, int dtype, long long B, uint8_t *d_out,
                           cudaStream_t s, int *d_err, double *dbg_pm, int *dbg_win, int sm_count) {
    const int L = 1 << pl.logL, FPW = 32 / L;
    long long groups = (B + FPW - 1) / FPW;
    int grid = (int)std::min<long long>(groups, (long long)sm_count * pl.ctas_per_sm);
#define PB_LAUNCH(LOGL, CAF) scl_lut_warp_kernel<LOGL, CAF><<<grid, 32, pl.smem, s>>>(d, pl.p, d_in, dtype, d_out, B, d_err, dbg_pm, dbg_win)
    switch (pl.logL) {
    case 0: PB_LAUNCH(0, false); break;
    case 1: if (pl.ca) PB_LAUNCH(1, true); else PB_LAUNCH(1, false); break;
    case 2: if (pl.ca) PB_LAUNCH(2, true); else PB_LAUNCH(2, false); break;
    default: if (pl.ca) PB_LAUNCH(3, true); else PB_LAUNCH(3, false); break;
    }
#undef PB_LAUNCH
    return (int)cudaGetLastError();
}

}  // namespace pb
