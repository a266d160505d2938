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
    const uint32_t *stream;        // upper-level table stream, 32 words per line
    const uint32_t *bundles;       // [N/8][8][32 lanes][4] words: the 29 lines of each 8-leaf subtree, lane-transposed
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
scl_lut_warp_kernel(const __grid_constant__ Dev d, const __grid_constant__ FastParams fp, const void *__restrict__ in, int in_dtype,
                    uint8_t *__restrict__ out, long long B, int *err_flag, double *dbg_pm, int *dbg_win) {
    constexpr int L = 1 << LOGL;
    constexpr int FPW = 32 / L;
    extern __shared__ __align__(16) uint32_t sm[];
    const int lane = threadIdx.x;
    const int grp = lane >> LOGL, me = lane & (L - 1), gbase = lane & ~(L - 1);
    const int N = d.N, n = d.n;
    const int NS = N >> 3;          // number of 8-leaf subtrees
    const int top = n - 3;          // depth of the subtree roots = deepest level kept in shared memory

    uint32_t *V = sm;
    uint32_t *X = V + fp.vwords * 32;
    uint32_t *IN = X + fp.xwords * 32;
    uint32_t *RING = IN + fp.inwords * FPW;
    double *KS = reinterpret_cast<double *>(RING + kRing * 32);      // [32][2]
    uint32_t *SEL = reinterpret_cast<uint32_t *>(KS + 64);           // [32]

    // ---- upper-level table stream: private ring per lane (lane l only ever touches word l of a line) ----
    int fetch_pos = 0;      // stream position of the next line to prefetch
    unsigned line_no = 0;   // lines consumed so far (ring slot = line_no & 15)
    for (int i = 0; i < kRing - 1; ++i) {
        cp_async4(&RING[i * 32 + lane], fp.stream + (size_t)fetch_pos * 32 + lane);
        fetch_pos = (fetch_pos + 1 == fp.n_lines) ? 0 : fetch_pos + 1;
    }
    auto next_line = [&]() -> uint32_t {
        cp_async_wait<kRing - 2>();
        uint32_t v = RING[(line_no & (kRing - 1)) * 32 + lane];
        cp_async4(&RING[((line_no + kRing - 1) & (kRing - 1)) * 32 + lane], fp.stream + (size_t)fetch_pos * 32 + lane);
        fetch_pos = (fetch_pos + 1 == fp.n_lines) ? 0 : fetch_pos + 1;
        ++line_no;
        return v;
    };
    // ---- bottom-subtree bundles: 32 lines (29 used) per 8-leaf subtree, lane-transposed so that 8 LDG.128 per
    //      lane fetch the lane's word of every line; double buffered in registers one subtree ahead ----
    auto load_bundle = [&](uint32_t (&tb)[32], int sidx) {
        const uint4 *src = reinterpret_cast<const uint4 *>(fp.bundles) + (size_t)sidx * 256 + lane;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint4 v = __ldg(src + j * 32);
            tb[4 * j] = v.x; tb[4 * j + 1] = v.y; tb[4 * j + 2] = v.z; tb[4 * j + 3] = v.w;
        }
    };
    auto lut16 = [&](uint32_t treg, uint32_t a, uint32_t b) -> uint32_t {
        const uint32_t w = __shfl_sync(kFull, treg, a * 2 + (b >> 3));
        return (w >> ((b & 7u) * 4)) & 15u;
    };
    auto nib = [](uint32_t w, int k) -> uint32_t { return (w >> (4 * k)) & 15u; };

    uint32_t bufA[32], bufB[32];
    load_bundle(bufA, 0);

    const unsigned gmask_below = (L == 32 ? kFull : (((1u << L) - 1u) << gbase)) & ((1u << lane) - 1u);

    const long long n_groups = (B + FPW - 1) / FPW;
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        // ---- stage the channel symbols of the FPW frames as nibbles: IN[w*FPW + frame_in_warp] ----
        for (int fi = 0; fi < FPW; ++fi) {
            long long frame = g * FPW + fi;
            if (frame >= B) frame = B - 1;
            for (int c = lane; c < N / 16; c += 32) {
                uint32_t x0, x1, x2, x3;
                bool bad = false;
                if (in_dtype == 0) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(in) + (size_t)frame * N) + c);
                    x0 = v.x; x1 = v.y; x2 = v.z; x3 = v.w;
                } else {
                    const uint4 *p = reinterpret_cast<const uint4 *>(reinterpret_cast<const int32_t *>(in) + (size_t)frame * N) + c * 4;
                    auto ld4 = [&](int q) -> uint32_t {
                        const uint4 v = __ldg(p + q);
                        bad |= ((v.x | v.y | v.z | v.w) & 0xffffff00u) != 0;
                        return (v.x & 0xff) | ((v.y & 0xff) << 8) | ((v.z & 0xff) << 16) | ((v.w & 0xff) << 24);
                    };
                    x0 = ld4(0); x1 = ld4(1); x2 = ld4(2); x3 = ld4(3);
                }
                const uint32_t bound = (c * 16 < N / 2) ? (uint32_t)d.root_qa : (uint32_t)d.root_qb;
                auto chk = [&](uint32_t x) {
#pragma unroll
                    for (int bb = 0; bb < 4; ++bb) bad |= ((x >> (8 * bb)) & 0xffu) >= bound;
                };
                chk(x0); chk(x1); chk(x2); chk(x3);
                if (bad) { *err_flag = 1; x0 = x1 = x2 = x3 = 0; }
                IN[(2 * c) * FPW + fi] = pack4(x0) | (pack4(x1) << 16);
                IN[(2 * c + 1) * FPW + fi] = pack4(x2) | (pack4(x3) << 16);
            }
        }
        __syncwarp();

        double PM = (me == 0) ? 0.0 : d.pm_init;
        // 3-bit-per-level slot pointers for levels 1..top: values (pv) and left-child partial sums (pu)
        uint32_t pv = 0x09249249u * (uint32_t)me, pu = pv;
        auto getp = [&](uint32_t pw, int lev) -> int { return (int)((pw >> (3 * (lev - 1))) & 7u); };
        auto setown = [&](uint32_t &pw, int lev) { pw = (pw & ~(7u << (3 * (lev - 1)))) | ((uint32_t)me << (3 * (lev - 1))); };
        auto vslot = [&](int lev) -> int { return L == 1 ? lane : (gbase | getp(pv, lev)); };
        auto uslot = [&](int lev) -> int { return L == 1 ? lane : (gbase | getp(pu, lev)); };

        // f / g step at depth dd (<= top-1): level dd -> level dd+1 (>= 8 symbols), written to the lane's own slot
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
            if (L > 1) setown(pv, dd + 1);
            __syncwarp();
        };

        // combine at depth dc (<= top-1): own[left range] = ptr-slot[left range] ^ own[right range]  (u(), utils.cpp:62-67)
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
            if (L > 1 && dc >= 1 && (node & 1u) == 0) setown(pu, dc);
            __syncwarp();
        };

        // ---- by-value state of the 8-leaf subtree being decoded (shuffled wholesale on a fork) ----
        uint32_t w3 = 0;     // 8 symbols entering the subtree root (depth top)
        uint32_t w21 = 0;    // [15:0] 4 symbols of the depth top+1 node, [23:16] 2 symbols of the depth top+2 node
        uint32_t xb = 0;     // partial sums of the subtree, in place (bit i = leaf i)

        // leaf decision (PD/src/SCLUTDecoder.cpp:59-67 / SCLLUTDecoder.cpp:92-145); sets bit `pos` of xb
        auto leaf = [&](uint32_t phi, int pos, uint32_t sym, uint32_t lr, bool frozen) {
            uint32_t bit = 0;
            if (L == 1) {
                if (!frozen) {
                    const int lo_ = __shfl_sync(kFull, (int)lr, 2 * sym), hi_ = __shfl_sync(kFull, (int)lr, 2 * sym + 1);
                    bit = (__hiloint2double(hi_, lo_) <= 0) ? 1u : 0u;
                }
            } else {
                const int lo_ = __shfl_sync(kFull, (int)lr, 2 * sym), hi_ = __shfl_sync(kFull, (int)lr, 2 * sym + 1);
                const double DM = __hiloint2double(hi_, lo_);
                if (frozen) {
                    PM += fabs(DM) * (double)(DM < 0);
                } else {
                    const uint32_t dec = (DM < 0) ? 1u : 0u;
                    const double K0 = PM, K1 = PM + fabs(DM);
                    __syncwarp();
                    *reinterpret_cast<double2 *>(&KS[lane * 2]) = make_double2(K0, K1);
                    // stable rank of (key, index) among the group's 2L keys == libstdc++ insertion sort for 2L <= 16
                    int r0 = __popc(__match_any_sync(kFull, (unsigned long long)__double_as_longlong(K0)) & gmask_below);
                    int r1 = __popc(__match_any_sync(kFull, (unsigned long long)__double_as_longlong(K1)) & gmask_below);
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < L; ++j) {
                        const double2 kf = *reinterpret_cast<const double2 *>(&KS[(gbase + j) * 2]);
                        r0 += (kf.x < K0) + (kf.y < K0);
                        r1 += !(K1 < kf.x) + (kf.y < K1);
                    }
                    if (r0 < L) SEL[gbase + r0] = (uint32_t)me;
                    if (r1 < L) SEL[gbase + r1] = (uint32_t)me | 16u;
                    __syncwarp();
                    const uint32_t s = SEL[lane];
                    const int p = gbase | (int)(s & 15u);
                    const uint32_t fl = s >> 4;
                    PM = KS[p * 2 + fl];
                    bit = __shfl_sync(kFull, dec, p) ^ fl;
                    w3 = __shfl_sync(kFull, w3, p);
                    w21 = __shfl_sync(kFull, w21, p);
                    xb = __shfl_sync(kFull, xb, p);
                    pv = __shfl_sync(kFull, pv, p);
                    pu = __shfl_sync(kFull, pu, p);
                }
            }
            xb = (xb & ~(1u << pos)) | (bit << pos);
            (void)phi;
        };

        // one 8-leaf subtree, fully unrolled; tb = its bundle (line order = consumption order, see plan_fast_lut)
        auto subtree8 = [&](int sidx, const uint32_t (&tb)[32]) {
            const uint32_t fz = (__ldg(fp.frozen_words + (sidx >> 2)) >> ((sidx & 3) * 8)) & 0xffu;
            w3 = V[fp.voff[top] * 32 + vslot(top)];
            xb = 0;
            // A.f : 8 -> 4 symbols
            uint32_t w2 = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) w2 |= lut16(tb[0], nib(w3, k), nib(w3, k + 4)) << (4 * k);
            w21 = w2;
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int bl = 1 + 15 * m;
                // B.f : 4 -> 2 symbols
                {
                    const uint32_t c2 = w21 & 0xffffu;
                    const uint32_t w1 = lut16(tb[bl], nib(c2, 0), nib(c2, 2)) | (lut16(tb[bl], nib(c2, 1), nib(c2, 3)) << 4);
                    w21 = c2 | (w1 << 16);
                }
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int cl = bl + 1 + 7 * c, pos = 4 * m + 2 * c;
                    const uint32_t phi = 8u * sidx + pos;
                    {   // left leaf through the f table of the depth n-1 node
                        const uint32_t w1 = w21 >> 16;
                        const uint32_t sym = lut16(tb[cl], nib(w1, 0), nib(w1, 1));
                        leaf(phi, pos, sym, tb[cl + 1], (fz >> pos) & 1u);
                    }
                    {   // right leaf through the g tables, u = left leaf's bit
                        const uint32_t w1 = w21 >> 16;
                        const uint32_t a = nib(w1, 0), b = nib(w1, 1);
                        const uint32_t s0 = lut16(tb[cl + 2], a, b), s1 = lut16(tb[cl + 3], a, b);
                        const uint32_t sym = ((xb >> pos) & 1u) ? s1 : s0;
                        leaf(phi + 1, pos + 1, sym, tb[cl + 4], (fz >> (pos + 1)) & 1u);
                    }
                    xb ^= ((xb >> (pos + 1)) & 1u) << pos;    // combine of the depth n-1 node
                    if (c == 0) {   // B.g : u = the 2 partial sums just formed
                        const uint32_t c2 = w21 & 0xffffu;
                        uint32_t w1 = 0;
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const uint32_t a = nib(c2, k), b = nib(c2, k + 2);
                            const uint32_t s0 = lut16(tb[bl + 6], a, b), s1 = lut16(tb[bl + 7], a, b);
                            w1 |= (((xb >> (pos + k)) & 1u) ? s1 : s0) << (4 * k);
                        }
                        w21 = c2 | (w1 << 16);
                    }
                }
                xb ^= ((xb >> (4 * m + 2)) & 3u) << (4 * m);    // combine of the depth n-2 node
                if (m == 0) {   // A.g : u = the 4 partial sums of the left half
                    uint32_t w2g = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t a = nib(w3, k), b = nib(w3, k + 4);
                        const uint32_t s0 = lut16(tb[14], a, b), s1 = lut16(tb[15], a, b);
                        w2g |= (((xb >> k) & 1u) ? s1 : s0) << (4 * k);
                    }
                    w21 = w2g;
                }
            }
            xb ^= (xb >> 4) & 15u;                               // combine of the subtree root
            // publish the 8 partial sums in the lane's own slot (read-modify-write keeps neighbouring ranges)
            uint32_t *xo = X + ((8u * sidx) >> 5) * 32 + lane;
            const int sh = (8 * sidx) & 31;
            *xo = (*xo & ~(0xffu << sh)) | ((xb & 0xffu) << sh);
            if (L > 1 && (sidx & 1) == 0) setown(pu, top);
            __syncwarp();
        };

        auto iteration = [&](int s, const uint32_t (&cur)[32], uint32_t (&nxt)[32]) {
            load_bundle(nxt, (s + 1 == NS) ? 0 : s + 1);   // next subtree (wraps into the next pass)
            int dstart = 0;
            if (s != 0) {
                const int t = __ffs(s) - 1;
                const int dg = top - 1 - t;
                fg_step(dg, (uint32_t)s >> (t + 1), std::true_type{});
                dstart = dg + 1;
            }
            for (int dd = dstart; dd <= top - 1; ++dd) fg_step(dd, (uint32_t)s >> (top - dd), std::false_type{});
            subtree8(s, cur);
            const int t1n = __ffs(~s) - 1;   // trailing ones of s
            for (int k = 0; k < t1n; ++k) combine(top - 1 - k, (uint32_t)s >> (k + 1));
        };
        for (int s = 0; s < NS; s += 2) {
            iteration(s, bufA, bufB);
            iteration(s + 1, bufB, bufA);
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
            KS[lane] = PM;
            __syncwarp();
            const double mine = KS[lane];
            int rank = 0, best = 0;
            double bk = KS[gbase];
#pragma unroll
            for (int j = 0; j < L; ++j) {
                const double kj = KS[gbase + j];
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
    // (a) upper levels (depths 0..n-4), in the order iteration() issues its f/g steps
    const int top = n - 3, NS = N >> 3;
    for (int sidx = 0; sidx < NS; ++sidx) {
        int dstart = 0;
        if (sidx != 0) {
            int t = __builtin_ctz((unsigned)sidx);
            int dg = top - 1 - t;
            push_g((1 << dg) + (sidx >> (t + 1)) - 1);
            dstart = dg + 1;
        }
        for (int dd = dstart; dd <= top - 1; ++dd) push_f((1 << dd) + (sidx >> (top - dd)) - 1);
    }
    std::vector<uint32_t> upper;
    upper.swap(stream);
    // (b) one bundle of 29 (+3 pad) lines per 8-leaf subtree, in the order subtree8() indexes them:
    //     0 A.f | per half m: 1+15m B.f, per pair c: +1+7c C.f, llr(left), C.g0, C.g1, llr(right); +6,+7 B.g0,B.g1 | 14,15 A.g0,A.g1
    std::vector<uint32_t> bundles((size_t)NS * 32 * 32, 0);
    for (int sidx = 0; sidx < NS; ++sidx) {
        stream.clear();
        auto heap = [&](int depth, int node) { return (1 << depth) + node - 1; };
        const int A = heap(top, sidx);
        std::vector<uint32_t> lines((size_t)32 * 32, 0);
        auto put = [&](int line) {   // moves the most recently pushed line(s) of `stream` into `lines`
            memcpy(&lines[(size_t)line * 32], &stream[stream.size() - 32], 32 * sizeof(uint32_t));
        };
        auto put2 = [&](int line) {
            memcpy(&lines[(size_t)line * 32], &stream[stream.size() - 64], 64 * sizeof(uint32_t));
        };
        push_f(A); put(0);
        push_g(A); put2(14);
        for (int m = 0; m < 2; ++m) {
            const int Bn = heap(top + 1, 2 * sidx + m), bl = 1 + 15 * m;
            push_f(Bn); put(bl);
            push_g(Bn); put2(bl + 6);
            for (int c = 0; c < 2; ++c) {
                const int Cn = heap(top + 2, 4 * sidx + 2 * m + c), cl = bl + 1 + 7 * c;
                const int leaf0 = 8 * sidx + 4 * m + 2 * c;
                push_f(Cn); put(cl);
                push_llr(leaf0); put(cl + 1);
                push_g(Cn); put2(cl + 2);
                push_llr(leaf0 + 1); put(cl + 4);
            }
        }
        // lane-transposed: chunk j (lines 4j..4j+3), lane l -> 4 consecutive words
        uint32_t *dst = &bundles[(size_t)sidx * 1024];
        for (int j = 0; j < 8; ++j)
            for (int l = 0; l < 32; ++l)
                for (int q = 0; q < 4; ++q) dst[(j * 32 + l) * 4 + q] = lines[(size_t)(4 * j + q) * 32 + l];
    }
    stream.swap(upper);
    FastParams &P = pl->p;
    P = FastParams{};
    P.n_lines = (int)(stream.size() / 32);
    if (P.n_lines < 1) return;
    std::vector<uint32_t> fw((N + 31) / 32, 0);
    for (int i = 0; i < N; ++i) if (frozen[i] == 1) fw[i >> 5] |= 1u << (i & 31);
    if (!fast_upload(pl, stream, &P.stream) || !fast_upload(pl, bundles, &P.bundles) || !fast_upload(pl, fw, &P.frozen_words)) { free_fast_plan(pl); return; }
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
    for (int lev = 1; lev <= n - 3; ++lev) {
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
