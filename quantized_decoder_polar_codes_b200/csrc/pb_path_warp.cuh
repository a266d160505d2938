// path_warp: warp-level, schedule-driven decoder for every class and every list size L in {1,2,4,8,16,32}.
//
// Same mapping as scl_lut_warp (one warp decodes 32/L frames, ONE LANE PER LIST PATH, lazy pointer copy,
// in-place bit partial sums) but generic in the value domain:
//   DOM_FLOAT / DOM_UNIFORM / DOM_LLOYD  fp64 LLRs, f/g evaluated with the reference's expressions
//                                         (PD/src/utils.cpp:8-60) -- rows a9, a11, a13, a15, a17, a20, a21 of SURVEY 8a
//   DOM_LUT                               byte symbols + tables read through the read-only cache -- the LUT shapes the
//                                         nibble kernel does not take (L = 16/32, tables wider than 16, per-position tables)
// It interprets the same compiled schedule as the CTA-per-frame generic kernel (pb::Step), so the two are
// interchangeable; this one keeps 32/L frames per warp in flight instead of one per CTA.
//
// Values of level d (1..n) live at  base[(voff[d] + j) * 32 + slot_lane]  -- lane-interleaved, so a warp access is
// one coalesced row whatever the slot permutation -- in an L2-resident global workspace for the large levels
// and in shared memory for the last three (4+2+1 elements).  Two 5-bit-per-level pointer words per lane say
// which physical slot holds level d of this logical path.
//
// 2L > 16 keys: libstdc++'s std::sort is an introsort whose order of EQUAL keys is algorithm defined (SURVEY
// App. B1).  For the plain SCL kinds a parallel stable rank is used when no two live (< 1e300) keys are equal --
// any correct sort then gives the same permutation of the live paths, and the order of dead (1e300 / inf) paths is
// unobservable.  Otherwise, and always for the Fast list kinds (their R1 rule reads the ordering held by the
// destination slot, dead or not), one lane per frame runs the exact emulation (pb::std_sort_idx).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "pb_generic.cuh"
#include "pb_internal.h"

namespace pb {

struct PathParams {
    int gl;            // value levels 1..gl live in the global workspace, gl+1..n in shared memory
    int gelems;        // workspace elements per lane (value levels)
    int velems;        // shared-memory elements per lane (value levels)
    int xwords;        // N/32
    int scrwords;      // epilogue scratch words per warp
    int r1_tmax;       // largest R1 node (list Fast kinds), 0 if none
    size_t r1_off;     // byte offset of the R1 scratch inside the per-CTA workspace
    int voff[kMaxLog + 2];
};

struct PathPlan {
    bool ok = false;
    int logL = 0;
    PathParams p{};
    size_t smem = 0;
    size_t ws_bytes_per_cta = 0;
    int ctas_per_sm = 0;
};

#ifndef PB_PATH_MINBLOCKS
#define PB_PATH_MINBLOCKS 20
#endif
template <int DOM, int LOGL>
__global__ void __launch_bounds__(32, PB_PATH_MINBLOCKS)
path_warp_kernel(const __grid_constant__ Dev d, const __grid_constant__ PathParams pp, const void *__restrict__ in, int in_dtype,
                 uint8_t *__restrict__ out, long long B, char *__restrict__ ws, int *err_flag, double *dbg_pm, int *dbg_win) {
    using T = typename Val<DOM>::T;
    constexpr int L = 1 << LOGL;
    constexpr int FPW = 32 / L;
    constexpr unsigned kAll = 0xffffffffu;
    PB_DYN_SMEM(unsigned char, smraw);
    const int lane = threadIdx.x;
    const int grp = lane >> LOGL, me = lane & (L - 1), gbase = lane & ~(L - 1);
    const int N = d.N, n = d.n;

    T *V = reinterpret_cast<T *>(smraw);                                           // levels gl+1..n, [elem][lane]
    uint32_t *X = reinterpret_cast<uint32_t *>(smraw + (size_t)pp.velems * 32 * sizeof(T));   // partial sums, in place
    uint32_t *SCR = X + pp.xwords * 32;
    double *KS = reinterpret_cast<double *>(SCR + pp.scrwords);                    // [32][2] fork keys
    int *ORD = reinterpret_cast<int *>(KS + 64);                                   // [64] sort order (2L > 16)
    uint32_t *SEL = reinterpret_cast<uint32_t *>(ORD + 64);                        // [32]
    char *wsc = ws + (size_t)blockIdx.x * (pp.r1_off + (size_t)32 * pp.r1_tmax * 12);
    T *G = reinterpret_cast<T *>(wsc);                                             // levels 1..gl, [elem][lane]
    double *R1K = reinterpret_cast<double *>(wsc + pp.r1_off) + (size_t)lane * pp.r1_tmax;            // per lane, contiguous
    int *R1I = reinterpret_cast<int *>(wsc + pp.r1_off + (size_t)32 * pp.r1_tmax * 8) + (size_t)lane * pp.r1_tmax;

    const long long n_groups = (B + FPW - 1) / FPW;
    for (long long g = blockIdx.x; g < n_groups; g += gridDim.x) {
        long long my_frame = g * FPW + grp;
        if (my_frame >= B) my_frame = B - 1;
        const uint8_t *in8 = reinterpret_cast<const uint8_t *>(in) + (size_t)my_frame * N;
        const int32_t *in32 = reinterpret_cast<const int32_t *>(in) + (size_t)my_frame * N;
        const double *in64 = reinterpret_cast<const double *>(in) + (size_t)my_frame * N;
        auto in0 = [&](int j) -> T {
            if (DOM == DOM_LUT) {
                int s = (in_dtype == 0) ? (int)__ldg(in8 + j) : __ldg(in32 + j);
                const int bound = (j < N / 2) ? d.root_qa : d.root_qb;
                if (s < 0 || s >= bound) { *err_flag = 1; s = 0; }
                return (T)s;
            } else {
                return (T)__ldg(in64 + j);
            }
        };

        double PM = (me == 0) ? 0.0 : d.pm_init;
        // 5-bit-per-level slot pointers, levels 1..n: values (pv) and left-child partial sums (pu)
        unsigned long long pv = 0x84210842108421ull * (unsigned long long)me, pu = pv;   // every 5-bit field = me
        auto getp = [&](unsigned long long pw, int lev) -> int { return (int)((pw >> (5 * (lev - 1))) & 31ull); };
        auto setown = [&](unsigned long long &pw, int lev) {
            pw = (pw & ~(31ull << (5 * (lev - 1)))) | ((unsigned long long)me << (5 * (lev - 1)));
        };
        auto vslot = [&](int lev) -> int { return L == 1 ? lane : (gbase | getp(pv, lev)); };
        auto uslot = [&](int lev) -> int { return L == 1 ? lane : (gbase | getp(pu, lev)); };
        auto level_ptr = [&](int lev, int sl) -> T * { return (lev <= pp.gl ? G : V) + (size_t)pp.voff[lev] * 32 + sl; };
        auto read_val = [&](int dd, const T *src, int j) -> T { return dd == 0 ? in0(j) : src[(size_t)j * 32]; };
        // LLR of element j of the node at (dd >= 1, node): float domains the value itself, LUT the table row of level dd-1
        auto elem_llr = [&](int dd, unsigned node, const T *src, int j) -> double {
            const T v = src[(size_t)j * 32];
            if (DOM == DOM_LUT) {
                const int pos = (int)(node << (n - dd)) + j;
                return __ldg(d.llr + __ldg(d.llr_off + (size_t)(dd - 1) * N + pos) + (int)v);
            } else {
                return (double)v;
            }
        };
        auto xbit_range_write = [&](int dd, unsigned node, auto word_of) {
            // word_of(w) -> the 32 result bits of word w of the node's range (or the low bits for a sub-word range)
            const int temp = N >> dd;
            const unsigned base = node * (unsigned)temp;
            uint32_t *xo = X + (base >> 5) * 32 + lane;
            if (temp >= 32) {
                for (int w = 0; w < (temp >> 5); ++w) xo[w * 32] = word_of(w);
            } else {
                const int sh = (int)(base & 31u);
                const uint32_t mask = ((1u << temp) - 1u) << sh;
                *xo = (*xo & ~mask) | ((word_of(0) << sh) & mask);
            }
            if (L > 1 && dd >= 1 && (node & 1u) == 0) setown(pu, dd);
            __syncwarp();
        };

        // mink + list permutation; see the header comment for the 2L > 16 case
        auto fork = [&](double K0, double K1, int &p, uint32_t &fl) {
            __syncwarp();
            *reinterpret_cast<double2 *>(&KS[lane * 2]) = make_double2(K0, K1);
            __syncwarp();
            int r0 = 0, r1 = 0;
            bool tie = false;
#pragma unroll 4
            for (int j = 0; j < L; ++j) {
                const double2 kf = *reinterpret_cast<const double2 *>(&KS[(gbase + j) * 2]);
                const bool jb = j < me;
                r0 += (jb ? !(K0 < kf.x) : (kf.x < K0)) + (kf.y < K0);
                r1 += !(K1 < kf.x) + (jb ? !(K1 < kf.y) : (kf.y < K1));
                if (2 * L > 16) {
                    tie |= (j != me && kf.x == K0 && K0 < 1e300) || (kf.y == K0 && K0 < 1e300);
                    tie |= (j != me && kf.y == K1 && K1 < 1e300);
                }
            }
            if (2 * L > 16) {
                // exact libstdc++ order is needed when two live keys are equal, and always for the Fast list kinds: their
                // R1 rule flips the bit named by the ordering the DESTINATION slot held before the permutation
                // (FastSCLDecoder.cpp:197-233), so which dead (PM = inf) path sits in which slot is observable there.
                // Decided for the whole warp so that the frames sharing it stay convergent.
                const bool need = d.r1_tmax > 0 || __any_sync(kAll, tie);
                if (need) {
                    if (me == 0) {
                        int *ord = ORD + gbase * 2;
                        // keys in PM2 order: index i < L keep of slot i, i >= L flip of slot i-L
                        double *k2 = KS + gbase * 2;   // interleaved (keep,flip) -> gather into a local copy
                        double keys[2 * L];
                        for (int i = 0; i < L; ++i) { keys[i] = k2[2 * i]; keys[L + i] = k2[2 * i + 1]; }
                        for (int i = 0; i < 2 * L; ++i) ord[i] = i;
                        std_sort_idx(ord, 2 * L, keys);
                    }
                    __syncwarp();
                    const int idx = ORD[gbase * 2 + me];
                    SEL[lane] = (uint32_t)(idx >= L ? (idx - L) | 32 : idx);
                    __syncwarp();
                } else {
                    if (r0 < L) SEL[gbase + r0] = (uint32_t)me;
                    if (r1 < L) SEL[gbase + r1] = (uint32_t)me | 32u;
                    __syncwarp();
                }
            } else {
                if (r0 < L) SEL[gbase + r0] = (uint32_t)me;
                if (r1 < L) SEL[gbase + r1] = (uint32_t)me | 32u;
                __syncwarp();
            }
            const uint32_t sv = SEL[lane];
            p = gbase | (int)(sv & 31u);
            fl = sv >> 5;
            PM = KS[p * 2 + fl];
            pv = __shfl_sync(kAll, pv, p);
            pu = __shfl_sync(kAll, pu, p);
        };

        double dmetric = 0.0;   // PD_BD_DMETRIC accumulator (one frame per lane there)
        Step st_next = d.steps[0];
        for (int si = 0; si < d.n_steps; ++si) {
            const Step st = st_next;
            if (si + 1 < d.n_steps) st_next = d.steps[si + 1];
            const int dd = st.depth;
            const unsigned node = st.node;
            const int temp = N >> dd;
            switch (st.op) {
            case OP_F:
            case OP_G: {
                const int ct = temp >> 1;
                const int p = (1 << dd) + (int)node - 1;
                const bool isg = st.op == OP_G;
                NodeTab tb;
                if (DOM == DOM_LUT) tb = d.tabs[p];
                double r = 0, M = 0;
                const double *bnd = nullptr, *rec = nullptr;
                if (DOM == DOM_UNIFORM) { r = isg ? d.r_g[p] : d.r_f[p]; M = (isg ? d.mg_mul : d.mf_mul) * r; }
                if (DOM == DOM_LLOYD) {
                    bnd = (isg ? d.bnd_g : d.bnd_f) + (size_t)p * d.nb;
                    rec = (isg ? d.rec_g : d.rec_f) + (size_t)p * d.nr;
                }
                const T *src = dd == 0 ? nullptr : level_ptr(dd, vslot(dd));
                T *dst = level_ptr(dd + 1, lane);
                const uint32_t *xsrc = X + uslot(dd + 1);
                const unsigned ub0 = (2u * node) * (unsigned)ct;
                auto one = [&](int j, T a, T b) {
                    int u = 0;
                    if (isg) { const unsigned bit = ub0 + (unsigned)j; u = (int)((xsrc[(bit >> 5) * 32] >> (bit & 31u)) & 1u); }
                    T o;
                    if (DOM == DOM_LUT) {
                        if (!isg) o = (T)__ldg(d.lut + tb.f_off + (size_t)j * tb.f_pstride + (unsigned)a * tb.f_qb + (unsigned)b);
                        else o = (T)__ldg(d.lut + tb.g_off + (size_t)j * tb.g_pstride + (unsigned)u * tb.g_sz + (unsigned)a * tb.g_qb + (unsigned)b);
                    } else {
                        double x = isg ? dev_g((double)a, (double)b, u) : dev_minsum((double)a, (double)b);
                        if (DOM == DOM_UNIFORM) x = dev_Q(x, r, M);
                        if (DOM == DOM_LLOYD) x = dev_bisect(x, bnd, d.nb, rec);
                        o = (T)x;
                    }
                    dst[(size_t)j * 32] = o;
                };
                int j = 0;
                for (; j + 4 <= ct; j += 4) {      // operands first (the big levels come from L2), then the math
                    T a[4], b[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) { a[q] = read_val(dd, src, j + q); b[q] = read_val(dd, src, j + q + ct); }
#pragma unroll
                    for (int q = 0; q < 4; ++q) one(j + q, a[q], b[q]);
                }
                for (; j < ct; ++j) one(j, read_val(dd, src, j), read_val(dd, src, j + ct));
                if (L > 1) setown(pv, dd + 1);
                __syncwarp();
                break;
            }
            case OP_C: {
                const int ct = temp >> 1;
                const unsigned lo = node * 2u * (unsigned)ct;
                const uint32_t *xl = X + uslot(dd + 1);
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
                if (L > 1 && dd >= 1 && (node & 1u) == 0) setown(pu, dd);
                __syncwarp();
                break;
            }
            case OP_LEAF: {
                uint32_t bit = 0;
                const T *src = level_ptr(n, vslot(n));
                if (L == 1) {
                    if (!st.flag) bit = elem_llr(n, node, src, 0) <= 0 ? 1u : 0u;   // PD/src/SCDecoder.cpp:31
                } else {
                    const double DM = elem_llr(n, node, src, 0);
                    if (st.flag) {
                        PM += fabs(DM) * (double)(DM < 0);                         // PD/src/SCLDecoder.cpp:62-66
                    } else {
                        const uint32_t dec = DM < 0 ? 1u : 0u;
                        int p;
                        uint32_t fl;
                        fork(PM, PM + fabs(DM), p, fl);
                        bit = __shfl_sync(kAll, dec, p) ^ fl;
                    }
                }
                uint32_t *xo = X + (node >> 5) * 32 + lane;
                *xo = (*xo & ~(1u << (node & 31u))) | (bit << (node & 31u));
                if (L > 1 && (node & 1u) == 0) setown(pu, n);
                __syncwarp();
                break;
            }
            case OP_R0: {
                if (L == 1 && d.bd == 1) {   // DMetric.cpp:56-61: DMetric += (sum of the node's LLRs) / temp
                    const T *src = level_ptr(dd, vslot(dd));
                    double tmp = 0;
                    for (int j = 0; j < temp; ++j) tmp += elem_llr(dd, node, src, j);
                    dmetric += tmp / temp;
                }
                if (L > 1) {   // PD/src/FastSCLDecoder.cpp:124-136
                    const T *src = level_ptr(dd, vslot(dd));
                    for (int j = 0; j < temp; ++j) {
                        const double l = elem_llr(dd, node, src, j);
                        PM += (double)(float)(l < 0) * fabs(l);
                    }
                }
                xbit_range_write(dd, node, [&](int) -> uint32_t { return 0u; });
                break;
            }
            case OP_REP: {
                const T *src = level_ptr(dd, vslot(dd));
                uint32_t fill;
                if (L == 1) {  // PD/src/FastSCDecoder.cpp:66-79
                    double S = 0;
                    for (int j = 0; j < temp; ++j) S += elem_llr(dd, node, src, j);
                    fill = S <= 0 ? 0xffffffffu : 0u;
                    if (d.bd == 1) dmetric += fabs(S) / temp;   // DMetric.cpp:87-93
                } else {       // PD/src/FastSCLDecoder.cpp:208-251
                    double a0 = PM, a1 = PM;
                    for (int j = 0; j < temp; ++j) {
                        const double l = elem_llr(dd, node, src, j);
                        a0 += (double)(l < 0) * fabs(l);
                        a1 += (double)(l >= 0) * fabs(l);
                    }
                    int p;
                    uint32_t fl;
                    fork(a0, a1, p, fl);
                    fill = fl ? 0xffffffffu : 0u;
                }
                xbit_range_write(dd, node, [&](int) -> uint32_t { return fill; });
                break;
            }
            case OP_SPC: {     // non-list only: PD/src/FastSCDecoder.cpp:80-106 (Wagner, first arg-min |llr|)
                const T *src = level_ptr(dd, vslot(dd));
                int parity = 0, amin = 0;
                double best = 0;
                for (int j = 0; j < temp; ++j) {
                    const double l = elem_llr(dd, node, src, j);
                    parity ^= (l <= 0) ? 1 : 0;
                    if (j == 0 || fabs(l) < best) { best = fabs(l); amin = j; }
                }
                xbit_range_write(dd, node, [&](int w) -> uint32_t {
                    uint32_t bits = 0;
                    const int cnt = temp < 32 ? temp : 32;
                    for (int k = 0; k < cnt; ++k) bits |= (elem_llr(dd, node, src, w * 32 + k) <= 0 ? 1u : 0u) << k;
                    if (parity && (amin >> 5) == w) bits ^= 1u << (amin & 31);
                    return bits;
                });
                break;
            }
            case OP_R1: {
                const T *src = level_ptr(dd, vslot(dd));
                if (L == 1) {  // PD/src/FastSCDecoder.cpp:54-65
                    xbit_range_write(dd, node, [&](int w) -> uint32_t {
                        uint32_t bits = 0;
                        const int cnt = temp < 32 ? temp : 32;
                        for (int k = 0; k < cnt; ++k) bits |= (elem_llr(dd, node, src, w * 32 + k) <= 0 ? 1u : 0u) << k;
                        return bits;
                    });
                    break;
                }
                // PD/src/FastSCLDecoder.cpp:139-205 incl. the flip-index quirk
                const unsigned base = node * (unsigned)temp;
                const int w0 = (int)(base >> 5), sh0 = (int)(base & 31u);
                const int nwords = temp >= 32 ? (temp >> 5) : 1;
                // hard decisions straight into the own X range; |llr| and the identity permutation into the lane's scratch
                {
                    uint32_t bits = 0;
                    for (int j = 0; j < temp; ++j) {
                        const double l = elem_llr(dd, node, src, j);
                        R1K[j] = fabs(l);
                        R1I[j] = j;
                        bits |= (l < 0 ? 1u : 0u) << (j & 31);
                        if ((j & 31) == 31 || j == temp - 1) {
                            uint32_t *xo = X + (w0 + (j >> 5)) * 32 + lane;
                            if (temp >= 32) *xo = bits;
                            else { const uint32_t mask = ((1u << temp) - 1u) << sh0; *xo = (*xo & ~mask) | ((bits << sh0) & mask); }
                            bits = 0;
                        }
                    }
                }
                std_sort_idx(R1I, temp, R1K);          // argsort(abs_llr), libstdc++ tie order, one path per lane
                __syncwarp();
                const int rounds = (L - 1 < temp) ? L - 1 : temp;
                int rowl = lane;                        // lane whose scratch rows (abs_llr, sorted idx) this path carries
                for (int layer = 0; layer < rounds; ++layer) {
                    const double *rk = R1K + ((long long)rowl - lane) * pp.r1_tmax;
                    const int *ri = R1I + ((long long)rowl - lane) * pp.r1_tmax;
                    const int q = ri[layer];
                    int p;
                    uint32_t fl;
                    fork(PM, PM + rk[q], p, fl);
                    // decision rows follow the parent; the flip goes to this slot's OWN pre-permutation position q
                    for (int w = 0; w < nwords; ++w) {
                        uint32_t v = X[(w0 + w) * 32 + p];
                        if (temp < 32) {
                            const uint32_t mask = ((1u << temp) - 1u) << sh0;
                            const uint32_t own = X[(w0 + w) * 32 + lane];
                            v = (own & ~mask) | (v & mask);
                        }
                        if (fl && (q >> 5) == w) v ^= 1u << ((q & 31) + (temp < 32 ? sh0 : 0));
                        __syncwarp();
                        X[(w0 + w) * 32 + lane] = v;
                        __syncwarp();
                    }
                    rowl = __shfl_sync(kAll, rowl, p);
                }
                if ((node & 1u) == 0) setown(pu, dd);
                __syncwarp();
                break;
            }
            default: break;
            }
        }

        // ---------------- epilogue: choose the path, u = x F^{(x)n}, gather ----------------
        const int NW = N >> 5;
        uint32_t *SC = (L == 1) ? X : SCR;
        auto transform_from = [&](int src_lane) {
            __syncwarp();
            for (int w = me; w < NW; w += L) {
                uint32_t x = X[w * 32 + src_lane];
                x ^= (x >> 1) & 0x55555555u;
                x ^= (x >> 2) & 0x33333333u;
                x ^= (x >> 4) & 0x0f0f0f0fu;
                x ^= (x >> 8) & 0x00ff00ffu;
                x ^= (x >> 16) & 0x0000ffffu;
                SC[w * FPW + grp] = x;
            }
            __syncwarp();
            for (int m = 1; m < NW; m <<= 1) {
                for (int t = me; t < NW / 2; t += L) {
                    const int w = ((t & ~(m - 1)) << 1) | (t & (m - 1));
                    SC[w * FPW + grp] ^= SC[(w + m) * FPW + grp];
                }
                __syncwarp();
            }
        };
        auto ubit = [&](int pos) -> uint32_t { return (SC[(pos >> 5) * FPW + grp] >> (pos & 31)) & 1u; };

        int winner = lane;
        double bd_pm = 0.0;
        bool bd_passed = false;
        if (L > 1) {
            __syncwarp();
            KS[lane] = PM;
            __syncwarp();
            int best = 0;
            double bk = KS[gbase];
            for (int j = 1; j < L; ++j) {
                const double kj = KS[gbase + j];
                if (kj < bk) { bk = kj; best = j; }     // std::min_element: first minimum
            }
            winner = gbase | best;
            if (d.ca) {   // candidates in argsort(PML) order (std::sort), first CRC pass wins; PD/src/CASCLDecoder.cpp:203-234
                if (me == 0) {
                    int *ord = ORD + gbase * 2;
                    for (int i = 0; i < L; ++i) ord[i] = i;
                    std_sort_idx(ord, L, KS + gbase);
                }
                __syncwarp();
                winner = gbase | ORD[gbase * 2];
                bd_pm = KS[gbase];                       // CASCLWithRNTI.cpp:205: PM = PML[0]
                bool decided = false;
                for (int t = 0; t < L; ++t) {
                    const int cand = gbase | ORD[gbase * 2 + t];
                    transform_from(cand);
                    bool pass = false;
                    if (me == 0) {   // CRC long division of the first A info bits, compare crc_check bits (utils.cpp:77-93)
                        uint32_t reg = 0;
                        const uint32_t msb = 1u << (d.crc_n - 1);
                        const uint32_t mask = (d.crc_n >= 32) ? 0xffffffffu : ((1u << d.crc_n) - 1u);
                        for (int k = 0; k < d.A; ++k) {
                            const uint32_t topb = ((reg & msb) ? 1u : 0u) ^ ubit(d.info_pos[k]);
                            reg = (reg << 1) & mask;
                            if (topb) reg ^= d.crc_taps;
                        }
                        pass = true;
                        for (int k = 0; k < d.crc_check; ++k) {
                            uint32_t bit = (reg >> (d.crc_n - 1 - k)) & 1u;
                            // CASCLWithRNTI.cpp:224-226: the RNTI is added onto the last RNTILength check bits
                            if (d.bd_rnti_len > 0 && k >= d.crc_n - d.bd_rnti_len) bit ^= (uint32_t)(d.bd_rnti[k - (d.crc_n - d.bd_rnti_len)] & 1);
                            if (bit != ubit(d.info_pos[d.A + k])) { pass = false; break; }
                        }
                    }
                    pass = __shfl_sync(kAll, (int)pass, gbase) != 0;
                    if (pass && !decided) { winner = cand; decided = true; bd_passed = true; bd_pm = KS[gbase + t]; }   // :236 PM = PML[i]
                    if (__all_sync(kAll, decided)) break;
                }
            }
        }
        transform_from(winner);
        {
            const long long frame = g * FPW + grp;
            if (frame < B) {
                if (out) {
                    uint8_t *o = out + (size_t)frame * d.Kout;
                    for (int k = me; k < d.Kout; k += L) o[k] = (uint8_t)ubit(d.info_pos[k]);
                }
                if (d.bd_metric && me == 0) d.bd_metric[frame] = d.bd == 1 ? dmetric : bd_pm;
                if (d.bd_pass && me == 0) d.bd_pass[frame] = (uint8_t)bd_passed;
                if (dbg_pm) dbg_pm[(size_t)frame * L + me] = L > 1 ? PM : 0.0;
                if (dbg_win && me == 0) dbg_win[frame] = winner - gbase;
            }
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
const void *path_kernel_fn(int dom, int logL);   // instantiations: pb_kernels.cu
#ifdef PB_TU_PATH
template <int DOM>
inline const void *path_kernel_fn_d(int logL) {
    switch (logL) {
    case 0: return PB_KFN(path_warp_kernel<DOM, 0>);
    case 1: return PB_KFN(path_warp_kernel<DOM, 1>);
    case 2: return PB_KFN(path_warp_kernel<DOM, 2>);
    case 3: return PB_KFN(path_warp_kernel<DOM, 3>);
    case 4: return PB_KFN(path_warp_kernel<DOM, 4>);
    default: return PB_KFN(path_warp_kernel<DOM, 5>);
    }
}
#endif

inline void plan_path_warp(const Dev &d, PathPlan *pl) {
    pl->ok = false;
    const int N = d.N, n = d.n, L = d.list ? d.L : 1;
    int logL = 0;
    while ((1 << logL) < L) logL++;
    if ((1 << logL) != L || L > 32) return;
    if (d.list && L == 1) return;   // a list decoder with L=1 keeps the list rules (bit = llr<0, REP tie -> all-zero): CTA kernel
    if (N < 32) return;      // sub-word codes stay on the CTA kernel
    PathParams &P = pl->p;
    P = PathParams{};
    const size_t sz = d.domain == DOM_LUT ? 1 : 8;
    int gl = 0, goff = 0, soff = 0;
    // levels of <= 4 elements (4+2+1) stay in shared memory, the rest lives in the workspace; measured against 8 and 2:
    // +10 % on the L=32 uniform and the Lloyd decoders, +3 % on float SCL (more resident warps), never worse
    const int smem_elems = getenv("POLAR_B200_PATH_SMEM_ELEMS") ? atoi(getenv("POLAR_B200_PATH_SMEM_ELEMS")) : 4;
    for (int lev = 1; lev <= n; ++lev) {
        const int elems = N >> lev;
        if (elems > smem_elems) { P.voff[lev] = goff; goff += elems; gl = lev; }
        else { P.voff[lev] = soff; soff += elems; }
    }
    P.gl = gl;
    P.gelems = std::max(goff, 1);
    P.velems = std::max(soff, 1);
    P.xwords = std::max(N / 32, 1);
    const int FPW = 32 / L;
    P.scrwords = (L == 1) ? 0 : ((P.xwords * FPW + 3) & ~3);   // keeps the fp64 key pairs behind it 16-byte aligned
    P.r1_tmax = d.list ? d.r1_tmax : 0;
    P.r1_off = (((size_t)P.gelems * 32 * sz) + 255) & ~(size_t)255;
    size_t vbytes = (((size_t)P.velems * 32 * sz) + 15) & ~(size_t)15;
    pl->smem = vbytes + (size_t)P.xwords * 32 * 4 + (size_t)P.scrwords * 4 + 64 * 8 + 64 * 4 + 32 * 4;
    pl->ws_bytes_per_cta = P.r1_off + (size_t)32 * P.r1_tmax * 12;
    pl->logL = logL;
    const void *fn = path_kernel_fn(d.domain, logL);
    if (pl->smem > 160 * 1024) return;
    // (per-function state shared by every decoder of this domain and list size: always the cap, see plan_generic)
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess) { cudaGetLastError(); return; }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 32, pl->smem) != cudaSuccess || occ < 1) { cudaGetLastError(); return; }
    pl->ctas_per_sm = occ;
    if (getenv("POLAR_B200_PATH_CTAS")) pl->ctas_per_sm = std::max(1, std::min(occ, atoi(getenv("POLAR_B200_PATH_CTAS"))));   // tuning knob
    pl->ok = true;
}

inline int path_grid(const PathPlan &pl, long long B, int sm_count) {
    const int L = 1 << pl.logL, FPW = 32 / L;
    long long groups = (B + FPW - 1) / FPW;
    return (int)std::max<long long>(1, std::min<long long>(groups, (long long)sm_count * pl.ctas_per_sm));
}

inline int launch_path_warp(const Dev &d, const PathPlan &pl, const void *d_in, int dtype, long long B, uint8_t *d_out,
                            cudaStream_t s, char *ws, int *d_err, double *dbg_pm, int *dbg_win, int sm_count) {
    const int grid = path_grid(pl, B, sm_count);
    void *args[] = {(void *)&d, (void *)&pl.p, (void *)&d_in, (void *)&dtype, (void *)&d_out, (void *)&B, (void *)&ws,
                    (void *)&d_err, (void *)&dbg_pm, (void *)&dbg_win};
    cudaError_t e = cudaLaunchKernel(path_kernel_fn(d.domain, pl.logL), dim3(grid), dim3(32), args, pl.smem, s);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaGetLastError();
}

}  // namespace pb
