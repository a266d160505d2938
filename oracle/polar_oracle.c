/*
 * polar_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference decoders' decode() path.
 *
 * This file is the parity oracle for the CUDA decoders in quantized_decoder_polar_codes_b200/csrc.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it;
 * the product path never does (it fails loudly when the CUDA library is missing).
 *
 * Parity pinning: the reference ships no tests / golden vectors, so this
 * restatement is pinned against the reference ITSELF, compiled unmodified into oracle/_ref by
 * oracle/Makefile and exercised by tests/test_oracle_vs_ref.py and the fixtures in tests/golden/
 * (generated from oracle/_ref by tests/golden/make_golden.py).
 *
 * It is written for clarity, not speed: one recursive tree walk per frame with eager whole-state copies
 * on every list permutation -- exactly the observable semantics of the reference, whose files are cited
 * as PD/... = PolarDecoder/PolarDecoder/_cpp/... relative to the reference root.
 *
 * All per-path values (float LLRs *and* LUT symbols) live in one double array val[(n+1)*N]; a LUT symbol
 * is a small non-negative integer held exactly in a double.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum {
    PO_SC = 0, PO_FASTSC = 1, PO_SCL = 2, PO_FASTSCL = 3, PO_CASCL = 4,
    PO_SCLUT = 5, PO_FASTSCLUT = 6, PO_SCLLUT = 7, PO_FASTSCLLUT = 8, PO_CASCLLUT = 9, PO_CAFASTSCLLUT = 10,
    PO_SC_UNIFORM = 11, PO_SCL_UNIFORM = 12, PO_SC_LLOYD = 13, PO_SCL_LLOYD = 14,
    /* blind-detection helpers, BD/... = PolarEncoder/PolarBD/_cpp/... */
    PO_BD_DMETRIC = 15,   /* DMetricCalculator::calculate   BD/src/DMetric.cpp:25-176 */
    PO_BD_CASCL = 16      /* CASCL::decode(llr, RNTI)       BD/src/CASCLWithRNTI.cpp:74-252 */
};

typedef struct {
    int32_t kind, N, K, A, L;
    const int32_t *frozen;     /* [N] 1 = frozen */
    const int32_t *node_type;  /* [2N-1] (-1 ordinary, 0 R0, 1 R1, 2 REP, 3 SPC) or NULL */
    int32_t crc_n;
    const int32_t *crc_p;      /* [crc_n+1] generator bits in PD/src/CASCLDecoder.cpp:49-53 order; float CASCL only */
    /* LUT family: tables flattened into one int32 pool.  f table of node p: pool[f_off[p] + (pos*fqa+a)*fqb+b]
     * with pos forced to 0 when f_pos[p]==1; g table: pool[g_off[p] + ((pos*2+u)*gqa+a)*gqb+b].          */
    const int32_t *lut_pool;
    const int64_t *f_off, *g_off;      /* [N-1] */
    const int32_t *f_pos, *g_pos;      /* [N-1] number of per-position tables stored (1 or N>>(depth+1)) */
    const int32_t *fqa, *fqb, *gqa, *gqb;
    const double *llr_pool;
    const int64_t *llr_off;            /* [llr_levels*N] start of row (level,pos) */
    int32_t llr_levels;
    /* uniform family (PD/src/SCUniformQuantizedDecoder.cpp:55-57,71-73) */
    const double *r_f, *r_g;           /* [N-1] */
    int32_t v;
    /* Lloyd family (PD/src/SCLloydQuantizedDecoder.cpp:57-59,73-75) */
    const double *bnd_f, *bnd_g, *rec_f, *rec_g; /* [N-1][nb] / [N-1][nr] */
    int32_t nb, nr;
} po_config;

/* ------------------------------------------------------------------------------------------------
 * libstdc++ std::sort(idx.begin(), idx.end(), [&](int a,int b){return key[a]<key[b];})   (GCC 13.3,
 * bits/stl_algo.h:1848-1950, bits/stl_heap.h) restated: introsort (median-of-3, unguarded Hoare partition,
 * heapsort fallback at depth 2*floor(lg n)) followed by the final insertion sort with threshold 16.
 * Call sites in the reference: mink (PD/src/SCLLUTDecoder.cpp:16 and the copies in every list decoder),
 * argsort (PD/src/FastSCLLUTDecoder.cpp:14, CASCLDecoder.cpp:37, CASCLLUTDecoder.cpp:38).
 * The ORDER OF EQUAL KEYS is decided here and is observable in the decoded bits.
 * ------------------------------------------------------------------------------------------------ */
typedef struct { const double *key; } po_cmp;
#define LESS(c, a, b) ((c)->key[(a)] < (c)->key[(b)])

static void ss_unguarded_linear_insert(int *last, const po_cmp *c) {
    int val = *last;
    int *next = last - 1;
    while (LESS(c, val, *next)) { *last = *next; last = next; --next; }
    *last = val;
}
static void ss_insertion_sort(int *first, int *last, const po_cmp *c) {
    if (first == last) return;
    for (int *i = first + 1; i != last; ++i) {
        if (LESS(c, *i, *first)) {
            int val = *i;
            memmove(first + 1, first, (size_t)(i - first) * sizeof(int));
            *first = val;
        } else {
            ss_unguarded_linear_insert(i, c);
        }
    }
}
static void ss_push_heap(int *first, long hole, long top, int value, const po_cmp *c) {
    long parent = (hole - 1) / 2;
    while (hole > top && LESS(c, first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
static void ss_adjust_heap(int *first, long hole, long len, int value, const po_cmp *c) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (LESS(c, first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    ss_push_heap(first, hole, top, value, c);
}
static void ss_heapsort(int *first, int *last, const po_cmp *c) {
    long len = last - first;
    if (len >= 2) { /* make_heap */
        long parent = (len - 2) / 2;
        for (;;) {
            int value = first[parent];
            ss_adjust_heap(first, parent, len, value, c);
            if (parent == 0) break;
            parent--;
        }
    }
    while (last - first > 1) { /* sort_heap via pop_heap */
        --last;
        int value = *last;
        *last = *first;
        ss_adjust_heap(first, 0, last - first, value, c);
    }
}
static void ss_swap(int *a, int *b) { int t = *a; *a = *b; *b = t; }
static void ss_move_median_to_first(int *r, int *a, int *b, int *cc, const po_cmp *c) {
    if (LESS(c, *a, *b)) {
        if (LESS(c, *b, *cc)) ss_swap(r, b);
        else if (LESS(c, *a, *cc)) ss_swap(r, cc);
        else ss_swap(r, a);
    } else if (LESS(c, *a, *cc)) ss_swap(r, a);
    else if (LESS(c, *b, *cc)) ss_swap(r, cc);
    else ss_swap(r, b);
}
static int *ss_unguarded_partition(int *first, int *last, int *pivot, const po_cmp *c) {
    for (;;) {
        while (LESS(c, *first, *pivot)) ++first;
        --last;
        while (LESS(c, *pivot, *last)) --last;
        if (!(first < last)) return first;
        ss_swap(first, last);
        ++first;
    }
}
static void ss_introsort_loop(int *first, int *last, long depth_limit, const po_cmp *c) {
    while (last - first > 16) {
        if (depth_limit == 0) { ss_heapsort(first, last, c); return; }
        --depth_limit;
        int *mid = first + (last - first) / 2;
        ss_move_median_to_first(first, first + 1, mid, last - 1, c);
        int *cut = ss_unguarded_partition(first + 1, last, first, c);
        ss_introsort_loop(cut, last, depth_limit, c);
        last = cut;
    }
}
/* idx must hold 0..n-1 on entry */
void po_std_sort_idx(int *idx, int n, const double *key) {
    po_cmp c = { key };
    if (n <= 0) return;
    long lg = 0;
    for (long t = n; t > 1; t >>= 1) lg++;
    ss_introsort_loop(idx, idx + n, 2 * lg, &c);
    if (n > 16) {
        ss_insertion_sort(idx, idx + 16, &c);
        for (int *i = idx + 16; i != idx + n; ++i) ss_unguarded_linear_insert(i, &c);
    } else {
        ss_insertion_sort(idx, idx + n, &c);
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* scalar primitives, PD/src/utils.cpp and PD/include/utils.h:13                                    */
#define SGN(x) (((x) < 0) ? -1 : ((x) > 0))
static inline double p_minsum(double a, double b) {            /* f, utils.cpp:26-30 */
    double fa = fabs(a), fb = fabs(b);
    return (SGN(a)) * (SGN(b)) * (fa < fb ? fa : fb);          /* std::min(x,y) = (y<x)?y:x ; same value */
}
static inline double p_g(double a, double b, int u) {          /* g, utils.cpp:32-36 */
    return (1 - 2 * u) * a + b;
}
static inline double p_Q(double x, double r, double M) {       /* Q, utils.cpp:8-10 */
    return fabs(x) > M ? SGN(x) * (M - 0.5 * r) : (floor(x / r) + 0.5) * r;
}
static inline double p_bisect(double a, const double *boundary, int nb, const double *reconstruct) {
    int lo = 0, hi = nb;                                        /* utils.cpp:12-24 */
    while (lo < hi) {
        int mid = (lo + hi) / 2;
        if (boundary[mid] < a) lo = mid + 1; else hi = mid;
    }
    return reconstruct[lo - 1];
}
/* CRC long division, utils.cpp:77-93 / CASCLDecoder.cpp:56-72.  poly[0..crc_n], out[crc_n] */
static void p_crc(const uint8_t *info, int len, const int32_t *poly, int crc_n, uint8_t *out) {
    uint8_t *u = (uint8_t *)calloc((size_t)(len + crc_n), 1);
    memcpy(u, info, (size_t)len);
    for (int i = 0; i < len; ++i)
        if (u[i] == 1)
            for (int j = 0; j < crc_n + 1; ++j) u[j + i] = (uint8_t)((u[j + i] + poly[j]) % 2);
    memcpy(out, u + len, (size_t)crc_n);
    free(u);
}

/* Encoder side (the PolarBDEnc package the drivers import is not in the reference tree; its conventions are fixed by
 * the code that IS there): CRC = msg || CRC::encoding(msg) (utils.cpp:77-93); polar encode = u[msgbits] = word,
 * x = u F^{(x)n} in natural order with the butterfly of the decoders' own re-encode (FastSCDecoder.cpp:153-164).
 * in [B][len] bits, out [B][len+crc_n]. */
void po_crc_attach(const uint8_t *in, int64_t B, int len, const int32_t *poly, int crc_n, uint8_t *out) {
    for (int64_t f = 0; f < B; ++f) {
        memcpy(out + f * (len + crc_n), in + f * len, (size_t)len);
        p_crc(in + f * len, len, poly, crc_n, out + f * (len + crc_n) + len);
    }
}
/* in [B][K] bits placed at info_pos[0..K), out [B][N] */
void po_polar_encode(const uint8_t *in, int64_t B, int K, const int32_t *info_pos, int N, uint8_t *out) {
    for (int64_t f = 0; f < B; ++f) {
        uint8_t *x = out + f * N;
        memset(x, 0, (size_t)N);
        for (int k = 0; k < K; ++k) x[info_pos[k]] = in[f * K + k];
        for (int m = 1; m < N; m *= 2)
            for (int i = 0; i < N; i += 2 * m)
                for (int j = 0; j < m; ++j) x[i + j] = x[i + j] ^ x[i + m + j];
    }
}

/* ------------------------------------------------------------------------------------------------ */
typedef struct {
    const po_config *c;
    int n, L, lut, list, fast, spc, ca, quant; /* quant: 0 float, 1 uniform, 2 lloyd */
    double pm_init;
    double **val;    /* [L][(n+1)*N] */
    uint8_t **uc;    /* [L][(n+1)*N] */
    double *PM;      /* [L] */
    double dmetric;  /* PO_BD_DMETRIC accumulator */
    /* scratch for permutations */
    double **tval; uint8_t **tuc;
} po_ctx;

static inline double llr_at(const po_ctx *x, int level, int pos, int sym) {
    const po_config *c = x->c;
    return c->llr_pool[c->llr_off[(int64_t)level * c->N + pos] + sym];
}
/* LLR seen by a leaf / special node element at (depth d, absolute position pos) of path i:
 * float domains: val[d*N+pos]; LUT: virtual_channel_llrs[d-1][pos][symbol] (PD/src/SCLLUTDecoder.cpp:97,
 * FastSCLLUTDecoder.cpp:89,112,179; SURVEY App. A.5 -- indexed exactly as the reference does). */
static inline double elem_llr(const po_ctx *x, int i, int d, int pos) {
    double v = x->val[i][(size_t)d * x->c->N + pos];
    if (!x->lut) return v;
    return llr_at(x, d - 1, pos, (int)v);
}

static void permute_paths(po_ctx *x, const int *parent) {
    /* new slot i <- old slot parent[i], whole state (PD/src/SCLLUTDecoder.cpp:132-144) */
    size_t sz = (size_t)(x->n + 1) * x->c->N;
    for (int i = 0; i < x->L; ++i) {
        memcpy(x->tval[i], x->val[parent[i]], sz * sizeof(double));
        memcpy(x->tuc[i], x->uc[parent[i]], sz);
    }
    double **tv = x->val; x->val = x->tval; x->tval = tv;
    uint8_t **tu = x->uc; x->uc = x->tuc; x->tuc = tu;
}

/* mink: indices of the L smallest of PM2[2L] in std::sort order; updates PM.  Returns parents + flip flags. */
static void fork_select(po_ctx *x, const double *PM2, int *parent, int *flip) {
    int L = x->L;
    int *idx = (int *)malloc(sizeof(int) * 2 * (size_t)L);
    for (int i = 0; i < 2 * L; ++i) idx[i] = i;
    po_std_sort_idx(idx, 2 * L, PM2);
    for (int i = 0; i < L; ++i) {
        x->PM[i] = PM2[idx[i]];
        flip[i] = idx[i] >= L;
        parent[i] = idx[i] >= L ? idx[i] - L : idx[i];
    }
    free(idx);
}

static void leaf_action(po_ctx *x, int leaf) {
    const po_config *c = x->c;
    int n = x->n, N = c->N, L = x->L;
    size_t at = (size_t)n * N + leaf;
    if (!x->list) {
        /* PD/src/SCDecoder.cpp:27-32, SCLUTDecoder.cpp:59-67 : bit = (llr <= 0) */
        if (c->frozen[leaf] == 1) x->uc[0][at] = 0;
        else x->uc[0][at] = (uint8_t)(elem_llr(x, 0, n, leaf) <= 0);
        return;
    }
    double *DM = (double *)malloc(sizeof(double) * (size_t)L);
    for (int i = 0; i < L; ++i) DM[i] = elem_llr(x, i, n, leaf);
    if (c->frozen[leaf] == 1) {
        /* PD/src/SCLLUTDecoder.cpp:99-104 */
        for (int i = 0; i < L; ++i) {
            x->uc[i][at] = 0;
            x->PM[i] += fabs(DM[i]) * (double)(DM[i] < 0);
        }
    } else {
        /* PD/src/SCLLUTDecoder.cpp:105-145 */
        double *PM2 = (double *)malloc(sizeof(double) * 2 * (size_t)L);
        int *parent = (int *)malloc(sizeof(int) * (size_t)L), *flip = (int *)malloc(sizeof(int) * (size_t)L);
        uint8_t *dec = (uint8_t *)malloc((size_t)L);
        for (int i = 0; i < L; ++i) {
            dec[i] = (uint8_t)(DM[i] < 0);
            PM2[i] = x->PM[i];
            PM2[i + L] = x->PM[i] + fabs(DM[i]);
        }
        fork_select(x, PM2, parent, flip);
        permute_paths(x, parent);
        for (int i = 0; i < L; ++i) x->uc[i][at] = (uint8_t)(flip[i] ? 1 - dec[parent[i]] : dec[parent[i]]);
        free(PM2); free(parent); free(flip); free(dec);
    }
    free(DM);
}

/* special nodes of the non-list Fast decoders: PD/src/FastSCDecoder.cpp:45-106, FastSCLUT.cpp:43-106 */
static void fastsc_special(po_ctx *x, int type, int d, int node) {
    int N = x->c->N, temp = 1 << (x->n - d), base = temp * node;
    uint8_t *pu = x->uc[0] + (size_t)d * N + base;
    if (type == 0) {
        memset(pu, 0, (size_t)temp);
        if (x->c->kind == PO_BD_DMETRIC) {   /* BD/src/DMetric.cpp:56-61 */
            double tmp = 0;
            for (int i = 0; i < temp; ++i) tmp += elem_llr(x, 0, d, base + i);
            x->dmetric += tmp / temp;
        }
        return;
    }
    if (type == 1) { for (int i = 0; i < temp; ++i) pu[i] = (uint8_t)(elem_llr(x, 0, d, base + i) <= 0); return; }
    if (type == 2) {
        double S = 0;
        for (int i = 0; i < temp; ++i) S += elem_llr(x, 0, d, base + i);
        memset(pu, (uint8_t)(S <= 0), (size_t)temp);
        if (x->c->kind == PO_BD_DMETRIC) x->dmetric += fabs(S) / temp;   /* BD/src/DMetric.cpp:87-93 */
        return;
    }
    /* SPC: Wagner, flip the FIRST arg-min |llr| when parity is odd */
    int parity = 0, amin = 0; double best = 0;
    for (int i = 0; i < temp; ++i) {
        double l = elem_llr(x, 0, d, base + i);
        pu[i] = (uint8_t)(l <= 0);
        parity += pu[i];
        if (i == 0 || fabs(l) < best) { best = fabs(l); amin = i; }
    }
    if (parity % 2) pu[amin] = (uint8_t)(1 - pu[amin]);
}

/* special nodes of the list Fast decoders: PD/src/FastSCLDecoder.cpp:121-251, FastSCLLUTDecoder.cpp:82-213 */
static void fastscl_special(po_ctx *x, int type, int d, int node) {
    int N = x->c->N, L = x->L, temp = 1 << (x->n - d), base = temp * node;
    size_t at = (size_t)d * N + base;
    if (type == 0) {
        for (int i = 0; i < L; ++i) {
            memset(x->uc[i] + at, 0, (size_t)temp);
            for (int j = 0; j < temp; ++j) {
                double l = elem_llr(x, i, d, base + j);
                x->PM[i] += (float)(l < 0) * fabs(l);
            }
        }
        return;
    }
    double *PM2 = (double *)malloc(sizeof(double) * 2 * (size_t)L);
    int *parent = (int *)malloc(sizeof(int) * (size_t)L), *flip = (int *)malloc(sizeof(int) * (size_t)L);
    if (type == 2) {
        for (int i = 0; i < L; ++i) { PM2[i] = x->PM[i]; PM2[i + L] = x->PM[i]; }
        for (int i = 0; i < L; ++i)
            for (int j = 0; j < temp; ++j) {
                double l = elem_llr(x, i, d, base + j);
                PM2[i] += (double)(l < 0) * fabs(l);
                PM2[i + L] += (double)(l >= 0) * fabs(l);
            }
        fork_select(x, PM2, parent, flip);
        permute_paths(x, parent);
        for (int i = 0; i < L; ++i) memset(x->uc[i] + at, flip[i] ? 1 : 0, (size_t)temp);
    } else { /* type 1: R1, including the reference's flip-index quirk */
        int rounds = (L - 1 < temp) ? L - 1 : temp;
        uint8_t *dec = (uint8_t *)malloc((size_t)L * temp), *tdec = (uint8_t *)malloc((size_t)L * temp);
        double *al = (double *)malloc(sizeof(double) * (size_t)L * temp), *tal = (double *)malloc(sizeof(double) * (size_t)L * temp);
        int *si = (int *)malloc(sizeof(int) * (size_t)L * temp), *tsi = (int *)malloc(sizeof(int) * (size_t)L * temp);
        for (int i = 0; i < L; ++i) {
            for (int j = 0; j < temp; ++j) {
                double l = elem_llr(x, i, d, base + j);
                dec[i * temp + j] = (uint8_t)(l < 0);
                al[i * temp + j] = fabs(l);
                si[i * temp + j] = j;
            }
            po_std_sort_idx(si + i * temp, temp, al + i * temp);
        }
        for (int layer = 0; layer < rounds; ++layer) {
            for (int i = 0; i < L; ++i) {
                PM2[i] = x->PM[i];
                PM2[i + L] = x->PM[i] + al[i * temp + si[i * temp + layer]];
            }
            fork_select(x, PM2, parent, flip);
            for (int i = 0; i < L; ++i) {
                memcpy(tdec + i * temp, dec + parent[i] * temp, (size_t)temp);
                if (flip[i]) {
                    int q = si[i * temp + layer];          /* slot i's OWN (pre-permutation) ordering */
                    tdec[i * temp + q] = (uint8_t)(1 - tdec[i * temp + q]);
                }
                memcpy(tal + i * temp, al + parent[i] * temp, sizeof(double) * (size_t)temp);
                memcpy(tsi + i * temp, si + parent[i] * temp, sizeof(int) * (size_t)temp);
            }
            permute_paths(x, parent);
            { uint8_t *t = dec; dec = tdec; tdec = t; }
            { double *t = al; al = tal; tal = t; }
            { int *t = si; si = tsi; tsi = t; }
        }
        for (int i = 0; i < L; ++i) memcpy(x->uc[i] + at, dec + i * temp, (size_t)temp);
        free(dec); free(tdec); free(al); free(tal); free(si); free(tsi);
    }
    free(PM2); free(parent); free(flip);
}

static void step_f(po_ctx *x, int d, int node) {
    const po_config *c = x->c;
    int N = c->N, temp = 1 << (x->n - d), ct = temp / 2, p = (1 << d) + node - 1;
    for (int i = 0; i < x->L; ++i) {
        const double *pa = x->val[i] + (size_t)d * N + (size_t)temp * node, *pb = pa + ct;
        double *out = x->val[i] + (size_t)(d + 1) * N + (size_t)ct * (2 * node);
        for (int j = 0; j < ct; ++j) {
            if (x->lut) {
                int pos = c->f_pos[p] == 1 ? 0 : j;
                out[j] = c->lut_pool[c->f_off[p] + ((int64_t)pos * c->fqa[p] + (int)pa[j]) * c->fqb[p] + (int)pb[j]];
            } else {
                double r = p_minsum(pa[j], pb[j]);
                if (x->quant == 1) { double rf = c->r_f[p]; r = p_Q(r, rf, (double)(c->v / 2 - 0.5) * rf); }
                else if (x->quant == 2) r = p_bisect(r, c->bnd_f + (size_t)p * c->nb, c->nb, c->rec_f + (size_t)p * c->nr);
                out[j] = r;
            }
        }
    }
}
static void step_g(po_ctx *x, int d, int node) {
    const po_config *c = x->c;
    int N = c->N, temp = 1 << (x->n - d), ct = temp / 2, p = (1 << d) + node - 1;
    for (int i = 0; i < x->L; ++i) {
        const double *pa = x->val[i] + (size_t)d * N + (size_t)temp * node, *pb = pa + ct;
        const uint8_t *ul = x->uc[i] + (size_t)(d + 1) * N + (size_t)ct * (2 * node);
        double *out = x->val[i] + (size_t)(d + 1) * N + (size_t)ct * (2 * node + 1);
        for (int j = 0; j < ct; ++j) {
            if (x->lut) {
                int pos = c->g_pos[p] == 1 ? 0 : j;
                out[j] = c->lut_pool[c->g_off[p] + (((int64_t)pos * 2 + ul[j]) * c->gqa[p] + (int)pa[j]) * c->gqb[p] + (int)pb[j]];
            } else {
                double r = p_g(pa[j], pb[j], ul[j]);
                if (x->quant == 1) { double rg = c->r_g[p]; r = p_Q(r, rg, (double)(c->v / 2 - 1) * rg); }
                else if (x->quant == 2) r = p_bisect(r, c->bnd_g + (size_t)p * c->nb, c->nb, c->rec_g + (size_t)p * c->nr);
                out[j] = r;
            }
        }
    }
}
static void step_combine(po_ctx *x, int d, int node) { /* u(), utils.cpp:62-67 */
    int N = x->c->N, temp = 1 << (x->n - d), ct = temp / 2;
    for (int i = 0; i < x->L; ++i) {
        const uint8_t *ul = x->uc[i] + (size_t)(d + 1) * N + (size_t)ct * (2 * node), *ur = ul + ct;
        uint8_t *o = x->uc[i] + (size_t)d * N + (size_t)temp * node;
        for (int j = 0; j < ct; ++j) { o[j] = ul[j] ^ ur[j]; o[j + ct] = ur[j]; }
    }
}

static void visit(po_ctx *x, int d, int node) {
    if (d == x->n) { leaf_action(x, node); return; }
    if (x->fast) {
        int t = x->c->node_type[(1 << d) + node - 1];
        if (t >= 0 && t <= (x->spc ? 3 : 2)) {
            if (x->list) fastscl_special(x, t, d, node); else fastsc_special(x, t, d, node);
            return;
        }
    }
    step_f(x, d, node);
    visit(x, d + 1, 2 * node);
    step_g(x, d, node);
    visit(x, d + 1, 2 * node + 1);
    step_combine(x, d, node);
}

static const int32_t CRC24_LOC[13] = {24, 23, 21, 20, 17, 15, 13, 12, 8, 4, 2, 1, 0}; /* PD/include/CASCLLUTDecoder.h:33 */

/* in: int32 [B][N] for LUT kinds, double [B][N] otherwise.  out: uint8 [B][Kout], Kout = A for CA kinds else K.
 * pm_out (optional): double [B][L] final path metrics in slot order; winner_out (optional): int32 [B].   */
static int decode_impl(const po_config *c, const void *in, int64_t B, uint8_t *out, double *pm_out, int32_t *winner_out,
                       const int32_t *rnti, int rnti_len, double *bd_metric, uint8_t *bd_pass) {
    po_ctx x; memset(&x, 0, sizeof x);
    x.c = c;
    int N = c->N, n = (int)log2f((float)N), k = c->kind;
    x.n = n;
    x.lut = (k >= PO_SCLUT && k <= PO_CAFASTSCLLUT);
    x.list = (k == PO_SCL || k == PO_FASTSCL || k == PO_CASCL || k == PO_SCLLUT || k == PO_FASTSCLLUT ||
              k == PO_CASCLLUT || k == PO_CAFASTSCLLUT || k == PO_SCL_UNIFORM || k == PO_SCL_LLOYD || k == PO_BD_CASCL);
    x.fast = (k == PO_FASTSC || k == PO_FASTSCL || k == PO_FASTSCLUT || k == PO_FASTSCLLUT || k == PO_CAFASTSCLLUT || k == PO_BD_DMETRIC);
    x.spc = (k == PO_FASTSC || k == PO_FASTSCLUT || k == PO_BD_DMETRIC);  /* list variants expand SPC nodes */
    x.ca = (k == PO_CASCL || k == PO_CASCLLUT || k == PO_CAFASTSCLLUT || k == PO_BD_CASCL);
    x.quant = (k == PO_SC_UNIFORM || k == PO_SCL_UNIFORM) ? 1 : (k == PO_SC_LLOYD || k == PO_SCL_LLOYD) ? 2 : 0;
    /* PD/include/SCLDecoder.h:8 (1e300) vs SCLLUTDecoder.h:16, FastSCLDecoder.h:7, FastSCLLUTDecoder.h:16 (+inf) */
    x.pm_init = (k == PO_SCL || k == PO_CASCL || k == PO_SCL_UNIFORM || k == PO_SCL_LLOYD) ? 1e300 : INFINITY;
    if (k == PO_BD_CASCL) x.pm_init = 1e30;   /* BD/src/CASCLWithRNTI.cpp:82 */
    x.L = x.list ? c->L : 1;
    int L = x.L, Kout = x.ca ? c->A : c->K;
    size_t sz = (size_t)(n + 1) * N;
    x.val = (double **)malloc(sizeof(double *) * L); x.tval = (double **)malloc(sizeof(double *) * L);
    x.uc = (uint8_t **)malloc(sizeof(uint8_t *) * L); x.tuc = (uint8_t **)malloc(sizeof(uint8_t *) * L);
    x.PM = (double *)malloc(sizeof(double) * L);
    for (int i = 0; i < L; ++i) {
        x.val[i] = (double *)malloc(sz * sizeof(double)); x.tval[i] = (double *)malloc(sz * sizeof(double));
        x.uc[i] = (uint8_t *)malloc(sz); x.tuc[i] = (uint8_t *)malloc(sz);
    }
    int32_t crc24[25]; memset(crc24, 0, sizeof crc24);
    for (int i = 0; i < 13; ++i) crc24[CRC24_LOC[i]] = 1;
    uint8_t *xw = (uint8_t *)malloc((size_t)N), *info = (uint8_t *)malloc((size_t)N), *chk = (uint8_t *)malloc(64 + (size_t)(c->crc_n > 0 ? c->crc_n : 0));
    int *order = (int *)malloc(sizeof(int) * L);

    for (int64_t b = 0; b < B; ++b) {
        for (int i = 0; i < L; ++i) {
            memset(x.val[i], 0, sz * sizeof(double)); memset(x.uc[i], 0, sz);
            for (int j = 0; j < N; ++j)
                x.val[i][j] = x.lut ? (double)((const int32_t *)in)[b * N + j] : ((const double *)in)[b * N + j];
            x.PM[i] = (i == 0) ? 0.0 : x.pm_init;
        }
        x.dmetric = 0;
        visit(&x, 0, 0);
        int winner = 0, passed = 0;
        double ret_pm = x.PM[0];   /* BD/src/CASCLWithRNTI.cpp:205 */
        if (x.list) {
            if (!x.ca) {
                for (int i = 1; i < L; ++i) if (x.PM[i] < x.PM[winner]) winner = i;   /* std::min_element: first minimum */
            } else {
                for (int i = 0; i < L; ++i) order[i] = i;
                po_std_sort_idx(order, L, x.PM);
                winner = order[0];
                const int own = (k == PO_CASCL || k == PO_BD_CASCL);
                const int32_t *poly = own ? c->crc_p : crc24;
                int crc_n = own ? c->crc_n : 24;
                int ncheck = own ? c->crc_n : c->K - c->A;    /* CASCLDecoder.cpp:222 vs CASCLLUTDecoder.cpp:280 */
                for (int t = 0; t < L; ++t) {
                    int cand = order[t], cnt = 0;
                    if (x.fast) {   /* CAFastSCLLUTDecoder.cpp:396-415: re-encode each candidate first */
                        memcpy(xw, x.uc[cand], (size_t)N);
                        for (int m = 1; m < N; m *= 2)
                            for (int i = 0; i < N; i += 2 * m)
                                for (int j = 0; j < m; ++j) xw[i + j] ^= xw[i + m + j];
                        for (int i = 0; i < N; ++i) if (c->frozen[i] == 0) info[cnt++] = xw[i];
                    } else {
                        for (int i = 0; i < N; ++i) if (c->frozen[i] == 0) info[cnt++] = x.uc[cand][(size_t)n * N + i];
                    }
                    p_crc(info, c->A, poly, crc_n, chk);
                    for (int l = 0; l < rnti_len; ++l)   /* BD/src/CASCLWithRNTI.cpp:222-224 */
                        chk[l + crc_n - rnti_len] = (uint8_t)((chk[l + crc_n - rnti_len] + rnti[l]) % 2);
                    int pass = 1;
                    for (int j = 0; j < ncheck; ++j) if (chk[j] != info[c->A + j]) { pass = 0; break; }
                    if (pass) { winner = cand; passed = 1; ret_pm = x.PM[t]; break; }   /* :236 PM = PML[i], the slot, not the candidate */
                }
            }
        }
        if (bd_metric) bd_metric[b] = (k == PO_BD_DMETRIC) ? x.dmetric : ret_pm;
        if (bd_pass) bd_pass[b] = (uint8_t)passed;
        if (!out) continue;
        uint8_t *o = out + b * Kout;
        int cnt = 0;
        if (x.fast) { /* re-encode the root codeword estimate: PD/src/FastSCLUT.cpp:186-205 */
            memcpy(xw, x.uc[winner], (size_t)N);
            for (int m = 1; m < N; m *= 2)
                for (int i = 0; i < N; i += 2 * m)
                    for (int j = 0; j < m; ++j) xw[i + j] ^= xw[i + m + j];
            for (int i = 0; i < N && cnt < Kout; ++i) if (c->frozen[i] == 0) o[cnt++] = xw[i];
        } else {
            for (int i = 0; i < N && cnt < Kout; ++i) if (c->frozen[i] == 0) o[cnt++] = x.uc[winner][(size_t)n * N + i];
        }
        if (pm_out) for (int i = 0; i < L; ++i) pm_out[b * L + i] = x.PM[i];
        if (winner_out) winner_out[b] = winner;
    }
    for (int i = 0; i < L; ++i) { free(x.val[i]); free(x.tval[i]); free(x.uc[i]); free(x.tuc[i]); }
    free(x.val); free(x.tval); free(x.uc); free(x.tuc); free(x.PM);
    free(xw); free(info); free(chk); free(order);
    return 0;
}

int po_decode(const po_config *c, const void *in, int64_t B, uint8_t *out, double *pm_out, int32_t *winner_out) {
    return decode_impl(c, in, B, out, pm_out, winner_out, NULL, 0, NULL, NULL);
}
/* PO_BD_DMETRIC: metric[b] = the D-metric, out may be NULL.  PO_BD_CASCL: out [B][A], metric[b] = returned PM, pass[b]. */
int po_decode_bd(const po_config *c, const double *in, int64_t B, const int32_t *rnti, int32_t rnti_len,
                 uint8_t *out, double *metric, uint8_t *pass) {
    return decode_impl(c, in, B, out, NULL, NULL, rnti, rnti_len, metric, pass);
}

/* ------------------------------------------------------------------------------------------------
 * Lookup-table design: the minimum-distortion quantizer the LLR-domain generator calls on every
 * tree node, LLRQuantizer.find_OptLS_quantizer (QuantizeDensityEvolution/QLLRDensityEvolution_MinDistortion.py:107-108).
 * The generator of record is C++ on OpenCV (Quantizers/quantizers/_cpp/LLRQuantizer/LLRQuantizer.cpp, cannot be built
 * here); the reference also ships a numpy restatement, QuantizeDensityEvolution/MinDistortionQuantizer.py:28-99 = MDQ,
 * which is what this follows INCLUDING numpy's summation order (np.sum of a contiguous float64 array = the pairwise
 * routine below; checked against numpy 2.3 in tests/test_lutgen.py and pinned by golden vectors made with MDQ itself).
 * ------------------------------------------------------------------------------------------------ */
static double np_pairwise(const double *a, int n) {
    if (n < 8) {
        double r = 0.;
        for (int i = 0; i < n; ++i) r += a[i];
        return r;
    }
    if (n <= 128) {
        double r[8];
        int i;
        for (i = 0; i < 8; ++i) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise(a, n2) + np_pairwise(a + n2, n - n2);
}
double po_np_sum(const double *a, int n) { return np_pairwise(a, n); }

/* MDQ:3-7 compute_partial_quantization_noise on the slice [lo,hi) */
static double optls_noise(const double *d, const double *q, int lo, int hi, double *tmp) {
    int n = hi - lo;
    for (int i = 0; i < n; ++i) tmp[i] = d[lo + i] * q[lo + i];
    double nq = np_pairwise(tmp, n) / np_pairwise(d + lo, n);
    for (int i = 0; i < n; ++i) { double e = q[lo + i] - nq; tmp[i] = (e * e) * d[lo + i]; }
    return np_pairwise(tmp, n);
}
/* density / quanta: M entries sorted by ascending quanta (the generator passes np.unique output, so MDQ's argsort is
 * the identity), M > K.  out_density[K], out_quanta[K], out_lut[M]. */
int po_optls_quantize(const double *d, const double *q, int M, int K, double *out_density, double *out_quanta, int32_t *out_lut) {
    if (M <= K || K < 2) return 1;
    const int W = M - K + 1;
    double *T = (double *)calloc((size_t)M * (M + 1), sizeof(double)), *tmp = (double *)malloc(sizeof(double) * M);
    double *state = (double *)calloc((size_t)W * (K + 1), sizeof(double));
    int *lm = (int *)calloc((size_t)W * (K + 1), sizeof(int)), *Az = (int *)calloc((size_t)K + 1, sizeof(int));
    for (int ap = 0; ap < M; ++ap) {                       /* MDQ:20-24 */
        int max_a = ap + W < M ? ap + W : M;
        for (int a = ap + 1; a <= max_a; ++a) T[(size_t)ap * (M + 1) + a] = optls_noise(d, q, ap, a, tmp);
    }
    for (int i = 0; i < W; ++i) state[(size_t)i * (K + 1) + 1] = T[1 + i];   /* MDQ:44 */
    for (int z = 2; z <= K; ++z) {                        /* MDQ:50-77: np.argmin / np.min = first minimum */
        int a_lo = z < K ? z : M, a_hi = z < K ? z + M - K : M;
        for (int a = a_lo; a <= a_hi; ++a) {
            int row = z < K ? a - z : W - 1, best_ap = z - 1;
            double best = state[(size_t)0 * (K + 1) + (z - 1)] + T[(size_t)(z - 1) * (M + 1) + a];
            for (int ap = z; ap <= a - 1; ++ap) {
                double v = state[(size_t)(ap - (z - 1)) * (K + 1) + (z - 1)] + T[(size_t)ap * (M + 1) + a];
                if (v < best) { best = v; best_ap = ap; }
            }
            state[(size_t)row * (K + 1) + z] = best;
            lm[(size_t)row * (K + 1) + z] = best_ap;
        }
    }
    Az[K] = M;                                             /* MDQ:80-84 backward tracing */
    int opt = lm[(size_t)(W - 1) * (K + 1) + K];
    Az[K - 1] = opt;
    for (int z = K - 1; z >= 2; --z) { opt = lm[(size_t)(opt - z) * (K + 1) + z]; Az[z - 1] = opt; }
    for (int i = 0; i < K; ++i) {                          /* MDQ:87-96 */
        int b = Az[i], e = Az[i + 1];
        for (int j = b; j < e; ++j) { out_lut[j] = i; tmp[j - b] = q[j] * d[j]; }
        double sd = np_pairwise(d + b, e - b);
        out_quanta[i] = np_pairwise(tmp, e - b) / sd;
        out_density[i] = sd;
    }
    free(T); free(tmp); free(state); free(lm); free(Az);
    return 0;
}

