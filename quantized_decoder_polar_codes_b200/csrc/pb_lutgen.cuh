// Lookup-table design on the GPU ("next" row f3): the minimum-distortion quantizer that the reference's LLR-domain
// generator runs on the f and g output distribution of every tree node --
// LLRQuantizer.find_OptLS_quantizer, QuantizeDensityEvolution/QLLRDensityEvolution_MinDistortion.py:107-108.
// The generator of record is C++ on OpenCV (Quantizers/quantizers/_cpp/LLRQuantizer/LLRQuantizer.cpp) and cannot be
// built here; this follows the numpy restatement the reference ships, QuantizeDensityEvolution/MinDistortionQuantizer.py
// (= MDQ below), bit for bit -- including numpy's summation order (np.sum of a contiguous float64 array is the pairwise
// routine np_pairwise) and np.argmin's first-minimum rule.  In pure Python that code needs minutes for N=128 and hours
// for N=1024; here every node of a tree level is one CTA of one launch.
//
// One CTA per problem: (1) the banded noise table T[a'][a] = distortion of merging sorted symbols [a',a) into one
// (MDQ:10-24), one thread per entry, three pairwise sums each; (2) the K-stage dynamic programme over cut positions
// (MDQ:50-77), threads across the end position, two ping-pong columns in shared memory; (3) back-tracing (MDQ:80-84)
// and the per-cluster centroid / mass (MDQ:87-96).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace pb {

constexpr int kOptlsMaxM = 1024;

// numpy's pairwise summation (n <= 128: eight running sums; above: split at n/2 rounded down to a multiple of 8).  The
// recursion is unrolled at compile time (DEPTH splits cover n <= 1024 with room to spare) so that no device stack is needed.
template <class F>
__device__ __forceinline__ double np_pairwise_block(F f, int lo, int n) {   // n <= 128
    if (n < 8) {
        double r = 0.;
        for (int i = 0; i < n; ++i) r += f(lo + i);
        return r;
    }
    double r[8];
    int i;
#pragma unroll
    for (i = 0; i < 8; ++i) r[i] = f(lo + i);
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += f(lo + i + j);
    }
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += f(lo + i);
    return res;
}
template <int DEPTH, class F>
__device__ double np_pairwise_d(F f, int lo, int n) {
    if (n <= 128) return np_pairwise_block(f, lo, n);
    if constexpr (DEPTH > 0) {
        int n2 = n / 2;
        n2 -= n2 % 8;
        const double a = np_pairwise_d<DEPTH - 1>(f, lo, n2);
        const double b = np_pairwise_d<DEPTH - 1>(f, lo + n2, n - n2);
        return a + b;
    } else {
        return __longlong_as_double(0x7ff8000000000000ll);   // unreachable for n <= kOptlsMaxM
    }
}
template <class F>
__device__ __forceinline__ double np_pairwise(F f, int lo, int n) { return np_pairwise_d<6>(f, lo, n); }

// T is stored banded: row a' holds a = a'+1 .. a'+W at T[a' * W + (a - a' - 1)]
__global__ void __launch_bounds__(256)
optls_kernel(const double *__restrict__ density, const double *__restrict__ quanta, const int32_t *__restrict__ Ms, long long stride,
             int K, double *__restrict__ out_density, double *__restrict__ out_quanta, int32_t *__restrict__ out_lut,
             double *__restrict__ ws_T, int32_t *__restrict__ ws_lm, long long t_stride, long long lm_stride) {
    extern __shared__ double sm_d[];
    const int p = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
    const int M = Ms[p], W = M - K + 1;
    double *d = sm_d, *q = sm_d + M, *col0 = q + M, *col1 = col0 + W;
    __shared__ int Az[64 + 1];
    for (int i = tid; i < M; i += nth) { d[i] = density[(size_t)p * stride + i]; q[i] = quanta[(size_t)p * stride + i]; }
    __syncthreads();
    double *T = ws_T + (size_t)p * t_stride;
    int32_t *lm = ws_lm + (size_t)p * lm_stride;     // [K+1][W]
    auto fd = [&](int i) { return d[i]; };
    auto fdq = [&](int i) { return d[i] * q[i]; };
    // (1) noise table, MDQ:3-7,20-24
    for (long long e = tid; e < (long long)M * W; e += nth) {
        const int ap = (int)(e / W), a = ap + 1 + (int)(e % W);
        if (a > M) continue;
        const int n = a - ap;
        const double nq = np_pairwise(fdq, ap, n) / np_pairwise(fd, ap, n);
        auto fn = [&](int i) { const double t = q[i] - nq; return (t * t) * d[i]; };
        T[e] = np_pairwise(fn, ap, n);
    }
    __syncthreads();
    // (2) dynamic programme, MDQ:44,50-77.  col[i] = state_table[i][z-1]
    for (int i = tid; i < W; i += nth) col0[i] = T[i];          // state[:,1] = T[0][1 .. W]
    __syncthreads();
    double *prev = col0, *cur = col1;
    for (int z = 2; z <= K; ++z) {
        const int a_lo = z < K ? z : M, cnt = z < K ? W : 1;
        for (int r = tid; r < cnt; r += nth) {
            const int a = a_lo + r;
            int best_ap = z - 1;
            double best = prev[0] + T[(size_t)(z - 1) * W + (a - z)];
            for (int ap = z; ap <= a - 1; ++ap) {
                const double v = prev[ap - (z - 1)] + T[(size_t)ap * W + (a - ap - 1)];
                if (v < best) { best = v; best_ap = ap; }
            }
            const int row = z < K ? r : W - 1;
            cur[row] = best;
            lm[(size_t)z * W + row] = best_ap;
        }
        __syncthreads();
        double *t = prev; prev = cur; cur = t;
    }
    // (3) back-tracing MDQ:80-84 and the clusters MDQ:87-96
    if (tid == 0) {
        Az[0] = 0;
        Az[K] = M;
        int opt = lm[(size_t)K * W + (W - 1)];
        Az[K - 1] = opt;
        for (int z = K - 1; z >= 2; --z) { opt = lm[(size_t)z * W + (opt - z)]; Az[z - 1] = opt; }
    }
    __syncthreads();
    for (int i = tid; i < K; i += nth) {
        const int b = Az[i], e = Az[i + 1];
        auto fqd = [&](int j) { return q[j] * d[j]; };
        const double sd = np_pairwise(fd, b, e - b);
        out_quanta[(size_t)p * K + i] = np_pairwise(fqd, b, e - b) / sd;
        out_density[(size_t)p * K + i] = sd;
    }
    for (int j = tid; j < M; j += nth) {
        int c = 0;
        while (j >= Az[c + 1]) ++c;
        out_lut[(size_t)p * stride + j] = c;
    }
}

// ------------------------------------------------------------------------------------------------
// Probability-domain (maximum mutual information) quantizer design, the `MMIQuantizer.find_opt_quantizer` every node of
// QuantizeDensityEvolution/QDensityEvolution_MMI.py:84,107 runs.  Followed bit for bit: the reference's numpy restatement
// QuantizeDensityEvolution/MMIQuantizer.py (= MMQ below; the quantizer of record is C++ on OpenCV and cannot be built here).
// Its cost table l(a',a) (MMQ:36-84) needs log2 of the cluster likelihoods; to stay identical with numpy the logarithms are
// taken by numpy on the host between two passes of mmi_table_kernel:
//   mode 0:  out1 = np.sum(p1[a':a]), out2 = np.sum(p2[a':a])                                         (MMQ:42-43)
//   mode 1:  out1 = c1 * np.sum(p1[a':a] * l1) + c2 * np.sum(p2[a':a] * l2)   with l = log2(p) per entry  (MMQ:55-67)
// both banded like T above (row a' holds a = a'+1 .. a'+W), one thread per entry, sums in numpy's pairwise order.
__global__ void __launch_bounds__(256)
mmi_table_kernel(const double *__restrict__ p1, const double *__restrict__ p2, int M, int W, int mode,
                 const double *__restrict__ l1, const double *__restrict__ l2, double c1, double c2,
                 double *__restrict__ out1, double *__restrict__ out2) {
    const int p = blockIdx.y;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)M * W) return;
    const double *q1 = p1 + (size_t)p * M, *q2 = p2 + (size_t)p * M;
    const size_t o = (size_t)p * M * W + (size_t)e;
    const int ap = (int)(e / W), a = ap + 1 + (int)(e % W);
    if (a > M) { out1[o] = 0.; if (mode == 0) out2[o] = 0.; return; }
    const int n = a - ap;
    if (mode == 0) {
        auto f1 = [&](int i) { return q1[i]; };
        auto f2 = [&](int i) { return q2[i]; };
        out1[o] = np_pairwise(f1, ap, n);
        out2[o] = np_pairwise(f2, ap, n);
    } else {
        const double a1 = l1[o], a2 = l2[o];
        auto f1 = [&](int i) { return q1[i] * a1; };
        auto f2 = [&](int i) { return q2[i] * a2; };
        const double s1 = np_pairwise(f1, ap, n), s2 = np_pairwise(f2, ap, n);
        out1[o] = c1 * s1 + c2 * s2;
    }
}

// The K-stage dynamic programme over cut positions (MMQ:175-207; maximisation, np.argmax = first maximum) and the
// back-trace (MMQ:210-215).  One CTA per problem; Az[p][0..K] are the cluster boundaries in sorted order.
__global__ void __launch_bounds__(256)
mmi_dp_kernel(const double *__restrict__ Tall, int M, int K, int W, int32_t *__restrict__ ws_lm, int32_t *__restrict__ Az_out) {
    extern __shared__ double sm_d[];
    const int p = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
    const double *T = Tall + (size_t)p * M * W;
    int32_t *lm = ws_lm + (size_t)p * (K + 1) * W;
    double *col0 = sm_d, *col1 = sm_d + W;
    for (int i = tid; i < W; i += nth) col0[i] = T[i];          // state[:,1] = table[0, 1 .. W]
    __syncthreads();
    double *prev = col0, *cur = col1;
    for (int z = 2; z <= K; ++z) {
        const int a_lo = z < K ? z : M, cnt = z < K ? W : 1;
        for (int r = tid; r < cnt; r += nth) {
            const int a = a_lo + r;
            int best_ap = z - 1;
            double best = prev[0] + T[(size_t)(z - 1) * W + (a - z)];
            for (int ap = z; ap <= a - 1; ++ap) {
                const double v = prev[ap - (z - 1)] + T[(size_t)ap * W + (a - ap - 1)];
                if (v > best) { best = v; best_ap = ap; }
            }
            const int row = z < K ? r : W - 1;
            cur[row] = best;
            lm[(size_t)z * W + row] = best_ap;
        }
        __syncthreads();
        double *t = prev; prev = cur; cur = t;
    }
    if (tid == 0) {
        int32_t *Az = Az_out + (size_t)p * (K + 1);
        Az[0] = 0;
        Az[K] = M;
        if (K >= 2) {
            int opt = lm[(size_t)K * W + (W - 1)];
            Az[K - 1] = opt;
            for (int z = K - 1; z >= 2; --z) { opt = lm[(size_t)z * W + (opt - z)]; Az[z - 1] = opt; }
        }
    }
}

}  // namespace pb
