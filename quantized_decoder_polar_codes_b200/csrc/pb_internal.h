// Internal structures shared by the host side (pb_capi.cu) and the kernels.  Not part of the ABI.
#pragma once
#include <cstdint>

// Dynamic shared memory / kernel handles.  PB_HOST_EMU (tools/emu: CPU fibers, development only) swaps them for
// host equivalents; the product build never defines it.
#ifdef PB_HOST_EMU
#define PB_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(pb_emu::dyn_smem())
#define PB_KFN(...) pb_emu::reg(__VA_ARGS__)
#else
#define PB_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#define PB_KFN(...) ((const void *)(__VA_ARGS__))
#endif

namespace pb {

constexpr int kMaxL = 32;
constexpr int kMaxLog = 12;  // N <= 4096

enum Op : uint8_t { OP_F = 0, OP_G = 1, OP_C = 2, OP_LEAF = 3, OP_R0 = 4, OP_R1 = 5, OP_REP = 6, OP_SPC = 7 };
enum Domain : int { DOM_LUT = 0, DOM_FLOAT = 1, DOM_UNIFORM = 2, DOM_LLOYD = 3 };

// One entry of the compiled tree walk.  The walk of every reference decoder is data-independent
// (SURVEY.md section 0), so it is replayed once on the host at pd_create and the kernels just run the list.
struct Step {
    uint8_t op;
    uint8_t depth;   // depth of the node the op belongs to (OP_LEAF: n)
    uint8_t flag;    // OP_LEAF: 1 = frozen
    uint8_t pad;
    uint32_t node;   // index of the node within its depth
};

// Per internal node (heap id p = (1<<depth)+index-1): where its tables live in the uint8 pool.
struct NodeTab {
    uint32_t f_off, g_off;          // byte offsets into Dev::lut
    uint32_t f_sz, g_sz;            // qa*qb of one table (g: one u-plane)
    uint32_t f_pstride, g_pstride;  // 0 = one table shared by all positions, else bytes between positions
    uint16_t f_qb, g_qb;
    uint32_t pad;
};

struct Dev {
    int kind, N, n, K, A, L, Kout;
    int domain, list, fast, ca;
    double pm_init;
    const Step *steps;
    int n_steps;
    const int32_t *info_pos;  // [K] ascending positions of the non-frozen bits
    // LUT family
    const uint8_t *lut;
    const NodeTab *tabs;      // [N-1]
    const double *llr;
    const uint32_t *llr_off;  // [levels*N]
    int root_qa, root_qb;
    // uniform family: M_f = mf_mul * r_f, M_g = mg_mul * r_g  (PD/src/SCUniformQuantizedDecoder.cpp:56,72)
    const double *r_f, *r_g;
    double mf_mul, mg_mul;
    // Lloyd family
    const double *bnd_f, *bnd_g, *rec_f, *rec_g;
    int nb, nr;
    // CRC (CA kinds): remainder width, number of compared bits, generator taps below the leading term
    int crc_n, crc_check;
    uint32_t crc_taps;
    // generic-kernel workspace geometry
    int r1_tmax;  // largest R1 node in the schedule (list Fast kinds), 0 if none
    // blind-detection kinds (PD_BD_*), per call: metric / pass outputs and the RNTI mask; all null otherwise
    int bd;                    // 0 none, 1 D-metric (DMetric.cpp), 2 CA-SCL with RNTI (CASCLWithRNTI.cpp)
    const int32_t *bd_rnti;
    int bd_rnti_len;
    double *bd_metric;         // [B]
    uint8_t *bd_pass;          // [B]
};

}  // namespace pb
