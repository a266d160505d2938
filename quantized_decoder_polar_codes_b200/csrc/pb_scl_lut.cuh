// Specialised LUT SC/SCL kernels (placeholder until the warp-level kernel lands).
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include "pb_internal.h"

namespace pb {
struct FastPlan {
    bool ok = false;
    const char *name = "generic";
};
inline void plan_fast_lut(const Dev &, const std::vector<Step> &, const std::vector<NodeTab> &, const std::vector<uint8_t> &,
                          const int32_t *, FastPlan *p) { p->ok = false; }
inline int launch_fast_lut(const Dev &, const FastPlan &, const void *, int, long long, uint8_t *, cudaStream_t, int *, double *, int *, int) { return 0; }
inline void free_fast_plan(FastPlan *) {}
}  // namespace pb
