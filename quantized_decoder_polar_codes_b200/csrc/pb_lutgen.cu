// Object of the lookup-table design kernel (pb_lutgen.cuh); pd_optls_quantize in pb_capi.cu launches it through this.
#include "pb_lutgen.cuh"

namespace pb {
cudaError_t launch_optls(int n_problems, size_t smem, const double *density, const double *quanta, const int32_t *M, long long stride, int K,
                         double *out_density, double *out_quanta, int32_t *out_lut, double *T, int32_t *lm, long long t_stride, long long lm_stride) {
    optls_kernel<<<n_problems, 256, smem>>>(density, quanta, M, stride, K, out_density, out_quanta, out_lut, T, lm, t_stride, lm_stride);
    return cudaGetLastError();
}
}  // namespace pb
