// Object of the lookup-table design kernel (pb_lutgen.cuh); pd_optls_quantize in pb_capi.cu launches it through this.
#include "pb_lutgen.cuh"

namespace pb {
cudaError_t launch_optls(int n_problems, size_t smem, const double *density, const double *quanta, const int32_t *M, long long stride, int K,
                         double *out_density, double *out_quanta, int32_t *out_lut, double *T, int32_t *lm, long long t_stride, long long lm_stride) {
    optls_kernel<<<n_problems, 256, smem>>>(density, quanta, M, stride, K, out_density, out_quanta, out_lut, T, lm, t_stride, lm_stride);
    return cudaGetLastError();
}
cudaError_t launch_mmi_table(int P, int M, int W, int mode, const double *p1, const double *p2, const double *l1, const double *l2,
                             double c1, double c2, double *out1, double *out2) {
    const long long entries = (long long)M * W;
    mmi_table_kernel<<<dim3((unsigned)((entries + 255) / 256), (unsigned)P), 256>>>(p1, p2, M, W, mode, l1, l2, c1, c2, out1, out2);
    return cudaGetLastError();
}
cudaError_t launch_mmi_dp(int P, int M, int K, int W, const double *T, int32_t *lm, int32_t *Az) {
    mmi_dp_kernel<<<P, 256, 2 * (size_t)W * sizeof(double)>>>(T, M, K, W, lm, Az);
    return cudaGetLastError();
}
}  // namespace pb
