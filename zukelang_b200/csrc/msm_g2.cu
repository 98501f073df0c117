// G2 instantiation of the MSM pipeline + its extern "C" entry points.
#include <string.h>
#include "msm_api.cuh"

namespace zk {
template struct BaseTable<G2Traits>;
template void finalize_points<G2Traits>(const XYZZ<Fp2>*, int, uint8_t*, cudaStream_t);
}  // namespace zk

extern "C" {
int zk_g2_msm(const uint8_t* bases, const uint8_t* inf_flags, const uint8_t* scalars, size_t n, uint8_t* out) {
  return zk::api_msm_oneshot<G2Traits>(bases, inf_flags, scalars, n, out);
}
int zk_g2_table_load(const uint8_t* bases, const uint8_t* inf_flags, size_t n, int precompute, int window_bits,
                     uint64_t* handle) {
  return zk::api_table_load<G2Traits>(bases, inf_flags, n, precompute, window_bits, handle);
}
int zk_g2_table_msm(uint64_t handle, const uint8_t* scalars, size_t n, uint8_t* out) {
  return zk::api_table_msm<G2Traits>(handle, scalars, n, out);
}
int zk_g2_table_msm_dev(uint64_t handle, const void* d_scalars, size_t n, void* d_out, void* stream) {
  return zk::api_table_msm_dev<G2Traits>(handle, d_scalars, n, d_out, stream);
}
int zk_g2_fixed_base_mul(const uint8_t* scalars, size_t n, uint8_t* out) {
  return zk::api_fixed_base_mul<G2Traits>(scalars, n, out);
}
int zk_g2_sum(const uint8_t* points, size_t k, uint8_t* out) { return zk::api_sum<G2Traits>(points, k, out); }
int zk_g2_sum_dev(const void* d_points, size_t k, void* d_out, void* stream) {
  return zk::api_sum_dev<G2Traits>(d_points, k, d_out, stream);
}
int zk_g2_sum_strided_dev(const void* d_points, size_t k, size_t batch, void* d_out, void* stream) {
  return zk::api_sum_strided_dev<G2Traits>(d_points, k, batch, d_out, stream);
}
int zk_g2_table_msm_batch(uint64_t handle, const uint8_t* const* scalars, size_t n, size_t count, uint8_t* out) {
  return zk::api_table_msm_batch<G2Traits>(handle, scalars, n, count, out);
}
}
