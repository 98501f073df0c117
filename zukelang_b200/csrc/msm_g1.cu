// G1 instantiation of the MSM pipeline + its extern "C" entry points.
#include <string.h>
#include "msm_api.cuh"

namespace zk {
__global__ void k_check_scalars(const uint32_t* __restrict__ scalars, uint32_t n, int* __restrict__ err) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr s;
  for (int j = 0; j < 8; j++) s.v[j] = scalars[8 * (size_t)i + j];
  if (!s.is_canonical_raw()) atomicExch(err, 1);
}
template struct BaseTable<G1Traits>;
template void finalize_points<G1Traits>(const XYZZ<Fp>*, int, uint8_t*, cudaStream_t);
}  // namespace zk

extern "C" {
int zk_g1_msm(const uint8_t* bases, const uint8_t* inf_flags, const uint8_t* scalars, size_t n, uint8_t* out) {
  return zk::api_msm_oneshot<G1Traits>(bases, inf_flags, scalars, n, out);
}
int zk_g1_table_load(const uint8_t* bases, const uint8_t* inf_flags, size_t n, int precompute, int window_bits,
                     uint64_t* handle) {
  return zk::api_table_load<G1Traits>(bases, inf_flags, n, precompute, window_bits, handle);
}
int zk_g1_table_msm(uint64_t handle, const uint8_t* scalars, size_t n, uint8_t* out) {
  return zk::api_table_msm<G1Traits>(handle, scalars, n, out);
}
int zk_g1_table_msm_dev(uint64_t handle, const void* d_scalars, size_t n, void* d_out, void* stream) {
  return zk::api_table_msm_dev<G1Traits>(handle, d_scalars, n, d_out, stream);
}
int zk_g1_fixed_base_mul(const uint8_t* scalars, size_t n, uint8_t* out) {
  return zk::api_fixed_base_mul<G1Traits>(scalars, n, out);
}
int zk_g1_sum(const uint8_t* points, size_t k, uint8_t* out) { return zk::api_sum<G1Traits>(points, k, out); }
int zk_g1_sum_dev(const void* d_points, size_t k, void* d_out, void* stream) {
  return zk::api_sum_dev<G1Traits>(d_points, k, d_out, stream);
}
int zk_g1_sum_strided_dev(const void* d_points, size_t k, size_t batch, void* d_out, void* stream) {
  return zk::api_sum_strided_dev<G1Traits>(d_points, k, batch, d_out, stream);
}
int zk_g1_table_msm_batch(uint64_t handle, const uint8_t* const* scalars, size_t n, size_t count, uint8_t* out) {
  return zk::api_table_msm_batch<G1Traits>(handle, scalars, n, count, out);
}
}
