// Gallery-network EMA (replaces ffc.py:139-145 `_momentum_update_gallery`): for every parameter
//     g = g * m + p * (1 - m)
// in ONE launch over all parameter tensors (the reference issues three eager kernels per tensor: 238 - 463 tensors per
// backbone).  HBM-bound: 12 bytes per parameter (read g, read p, write g).  The arithmetic reproduces the reference's
// three separate fp32 operations (two multiplies, one add, no FMA contraction), so the result is bit-identical.
#include <algorithm>

#include "ffc_common.cuh"

namespace ffc {

constexpr int EMA_CHUNK = 16384;   // elements per block

__global__ void __launch_bounds__(256) ema_update_kernel(const ffc_ema_chunk* __restrict__ table, float m, float one_minus_m) {
  const ffc_ema_chunk c = table[blockIdx.x];
  float* __restrict__ g = c.gallery;
  const float* __restrict__ p = c.probe;
  const int n = c.n;
  if (((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(p)) & 15u) == 0) {
    const int n4 = n >> 2;
    float4* g4 = reinterpret_cast<float4*>(g);
    const float4* p4 = reinterpret_cast<const float4*>(p);
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 a = g4[i];
      const float4 b = p4[i];
      a.x = __fadd_rn(__fmul_rn(a.x, m), __fmul_rn(b.x, one_minus_m));
      a.y = __fadd_rn(__fmul_rn(a.y, m), __fmul_rn(b.y, one_minus_m));
      a.z = __fadd_rn(__fmul_rn(a.z, m), __fmul_rn(b.z, one_minus_m));
      a.w = __fadd_rn(__fmul_rn(a.w, m), __fmul_rn(b.w, one_minus_m));
      g4[i] = a;
    }
    for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) g[i] = __fadd_rn(__fmul_rn(g[i], m), __fmul_rn(p[i], one_minus_m));
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) g[i] = __fadd_rn(__fmul_rn(g[i], m), __fmul_rn(p[i], one_minus_m));
  }
}

}  // namespace ffc

using namespace ffc;

extern "C" int ffc_ema_chunk_elems(void) { return EMA_CHUNK; }

extern "C" int ffc_ema_update(const ffc_ema_chunk* table_dev, int n_chunks, float m, float one_minus_m, void* stream) {
  FFC_REQUIRE(n_chunks >= 0 && (table_dev != nullptr || n_chunks == 0), "ffc_ema_update: bad arguments");
  if (n_chunks == 0) return FFC_OK;
  ema_update_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(table_dev, m, one_minus_m);
  FFC_LAUNCH_CHECK();
  return FFC_OK;
}
