// Backbone tail -> head hand-off: the last two operations of every reference backbone,
//     x = self.features(x)      # nn.BatchNorm1d(feat_dim, eps=1e-05)     resnet_arcface.py:99,151 / resnet_std.py:201
//     x = F.normalize(x)        # x / max(||x||_2, 1e-12) per row          resnet_arcface.py:151 / resnet_std.py:202 /
//                                                                          mobilefacenet_def.py:113-114 (no BatchNorm1d there)
// and their backward, so that the unit-norm fp32 embeddings the head consumes (ffc.py:157/209) are produced by one or two
// launches, optionally straight into the buffer the head (or its all-gather) reads: `p_stride` lets the rows land in a packed
// staging buffer.  Eager PyTorch runs ~6 kernels forward and ~12 backward for the same work.
//
// [B, D] fp32 is 2 MiB at B = 1024, D = 512: every kernel here is launch-latency bound; the byte counts are in DESIGN.md 4.6.
//   forward   BN(train): column statistics over (32 columns x row slab) CTAs -- per-slab mean and centred second moment, combined in
//                        fp64 by the last CTA of each column group; running statistics updated as torch does (unbiased variance,
//                        momentum) -- then the row kernel
//             BN(eval): invstd of the running statistics, then the row kernel;  plain normalise: the row kernel only (one warp per row)
//   backward  row kernel: dy = (dp - p (p . dp)) / max(||y||, eps)        (dp / eps on a clamped row, as autograd does); with BN it
//             also leaves per-slab column sums of dy and dy * xhat, and a column kernel finishes in place on dy:
//                 dx = gamma * invstd * (dy - mean_b(dy) - xhat * mean_b(dy * xhat))     (train; eval drops the two means)
//             dbeta / dgamma are the two column sums.  All reductions run in a fixed order: results are run-to-run bit-identical.
#include <algorithm>

#include "ffc_common.cuh"

namespace ffc {

constexpr float TAIL_NORM_EPS = 1e-12f;   // F.normalize default eps
constexpr int TAIL_ROW_WARPS = 4;

__device__ __forceinline__ float tail_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int TAIL_COL_WARPS = 8;      // column-organised kernels: CTA = 32 columns x 8 warps, each warp a strided share of the slab's rows

// Fixed-order sum over the CTA's 8 warps of one value per (warp, column); result valid in every thread (indexed by threadIdx.x).
template <typename T>
__device__ __forceinline__ T tail_col_reduce(T v, T (*red)[33]) {
  red[threadIdx.y][threadIdx.x] = v;
  __syncthreads();
  T t = red[0][threadIdx.x];
#pragma unroll
  for (int i = 1; i < TAIL_COL_WARPS; ++i) t += red[i][threadIdx.x];
  __syncthreads();
  return t;
}

// ---- forward, BatchNorm1d in train(): per-column batch statistics ----
// grid (ceil(D / 32), n_slabs): CTA (g, k) reduces rows [k * slab_rows, ...) of columns [32 g, 32 g + 32) to the slab's
// (mean_k, M2_k) -- two passes, the second about the slab mean (L1 hit) -- and the LAST CTA of a column group to arrive combines the
// slabs in slab order (Chan's update in fp64: exact to fp32 rounding, and independent of the arrival order, so run-to-run
// bit-identical), writes mean / invstd and updates the running statistics.  counters[g] is zero on entry and is left zero.
__global__ void __launch_bounds__(32 * TAIL_COL_WARPS) tail_col_stats_kernel(const float* __restrict__ x, int B, int D, int slab_rows, float eps,
                                                                              float momentum, float* __restrict__ running_mean,
                                                                              float* __restrict__ running_var, float* __restrict__ save_mean,
                                                                              float* __restrict__ save_invstd, float* __restrict__ partial,
                                                                              unsigned int* __restrict__ counters) {
  __shared__ float red[TAIL_COL_WARPS][33];
  __shared__ unsigned int ticket;
  const int c = blockIdx.x * 32 + threadIdx.x, n_slabs = gridDim.y, k = blockIdx.y;
  const bool ok = c < D;
  const int r0 = k * slab_rows, r1 = min(B, r0 + slab_rows);
  float s = 0.f;
  if (ok)
    for (int r = r0 + threadIdx.y; r < r1; r += TAIL_COL_WARPS) s += x[(int64_t)r * D + c];
  const float mean_k = tail_col_reduce(s, red) / (float)(r1 - r0);
  float q = 0.f;
  if (ok)
    for (int r = r0 + threadIdx.y; r < r1; r += TAIL_COL_WARPS) {
      const float d = x[(int64_t)r * D + c] - mean_k;
      q += d * d;
    }
  const float m2_k = tail_col_reduce(q, red);
  if (ok && threadIdx.y == 0) {
    partial[((int64_t)k * 2 + 0) * D + c] = mean_k;
    partial[((int64_t)k * 2 + 1) * D + c] = m2_k;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) ticket = atomicAdd(&counters[blockIdx.x], 1u);
  __syncthreads();
  if (ticket != (unsigned int)(n_slabs - 1)) return;
  __threadfence();
  // last CTA of the column group: warp y takes slabs y, y + 8, ... (loads batched), the 8 warp sums are added in warp order
  __shared__ double redd[TAIL_COL_WARPS][33];
  double ms = 0.0;
  if (ok) {
#pragma unroll 8
    for (int j = threadIdx.y; j < n_slabs; j += TAIL_COL_WARPS) {
      const int nj = min(B, (j + 1) * slab_rows) - j * slab_rows;
      ms += (double)nj * (double)__ldcg(&partial[((int64_t)j * 2 + 0) * D + c]);
    }
  }
  const double mean = tail_col_reduce(ms, redd) / (double)B;
  double m2s = 0.0;
  if (ok) {
#pragma unroll 8
    for (int j = threadIdx.y; j < n_slabs; j += TAIL_COL_WARPS) {
      const int nj = min(B, (j + 1) * slab_rows) - j * slab_rows;
      const double d = (double)__ldcg(&partial[((int64_t)j * 2 + 0) * D + c]) - mean;
      m2s += (double)__ldcg(&partial[((int64_t)j * 2 + 1) * D + c]) + (double)nj * d * d;
    }
  }
  const double m2 = tail_col_reduce(m2s, redd);
  if (threadIdx.y == 0) {
    if (ok) {
      const float var_b = (float)(m2 / (double)B);                                  // biased: what normalises the batch
      save_mean[c] = (float)mean;
      save_invstd[c] = 1.0f / sqrtf(var_b + eps);
      if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
      if (running_var) {
        const float var_u = (float)(m2 / (double)(B > 1 ? B - 1 : 1));              // unbiased: what the running estimate tracks
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * var_u;
      }
    }
    if (threadIdx.x == 0) counters[blockIdx.x] = 0u;
  }
}

// ---- BatchNorm1d in eval(): invstd of the running statistics ----
__global__ void __launch_bounds__(256) tail_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var, int D,
                                                              float eps, float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < D) {
    save_mean[c] = running_mean[c];
    save_invstd[c] = 1.0f / sqrtf(running_var[c] + eps);
  }
}

struct TailAffine {
  const float* gamma;   // may be null (1)
  const float* beta;    // may be null (0)
  const float* mean;
  const float* invstd;
};

template <bool BN>
__device__ __forceinline__ float tail_y(float x, int c, const TailAffine& a) {
  if (!BN) return x;
  const float xh = (x - a.mean[c]) * a.invstd[c];
  return xh * (a.gamma ? a.gamma[c] : 1.0f) + (a.beta ? a.beta[c] : 0.0f);
}

template <bool BN>
__device__ __forceinline__ float4 tail_y4(float4 v, int c, const TailAffine& a) {
  if (!BN) return v;
  const float4 m = *reinterpret_cast<const float4*>(a.mean + c), is = *reinterpret_cast<const float4*>(a.invstd + c);
  float4 g = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a.gamma) g = *reinterpret_cast<const float4*>(a.gamma + c);
  if (a.beta) b = *reinterpret_cast<const float4*>(a.beta + c);
  return make_float4((v.x - m.x) * is.x * g.x + b.x, (v.y - m.y) * is.y * g.y + b.y, (v.z - m.z) * is.z * g.z + b.z,
                     (v.w - m.w) * is.w * g.w + b.w);
}

// ---- forward row kernel: one warp per row; y recomputed in the second sweep of the row (L1 hit) instead of being held ----
template <bool BN, bool VEC>
__global__ void __launch_bounds__(32 * TAIL_ROW_WARPS) tail_rows_fwd_kernel(const float* __restrict__ x, int B, int D, TailAffine a,
                                                                            float* __restrict__ p, int64_t p_stride, float* __restrict__ inv_norm) {
  const int row = blockIdx.x * TAIL_ROW_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* xr = x + (int64_t)row * D;
  float* pr = p + (int64_t)row * p_stride;
  float ss = 0.f;
  if (VEC) {
    for (int c = 4 * lane; c < D; c += 128) {
      const float4 y = tail_y4<BN>(*reinterpret_cast<const float4*>(xr + c), c, a);
      ss += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
    }
  } else {
    for (int c = lane; c < D; c += 32) {
      const float y = tail_y<BN>(xr[c], c, a);
      ss += y * y;
    }
  }
  ss = tail_warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), TAIL_NORM_EPS);
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
  if (VEC) {
    for (int c = 4 * lane; c < D; c += 128) {
      const float4 y = tail_y4<BN>(*reinterpret_cast<const float4*>(xr + c), c, a);
      *reinterpret_cast<float4*>(pr + c) = make_float4(y.x * inv, y.y * inv, y.z * inv, y.w * inv);
    }
  } else {
    for (int c = lane; c < D; c += 32) pr[c] = tail_y<BN>(xr[c], c, a) * inv;
  }
}

// ---- backward row kernel: gradient of the L2 normalisation ----
template <bool VEC>
__global__ void __launch_bounds__(32 * TAIL_ROW_WARPS) tail_rows_bwd_kernel(const float* __restrict__ p, int64_t p_stride, const float* __restrict__ dp,
                                                                            int64_t dp_stride, const float* __restrict__ inv_norm, int B, int D,
                                                                            float* __restrict__ dy) {
  const int row = blockIdx.x * TAIL_ROW_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* pr = p + (int64_t)row * p_stride;
  const float* gr = dp + (int64_t)row * dp_stride;
  float* out = dy + (int64_t)row * D;
  const float inv = inv_norm[row];
  // a row whose norm was clamped to eps (inv == 1/eps): autograd sends no gradient through the norm, dy = dp / eps
  const bool clamped = inv >= 1.0f / TAIL_NORM_EPS;
  float s = 0.f;
  if (!clamped) {
    if (VEC) {
      for (int c = 4 * lane; c < D; c += 128) {
        const float4 a = *reinterpret_cast<const float4*>(pr + c), g = *reinterpret_cast<const float4*>(gr + c);
        s += a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w;
      }
    } else {
      for (int c = lane; c < D; c += 32) s += pr[c] * gr[c];
    }
    s = tail_warp_sum(s);
  }
  if (VEC) {
    for (int c = 4 * lane; c < D; c += 128) {
      const float4 a = *reinterpret_cast<const float4*>(pr + c), g = *reinterpret_cast<const float4*>(gr + c);
      *reinterpret_cast<float4*>(out + c) = make_float4((g.x - a.x * s) * inv, (g.y - a.y * s) * inv, (g.z - a.z * s) * inv, (g.w - a.w * s) * inv);
    }
  } else {
    for (int c = lane; c < D; c += 32) out[c] = (gr[c] - pr[c] * s) * inv;
  }
}

// ---- backward row kernel with BatchNorm1d behind the normalisation: dy as above, plus the slab's column sums of dy and dy * xhat ----
// CTA = 8 warps over rows [blockIdx.x * rows_per_cta, ...); every warp accumulates the rows it handles into its own shared-memory
// copy of the two column vectors (a lane always owns the same columns: no atomics, fixed order), the CTA adds the 8 copies in warp
// order and writes partial[blockIdx.x][2][D].
template <bool VEC>
__global__ void __launch_bounds__(32 * TAIL_COL_WARPS) tail_rows_bwd_bn_kernel(const float* __restrict__ p, int64_t p_stride, const float* __restrict__ dp,
                                                                                int64_t dp_stride, const float* __restrict__ inv_norm,
                                                                                const float* __restrict__ x, const float* __restrict__ mean,
                                                                                const float* __restrict__ invstd, int B, int D, int rows_per_cta,
                                                                                float* __restrict__ dy, float* __restrict__ partial) {
  extern __shared__ float acc[];   // [8 warps][2][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* a0 = acc + (int64_t)(warp * 2) * D;
  float* a1 = a0 + D;
  if (VEC) {
    for (int c = 4 * lane; c < D; c += 128) {
      *reinterpret_cast<float4*>(a0 + c) = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(a1 + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    for (int c = lane; c < D; c += 32) a0[c] = a1[c] = 0.f;
  }
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(B, r0 + rows_per_cta);
  for (int row = r0 + warp; row < r1; row += TAIL_COL_WARPS) {
    const float* pr = p + (int64_t)row * p_stride;
    const float* gr = dp + (int64_t)row * dp_stride;
    const float* xr = x + (int64_t)row * D;
    float* out = dy + (int64_t)row * D;
    const float inv = inv_norm[row];
    const bool clamped = inv >= 1.0f / TAIL_NORM_EPS;
    float s = 0.f;
    if (!clamped) {
      if (VEC) {
        for (int c = 4 * lane; c < D; c += 128) {
          const float4 a = *reinterpret_cast<const float4*>(pr + c), g = *reinterpret_cast<const float4*>(gr + c);
          s += a.x * g.x + a.y * g.y + a.z * g.z + a.w * g.w;
        }
      } else {
        for (int c = lane; c < D; c += 32) s += pr[c] * gr[c];
      }
      s = tail_warp_sum(s);
    }
    if (VEC) {
      for (int c = 4 * lane; c < D; c += 128) {
        const float4 a = *reinterpret_cast<const float4*>(pr + c), g = *reinterpret_cast<const float4*>(gr + c);
        const float4 xv = *reinterpret_cast<const float4*>(xr + c), m = *reinterpret_cast<const float4*>(mean + c),
                     is = *reinterpret_cast<const float4*>(invstd + c);
        const float4 d = make_float4((g.x - a.x * s) * inv, (g.y - a.y * s) * inv, (g.z - a.z * s) * inv, (g.w - a.w * s) * inv);
        *reinterpret_cast<float4*>(out + c) = d;
        float4 u = *reinterpret_cast<float4*>(a0 + c), v = *reinterpret_cast<float4*>(a1 + c);
        u.x += d.x, u.y += d.y, u.z += d.z, u.w += d.w;
        v.x += d.x * ((xv.x - m.x) * is.x), v.y += d.y * ((xv.y - m.y) * is.y), v.z += d.z * ((xv.z - m.z) * is.z), v.w += d.w * ((xv.w - m.w) * is.w);
        *reinterpret_cast<float4*>(a0 + c) = u;
        *reinterpret_cast<float4*>(a1 + c) = v;
      }
    } else {
      for (int c = lane; c < D; c += 32) {
        const float d = (gr[c] - pr[c] * s) * inv;
        out[c] = d;
        a0[c] += d;
        a1[c] += d * ((xr[c] - mean[c]) * invstd[c]);
      }
    }
  }
  __syncthreads();
  float* part = partial + (int64_t)blockIdx.x * 2 * D;
  for (int v = threadIdx.x; v < 2 * D; v += 32 * TAIL_COL_WARPS) {
    float t = acc[v];
#pragma unroll
    for (int w = 1; w < TAIL_COL_WARPS; ++w) t += acc[(int64_t)w * 2 * D + v];
    part[v] = t;
  }
}

// ---- backward column kernel (BatchNorm1d), in place on dy: dx = gamma * invstd * (dy - mean_b(dy) - xhat * mean_b(dy * xhat)) ----
// grid (ceil(D / 32), ceil(B / rows_per_cta)).  Every CTA re-adds the row kernel's slab partials of its 32 columns (fp64, slab order:
// identical in every CTA and every run); the CTAs of the first row slab also write dbeta / dgamma.
__global__ void __launch_bounds__(32 * TAIL_COL_WARPS) tail_cols_dx_kernel(const float* __restrict__ x, float* __restrict__ dy_dx, int B, int D,
                                                                            int rows_per_cta, TailAffine a, int train,
                                                                            const float* __restrict__ partial, int n_partial,
                                                                            float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ double red[TAIL_COL_WARPS][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const bool ok = c < D;
  double s1 = 0.0, s2 = 0.0;
  if (ok) {
#pragma unroll 8
    for (int j = threadIdx.y; j < n_partial; j += TAIL_COL_WARPS) {
      s1 += (double)partial[((int64_t)j * 2 + 0) * D + c];
      s2 += (double)partial[((int64_t)j * 2 + 1) * D + c];
    }
  }
  s1 = tail_col_reduce(s1, red);
  s2 = tail_col_reduce(s2, red);
  if (!ok) return;
  if (blockIdx.y == 0 && threadIdx.y == 0) {
    if (dbeta) dbeta[c] = (float)s1;
    if (dgamma) dgamma[c] = (float)s2;
  }
  const float mean = a.mean[c], invstd = a.invstd[c];
  const float w = (a.gamma ? a.gamma[c] : 1.0f) * invstd;
  const float m1 = train ? (float)(s1 / (double)B) : 0.f, m2 = train ? (float)(s2 / (double)B) : 0.f;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(B, r0 + rows_per_cta);
  for (int r = r0 + threadIdx.y; r < r1; r += TAIL_COL_WARPS) {
    const int64_t i = (int64_t)r * D + c;
    const float xh = (x[i] - mean) * invstd;
    dy_dx[i] = w * (dy_dx[i] - m1 - xh * m2);
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// slab geometry (shared by ffc_tail_workspace_bytes and the launches)
struct TailGeom {
  int fwd_slab_rows, fwd_slabs;      // column statistics: <= 64 slabs of >= 32 rows
  int bwd_rows_per_cta, bwd_ctas;    // backward row kernel: 16 rows per CTA up to 2048 rows, <= 128 CTAs beyond
  int64_t counter_bytes, partial_floats;
};
static TailGeom tail_geom(int B, int D) {
  TailGeom g;
  const int want = std::min((B + 31) / 32, 64);
  g.fwd_slab_rows = (B + want - 1) / want;
  g.fwd_slabs = (B + g.fwd_slab_rows - 1) / g.fwd_slab_rows;
  g.bwd_rows_per_cta = B <= 2048 ? 2 * TAIL_COL_WARPS : ((B + 127) / 128 + TAIL_COL_WARPS - 1) / TAIL_COL_WARPS * TAIL_COL_WARPS;
  g.bwd_ctas = (B + g.bwd_rows_per_cta - 1) / g.bwd_rows_per_cta;
  g.counter_bytes = ((int64_t)((D + 31) / 32) * 4 + 255) / 256 * 256;
  g.partial_floats = (int64_t)std::max(g.fwd_slabs, g.bwd_ctas) * 2 * D;
  return g;
}

}  // namespace ffc

using namespace ffc;

extern "C" int ffc_tail_workspace_bytes(int n_rows, int feat_dim, int64_t* bytes_out) {
  FFC_REQUIRE(n_rows >= 1 && feat_dim >= 1 && bytes_out, "ffc_tail_workspace_bytes: bad arguments");
  const TailGeom g = tail_geom(n_rows, feat_dim);
  *bytes_out = g.counter_bytes + g.partial_floats * 4;
  return FFC_OK;
}

static int tail_check(const ffc_tail_args* a, const char* who, bool backward) {
  FFC_REQUIRE(a != nullptr, "%s: null arguments", who);
  FFC_REQUIRE(a->n_rows >= 1 && a->feat_dim >= 1, "%s: n_rows %d / feat_dim %d must be >= 1", who, a->n_rows, a->feat_dim);
  FFC_REQUIRE(a->mode == FFC_TAIL_NORMALIZE || a->mode == FFC_TAIL_BN_EVAL || a->mode == FFC_TAIL_BN_TRAIN, "%s: unknown mode %d", who, a->mode);
  FFC_REQUIRE(a->x && a->p && a->inv_norm, "%s: x, p and inv_norm are required", who);
  FFC_REQUIRE(a->x != a->p, "%s: p must not alias x", who);
  FFC_REQUIRE(a->p_stride >= a->feat_dim, "%s: p_stride %lld < feat_dim %d", who, (long long)a->p_stride, a->feat_dim);
  if (a->mode != FFC_TAIL_NORMALIZE) {
    FFC_REQUIRE(a->save_mean && a->save_invstd, "%s: save_mean / save_invstd are required with BatchNorm1d", who);
    if (a->mode == FFC_TAIL_BN_EVAL) FFC_REQUIRE(a->running_mean && a->running_var, "%s: eval mode needs the running statistics", who);
    // torch raises "Expected more than 1 value per channel when training" (functional.py:_verify_batch_size)
    if (a->mode == FFC_TAIL_BN_TRAIN) FFC_REQUIRE(a->n_rows > 1, "%s: BatchNorm1d in training mode needs more than 1 row", who);
    if (backward || a->mode == FFC_TAIL_BN_TRAIN) {
      int64_t need = 0;
      ffc_tail_workspace_bytes(a->n_rows, a->feat_dim, &need);
      FFC_REQUIRE(a->workspace && a->workspace_bytes >= need && aligned16(a->workspace),
                  "%s: workspace of %lld bytes (16-byte aligned, zero-filled before its first use) is required", who, (long long)need);
    }
  }
  return FFC_OK;
}

static bool tail_vec(const ffc_tail_args* a) {
  bool v = a->feat_dim % 4 == 0 && a->p_stride % 4 == 0 && aligned16(a->x) && aligned16(a->p);
  if (a->mode != FFC_TAIL_NORMALIZE)
    v = v && aligned16(a->save_mean) && aligned16(a->save_invstd) && (!a->gamma || aligned16(a->gamma)) && (!a->beta || aligned16(a->beta));
  return v;
}

extern "C" int ffc_tail_forward(const ffc_tail_args* a, void* stream) {
  if (int rc = tail_check(a, "ffc_tail_forward", false)) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int B = a->n_rows, D = a->feat_dim;
  if (a->mode == FFC_TAIL_BN_TRAIN) {
    const TailGeom g = tail_geom(B, D);
    unsigned int* counters = reinterpret_cast<unsigned int*>(a->workspace);
    float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(a->workspace) + g.counter_bytes);
    tail_col_stats_kernel<<<dim3((D + 31) / 32, g.fwd_slabs), dim3(32, TAIL_COL_WARPS), 0, s>>>(a->x, B, D, g.fwd_slab_rows, a->eps, a->momentum,
                                                                                                 a->running_mean, a->running_var, a->save_mean,
                                                                                                 a->save_invstd, partial, counters);
    FFC_LAUNCH_CHECK();
  } else if (a->mode == FFC_TAIL_BN_EVAL) {
    tail_eval_stats_kernel<<<(D + 255) / 256, 256, 0, s>>>(a->running_mean, a->running_var, D, a->eps, a->save_mean, a->save_invstd);
    FFC_LAUNCH_CHECK();
  }
  const TailAffine af{a->gamma, a->beta, a->save_mean, a->save_invstd};
  const int grid = (B + TAIL_ROW_WARPS - 1) / TAIL_ROW_WARPS;
  const bool bn = a->mode != FFC_TAIL_NORMALIZE, vec = tail_vec(a);
#define FFC_TAIL_FWD(BN, VEC) tail_rows_fwd_kernel<BN, VEC><<<grid, 32 * TAIL_ROW_WARPS, 0, s>>>(a->x, B, D, af, a->p, a->p_stride, a->inv_norm)
  if (bn && vec) FFC_TAIL_FWD(true, true);
  else if (bn) FFC_TAIL_FWD(true, false);
  else if (vec) FFC_TAIL_FWD(false, true);
  else FFC_TAIL_FWD(false, false);
#undef FFC_TAIL_FWD
  FFC_LAUNCH_CHECK();
  return FFC_OK;
}

extern "C" int ffc_tail_backward(const ffc_tail_args* a, const float* dp_dev, int64_t dp_stride, float* dx_dev, float* dgamma_dev, float* dbeta_dev,
                                 void* stream) {
  if (int rc = tail_check(a, "ffc_tail_backward", true)) return rc;
  FFC_REQUIRE(dp_dev && dx_dev, "ffc_tail_backward: dp and dx are required");
  FFC_REQUIRE(dp_stride >= a->feat_dim, "ffc_tail_backward: dp_stride %lld < feat_dim %d", (long long)dp_stride, a->feat_dim);
  FFC_REQUIRE(a->mode != FFC_TAIL_NORMALIZE || (!dgamma_dev && !dbeta_dev), "ffc_tail_backward: no affine parameters in FFC_TAIL_NORMALIZE mode");
  cudaStream_t s = (cudaStream_t)stream;
  const int B = a->n_rows, D = a->feat_dim;
  bool vec = D % 4 == 0 && a->p_stride % 4 == 0 && dp_stride % 4 == 0 && aligned16(a->p) && aligned16(dp_dev) && aligned16(dx_dev);
  if (a->mode == FFC_TAIL_NORMALIZE) {
    const int grid = (B + TAIL_ROW_WARPS - 1) / TAIL_ROW_WARPS;
    if (vec)
      tail_rows_bwd_kernel<true><<<grid, 32 * TAIL_ROW_WARPS, 0, s>>>(a->p, a->p_stride, dp_dev, dp_stride, a->inv_norm, B, D, dx_dev);
    else
      tail_rows_bwd_kernel<false><<<grid, 32 * TAIL_ROW_WARPS, 0, s>>>(a->p, a->p_stride, dp_dev, dp_stride, a->inv_norm, B, D, dx_dev);
    FFC_LAUNCH_CHECK();
    return FFC_OK;
  }
  const TailGeom g = tail_geom(B, D);
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(a->workspace) + g.counter_bytes);
  vec = vec && aligned16(a->x) && aligned16(a->save_mean) && aligned16(a->save_invstd);
  const size_t smem = (size_t)TAIL_COL_WARPS * 2 * D * sizeof(float);
  FFC_REQUIRE(smem <= 227 * 1024, "ffc_tail_backward: feat_dim %d exceeds the shared-memory accumulators (max %d)", D, 227 * 1024 / (TAIL_COL_WARPS * 8));
  if (vec) {
    if (smem > 48 * 1024) FFC_CUDA(cudaFuncSetAttribute(tail_rows_bwd_bn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tail_rows_bwd_bn_kernel<true><<<g.bwd_ctas, 32 * TAIL_COL_WARPS, smem, s>>>(a->p, a->p_stride, dp_dev, dp_stride, a->inv_norm, a->x, a->save_mean,
                                                                                a->save_invstd, B, D, g.bwd_rows_per_cta, dx_dev, partial);
  } else {
    if (smem > 48 * 1024) FFC_CUDA(cudaFuncSetAttribute(tail_rows_bwd_bn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tail_rows_bwd_bn_kernel<false><<<g.bwd_ctas, 32 * TAIL_COL_WARPS, smem, s>>>(a->p, a->p_stride, dp_dev, dp_stride, a->inv_norm, a->x, a->save_mean,
                                                                                 a->save_invstd, B, D, g.bwd_rows_per_cta, dx_dev, partial);
  }
  FFC_LAUNCH_CHECK();
  const TailAffine af{a->gamma, a->beta, a->save_mean, a->save_invstd};
  const int dx_rows = 64;
  tail_cols_dx_kernel<<<dim3((D + 31) / 32, (B + dx_rows - 1) / dx_rows), dim3(32, TAIL_COL_WARPS), 0, s>>>(
      a->x, dx_dev, B, D, dx_rows, af, a->mode == FFC_TAIL_BN_TRAIN ? 1 : 0, partial, g.bwd_ctas, dgamma_dev, dbeta_dev);
  FFC_LAUNCH_CHECK();
  return FFC_OK;
}
