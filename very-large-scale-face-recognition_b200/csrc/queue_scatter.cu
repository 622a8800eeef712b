// Prototype-queue enqueue / rollback scatter (replaces ffc.py:179-182, 237-241, 255).
// One CTA per batch position: the CTA first scans the later positions for the same (row, col) pair
// -- "last occurrence wins", the reference's serial index_put behaviour -- and only a winner moves
// data: (optionally) save the old fp32 row, write the fp32 row with 16-byte stores and the bf16 mirror
// row next to it.  B*D*(4+4+2) bytes per pass: latency-bound by construction.
#include <algorithm>

#include "ffc_common.cuh"

namespace ffc {

// Is there a later position with the same (row, col)?  `packed`: the live positions come first and every position after the first
// padded one (cols < 0) is padded too -- the sharded head's lists, where a rank's own keys are compacted to the front of the R*B
// gathered positions -- so the scan stops there (at 8 ranks 7/8 of the list is padding: 80 us -> 10 us per scatter).
__device__ __forceinline__ bool later_duplicate(const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, int i, int B, bool packed) {
  const int32_t r = rows[i], c = cols[i];
  bool dup = false;
  for (int j0 = i + 1; j0 < B; j0 += blockDim.x) {
    const int j = j0 + threadIdx.x;
    const int32_t cj = j < B ? cols[j] : -1;
    dup |= (cj == c && rows[j < B ? j : i] == r);
    if (packed && __syncthreads_or(cj < 0)) break;
  }
  return __syncthreads_or(dup);
}

__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* dst, float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&lo);
  u.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(dst) = u;
}

__global__ void __launch_bounds__(128) queue_scatter_kernel(float* __restrict__ qf, __nv_bfloat16* __restrict__ qh, const int32_t* __restrict__ rows,
                                                            const int32_t* __restrict__ cols, const float* __restrict__ g, int B, int64_t Q, int D,
                                                            float* __restrict__ undo, const int32_t* __restrict__ src_row,
                                                            int32_t* __restrict__ ovl_map, int ovl_table) {
  const int i = blockIdx.x;
  if (cols[i] < 0) return;   // padded position (sharded callers)
  if (later_duplicate(rows, cols, i, B, src_row != nullptr)) return;
  const int64_t off = ((int64_t)rows[i] * Q + cols[i]) * D;
  if (ovl_map && threadIdx.x == 0) {
    // overlay of the sharded head's merged step (csrc/head.cu FinalizeArgs::ovl_map): where the sweep-time content of this (row, slot)
    // will still be found after later writes.  Table 0 (this enqueue's own source row, a rollback pass) always wins; table 1 (the
    // previous content this enqueue saves in undo[i], the commit pass that follows) only marks entries no rollback enqueue holds.
    int32_t* m = ovl_map + (int64_t)rows[i] * Q + cols[i];
    if (ovl_table == 0) *m = src_row ? src_row[i] : i;
    else if (*m < 0) *m = 0x40000000 | i;
  }
  const float4* src = reinterpret_cast<const float4*>(g + (int64_t)(src_row ? src_row[i] : i) * D);
  float4* dst = reinterpret_cast<float4*>(qf + off);
  float4* und = undo ? reinterpret_cast<float4*>(undo + (int64_t)i * D) : nullptr;
  for (int v = threadIdx.x; v < D / 4; v += blockDim.x) {
    if (und) und[v] = dst[v];
    const float4 x = src[v];
    dst[v] = x;
    if (qh) store_bf16x4(qh + off + 4 * v, x);
  }
}

__global__ void __launch_bounds__(128) queue_restore_kernel(float* __restrict__ qf, __nv_bfloat16* __restrict__ qh, const int32_t* __restrict__ rows,
                                                            const int32_t* __restrict__ cols, const float* __restrict__ undo, int B, int64_t Q, int D, int packed) {
  const int i = blockIdx.x;
  if (cols[i] < 0) return;
  if (later_duplicate(rows, cols, i, B, packed != 0)) return;
  const int64_t off = ((int64_t)rows[i] * Q + cols[i]) * D;
  const float4* src = reinterpret_cast<const float4*>(undo + (int64_t)i * D);
  float4* dst = reinterpret_cast<float4*>(qf + off);
  for (int v = threadIdx.x; v < D / 4; v += blockDim.x) {
    const float4 x = src[v];
    dst[v] = x;
    if (qh) store_bf16x4(qh + off + 4 * v, x);
  }
}

__global__ void __launch_bounds__(256) cast_bf16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    dst[i] = u;
  }
}

// Stable partition of the gathered gallery keys: the keys this rank owns (floor-mod identity hash, key mod R == rank) first,
// in global batch order, then the others; order[j] = source position of output j.  One CTA: two passes of a block-wide count.
__global__ void __launch_bounds__(1024) route_keys_kernel(const int64_t* __restrict__ keys, int n, int R, int rank, int64_t* __restrict__ keys_out,
                                                          int32_t* __restrict__ order, int32_t* __restrict__ n_mine_out) {
  __shared__ int warp_cnt[32];
  __shared__ int base_mine, base_other, total_mine;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto mine_of = [&](int64_t key) {
    int64_t m = key % R;
    if (m < 0) m += R;
    return m == rank;
  };
  // pass 1: how many keys are mine
  int cnt = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) cnt += mine_of(keys[i]) ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) warp_cnt[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 32; ++w) t += warp_cnt[w];
    total_mine = t;
    base_mine = 0;
    base_other = t;
    *n_mine_out = t;
  }
  __syncthreads();
  // pass 2: stable scatter, 1024 keys per round
  for (int i0 = 0; i0 < n; i0 += blockDim.x) {
    const int i = i0 + threadIdx.x;
    const bool valid = i < n;
    const int64_t key = valid ? keys[i] : 0;
    const bool mine = valid && mine_of(key);
    const unsigned bm = __ballot_sync(0xffffffffu, mine), bv = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) warp_cnt[warp] = __popc(bm) | (__popc(bv) << 16);
    __syncthreads();
    int before_mine = 0, before_valid = 0, all_mine = 0;
    for (int w = 0; w < 32; ++w) {
      const int c = warp_cnt[w];
      if (w < warp) {
        before_mine += c & 0xffff;
        before_valid += c >> 16;
      }
      all_mine += c & 0xffff;
    }
    const unsigned lt = (1u << lane) - 1u;
    const int my_mine = before_mine + __popc(bm & lt), my_valid = before_valid + __popc(bv & lt);
    if (valid) {
      const int dst = mine ? base_mine + my_mine : base_other + (my_valid - my_mine);
      keys_out[dst] = key;
      order[dst] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int round_valid = 0;
      for (int w = 0; w < 32; ++w) round_valid += warp_cnt[w] >> 16;
      base_mine += all_mine;
      base_other += round_valid - all_mine;
    }
    __syncthreads();
  }
}

}  // namespace ffc

using namespace ffc;

static int check_scatter_args(const void* qf, const void* rows, const void* cols, int B, int64_t Q, int D) {
  FFC_REQUIRE(qf && rows && cols, "queue scatter: NULL argument");
  FFC_REQUIRE(B >= 0 && Q >= 1 && D >= 4 && D % 4 == 0, "queue scatter: bad shape B=%d Q=%lld D=%d (D must be a multiple of 4)", B, (long long)Q, D);
  return FFC_OK;
}

extern "C" int ffc_queue_scatter(float* queue_f32_dev, void* queue_bf16_dev, const int32_t* rows_dev, const int32_t* cols_dev, const float* g_dev,
                                 int B, int64_t Q, int D, float* undo_f32_dev, void* stream) {
  int rc = check_scatter_args(queue_f32_dev, rows_dev, cols_dev, B, Q, D);
  if (rc) return rc;
  FFC_REQUIRE(g_dev != nullptr, "ffc_queue_scatter: g is NULL");
  if (B == 0) return FFC_OK;
  queue_scatter_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(queue_f32_dev, (__nv_bfloat16*)queue_bf16_dev, rows_dev, cols_dev, g_dev, B, Q, D,
                                                             undo_f32_dev, nullptr, nullptr, 0);
  FFC_LAUNCH_CHECK();
  return FFC_OK;
}

extern "C" int ffc_queue_scatter_indexed(float* queue_f32_dev, void* queue_bf16_dev, const int32_t* rows_dev, const int32_t* cols_dev,
                                         const float* g_dev, const int32_t* src_row_dev, int B, int64_t Q, int D, float* undo_f32_dev, void* stream) {
  return ffc_queue_scatter_overlay(queue_f32_dev, queue_bf16_dev, rows_dev, cols_dev, g_dev, src_row_dev, B, Q, D, undo_f32_dev, nullptr, 0, stream);
}

extern "C" int ffc_queue_scatter_overlay(float* queue_f32_dev, void* queue_bf16_dev, const int32_t* rows_dev, const int32_t* cols_dev,
                                         const float* g_dev, const int32_t* src_row_dev, int B, int64_t Q, int D, float* undo_f32_dev,
                                         int32_t* overlay_map_dev, int overlay_table, void* stream) {
  int rc = check_scatter_args(queue_f32_dev, rows_dev, cols_dev, B, Q, D);
  if (rc) return rc;
  FFC_REQUIRE(g_dev != nullptr && src_row_dev != nullptr, "ffc_queue_scatter_indexed: NULL argument");
  FFC_REQUIRE(!overlay_map_dev || overlay_table == 0 || (overlay_table == 1 && undo_f32_dev), "ffc_queue_scatter_overlay: table 1 marks need the undo buffer");
  if (B == 0) return FFC_OK;
  queue_scatter_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(queue_f32_dev, (__nv_bfloat16*)queue_bf16_dev, rows_dev, cols_dev, g_dev, B, Q, D,
                                                             undo_f32_dev, src_row_dev, overlay_map_dev, overlay_table);
  FFC_LAUNCH_CHECK();
  return FFC_OK;
}

namespace ffc {
__global__ void __launch_bounds__(256) overlay_clear_kernel(int32_t* __restrict__ map, const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, int n,
                                                            int64_t Q) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && cols[i] >= 0) map[(int64_t)rows[i] * Q + cols[i]] = -1;
}

// out[e] = sum over r (in rank order: deterministic) of slabs[r * stride + e]: the local half of the reduce-scatter that finalize's
// peer stores started
__global__ void __launch_bounds__(256) sum_slabs_kernel(const float4* __restrict__ slabs, int n_slabs, int64_t stride4, int64_t n4, float4* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 acc = __ldcs(slabs + i);
    for (int r = 1; r < n_slabs; ++r) {
      const float4 v = __ldcs(slabs + r * stride4 + i);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    out[i] = acc;
  }
}
// The same sum behind a barrier across the ranks, in ONE launch: the slabs are written by the PEERS' finalize kernels (stores over
// NVLink into this rank's peer-mapped staging buffer).  Block 0 tells every peer "my stores of step `epoch` are out" (the finalize
// kernels that issued them precede this launch in stream order; a system-scope fence, then a release store of the step number into
// the peer's flag word), every block waits until all peers have said so for this rank (acquire loads of its own flag words), then
// sums.  The flag words only ever grow, so a peer that is already a step ahead releases the wait as well.  A wait that lasts longer
// than `timeout_ns` raises *err_flag and goes on (wrong sums, no hang).
__global__ void __launch_bounds__(256) sum_slabs_barrier_kernel(const float4* __restrict__ slabs, int n_slabs, int64_t stride4, int64_t n4,
                                                                float4* __restrict__ out, int32_t* const* __restrict__ flag_ptrs, int my_rank,
                                                                int n_ranks, int32_t epoch, int32_t* err_flag, long long timeout_ns) {
  if (blockIdx.x == 0 && threadIdx.x < n_ranks) {
    __threadfence_system();
    int32_t* dst = flag_ptrs[threadIdx.x] + my_rank;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
  }
  if (threadIdx.x < n_ranks) {
    const int32_t* src = flag_ptrs[my_rank] + threadIdx.x;
    long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (true) {
      int32_t v;
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
      if (v - epoch >= 0) break;
      long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {
        *err_flag = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  // the slabs were written by other GPUs: read them at the L2 (the point of coherence for peer stores), never from a stale L1 line
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 acc = __ldcg(slabs + i);
    for (int r = 1; r < n_slabs; ++r) {
      const float4 v = __ldcg(slabs + r * stride4 + i);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    out[i] = acc;
  }
}
// The barrier alone (one block): "everything this rank stored into its peers' buffers before this launch is out" to every peer, then wait
// until every peer has said the same to this rank.  Same flag protocol as sum_slabs_barrier_kernel (own set of flag words).
__global__ void __launch_bounds__(64) peer_barrier_kernel(int32_t* const* __restrict__ flag_ptrs, int my_rank, int n_ranks, int32_t epoch, int32_t* err_flag,
                                                          long long timeout_ns) {
  if (threadIdx.x < n_ranks) {
    __threadfence_system();
    int32_t* dst = flag_ptrs[threadIdx.x] + my_rank;
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
    const int32_t* src = flag_ptrs[my_rank] + threadIdx.x;
    long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (true) {
      int32_t v;
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
      if (v - epoch >= 0) break;
      long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {
        *err_flag = 1;
        break;
      }
      __nanosleep(64);
    }
  }
}
}  // namespace ffc

extern "C" int ffc_sum_slabs_barrier(const float* slabs_dev, int n_slabs, int64_t slab_stride, int64_t n, float* out_dev, int32_t* const* flag_ptrs_dev,
                                     int my_rank, int n_ranks, int32_t epoch, int32_t* err_flag_dev, void* stream) {
  FFC_REQUIRE(slabs_dev && out_dev && flag_ptrs_dev && err_flag_dev && n_slabs >= 1 && n >= 0 && n % 4 == 0 && slab_stride % 4 == 0,
              "ffc_sum_slabs_barrier: bad arguments (n and the stride must be multiples of 4)");
  FFC_REQUIRE(n_ranks >= 1 && n_ranks <= 64 && my_rank >= 0 && my_rank < n_ranks, "ffc_sum_slabs_barrier: rank %d of %d", my_rank, n_ranks);
  // at most one resident wave of blocks: a block that is not resident cannot keep a peer waiting, but there is no point in more
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n / 4, 256), 148 * 4));
  sum_slabs_barrier_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)slabs_dev, n_slabs, slab_stride / 4, n / 4, (float4*)out_dev, flag_ptrs_dev,
                                                                      my_rank, n_ranks, epoch, err_flag_dev, 5000000000LL);
  FFC_LAUNCH_CHECK();
  return FFC_OK;
}

extern "C" int ffc_peer_barrier(int32_t* const* flag_ptrs_dev, int my_rank, int n_ranks, int32_t epoch, int32_t* err_flag_dev, void* stream) {
  FFC_REQUIRE(flag_ptrs_dev && err_flag_dev && n_ranks >= 1 && n_ranks <= 64 && my_rank >= 0 && my_rank < n_ranks, "ffc_peer_barrier: rank %d of %d", my_rank, n_ranks);
  peer_barrier_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(flag_ptrs_dev, my_rank, n_ranks, epoch, err_flag_dev, 5000000000LL);
  FFC_LAUNCH_CHECK();
  return FFC_OK;
}

extern "C" int ffc_overlay_clear(int32_t* overlay_map_dev, const int32_t* rows_dev, const int32_t* cols_dev, int n, int64_t Q, void* stream) {
  FFC_REQUIRE(overlay_map_dev && rows_dev && cols_dev && n >= 0 && Q >= 1, "ffc_overlay_clear: bad arguments");
  if (n == 0) return FFC_OK;
  overlay_clear_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(overlay_map_dev, rows_dev, cols_dev, n, Q);
  FFC_LAUNCH_CHECK();
  return FFC_OK;
}

extern "C" int ffc_sum_slabs(const float* slabs_dev, int n_slabs, int64_t slab_stride, int64_t n, float* out_dev, void* stream) {
  FFC_REQUIRE(slabs_dev && out_dev && n_slabs >= 1 && n >= 0 && n % 4 == 0 && slab_stride % 4 == 0, "ffc_sum_slabs: bad arguments (n and the stride must be multiples of 4)");
  if (n == 0) return FFC_OK;
  const int blocks = (int)std::min<int64_t>(ceil_div64(n / 4, 256), 148 * 8);
  sum_slabs_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)slabs_dev, n_slabs, slab_stride / 4, n / 4, (float4*)out_dev);
  FFC_LAUNCH_CHECK();
  return FFC_OK;
}

extern "C" int ffc_route_keys(const int64_t* keys_dev, int n, int n_ranks, int rank, int64_t* keys_out_dev, int32_t* order_out_dev,
                              int32_t* n_mine_out_dev, void* stream) {
  FFC_REQUIRE(keys_dev && keys_out_dev && order_out_dev && n_mine_out_dev && n >= 0 && n_ranks >= 1 && rank >= 0 && rank < n_ranks,
              "ffc_route_keys: bad arguments");
  if (n == 0) {
    FFC_CUDA(cudaMemsetAsync(n_mine_out_dev, 0, sizeof(int32_t), (cudaStream_t)stream));
    return FFC_OK;
  }
  route_keys_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(keys_dev, n, n_ranks, rank, keys_out_dev, order_out_dev, n_mine_out_dev);
  FFC_LAUNCH_CHECK();
  return FFC_OK;
}

extern "C" int ffc_queue_restore(float* queue_f32_dev, void* queue_bf16_dev, const int32_t* rows_dev, const int32_t* cols_dev,
                                 const float* undo_f32_dev, int B, int64_t Q, int D, void* stream) {
  int rc = check_scatter_args(queue_f32_dev, rows_dev, cols_dev, B, Q, D);
  if (rc) return rc;
  FFC_REQUIRE(undo_f32_dev != nullptr, "ffc_queue_restore: undo buffer is NULL");
  if (B == 0) return FFC_OK;
  queue_restore_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(queue_f32_dev, (__nv_bfloat16*)queue_bf16_dev, rows_dev, cols_dev, undo_f32_dev, B, Q, D, 0);
  FFC_LAUNCH_CHECK();
  return FFC_OK;
}

extern "C" int ffc_cast_bf16(const float* src_dev, void* dst_bf16_dev, int64_t n, void* stream) {
  FFC_REQUIRE(src_dev && dst_bf16_dev && n >= 0 && n % 4 == 0, "ffc_cast_bf16: bad arguments (n must be a multiple of 4)");
  if (n == 0) return FFC_OK;
  const int64_t n4 = n / 4;
  const int blocks = (int)std::min<int64_t>(ceil_div64(n4, 256), 148 * 16);
  cast_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)src_dev, (uint2*)dst_bf16_dev, n4);
  FFC_LAUNCH_CHECK();
  return FFC_OK;
}

extern "C" int ffc_queue_restore_packed(float* queue_f32_dev, void* queue_bf16_dev, const int32_t* rows_dev, const int32_t* cols_dev,
                                        const float* undo_f32_dev, int B, int64_t Q, int D, void* stream) {
  int rc = check_scatter_args(queue_f32_dev, rows_dev, cols_dev, B, Q, D);
  if (rc) return rc;
  FFC_REQUIRE(undo_f32_dev != nullptr, "ffc_queue_restore_packed: undo buffer is NULL");
  if (B == 0) return FFC_OK;
  queue_restore_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(queue_f32_dev, (__nv_bfloat16*)queue_bf16_dev, rows_dev, cols_dev, undo_f32_dev, B, Q, D, 1);
  FFC_LAUNCH_CHECK();
  return FFC_OK;
}
