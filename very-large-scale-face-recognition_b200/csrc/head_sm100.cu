// tcgen05 / TMEM / TMA sweep of the FFC head for sm_100a (the hot kernel).
//
// One work item = (128 probe rows) x (a contiguous range of 128-row queue tiles), processed by a
// CLUSTER OF TWO CTAs with different roles, because at D = 512 the fp32 gradient accumulator
// O[128 x 512] alone fills the 512 TMEM columns of one SM:
//
//   CTA rank 0, "S-CTA":  S = P . W_tile^T        (tcgen05.mma, M=128 N=128 K=D, A = P resident in
//                          smem, B = W K-chunks streamed by TMA);  8 epilogue warps read S from TMEM,
//                          apply exclusions / SV transform, p~ = 2^(a*cos - b), accumulate the softmax
//                          denominator and the running top-k, and store p~ as bf16 straight into the
//                          peer CTA's shared memory (st.shared::cluster) in the K-major SWIZZLE_128B
//                          layout the second GEMM reads.
//   CTA rank 1, "O-CTA":  O += P~ . W_tile          (tcgen05.mma, M=128 N<=256 per instruction, K=128
//                          queue rows per tile, A = P~ from smem, B = the same W tile streamed by TMA
//                          and read MN-major);  O stays in TMEM for the whole item and is written once.
//
// Executed tensor FLOPs == algorithmic FLOPs (4 * rows * cols * D): no recompute, no B x Q logits.
// Synchronisation is all mbarriers: TMA -> MMA (full/empty rings), MMA -> epilogue (tcgen05.commit),
// epilogue -> peer MMA (st.async complete_tx on the peer's mbarrier), peer MMA -> epilogue (multicast tcgen05.commit).
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <vector>

#include "head_internal.cuh"

namespace ffc {

constexpr int BM = 128;            // probe rows per item
constexpr int BN = 128;            // queue rows per tile
constexpr int KC = 64;             // bf16 elements per 128-byte swizzle row
#ifndef FFC_NS1
#define FFC_NS1 6
#endif
#ifndef FFC_JB
#define FFC_JB 32
#endif
constexpr int NS1 = FFC_NS1;       // S-CTA W K-chunk stages (16 KB each)
constexpr int NSB = 4;             // S accumulators in the S-CTA's TMEM (4 x 128 columns)
constexpr int NPB = 3;             // P~ buffers in the O-CTA's shared memory
constexpr int JB = FFC_JB;         // queue rows per O-CTA W stage
constexpr int NEPI = 3;             // epilogue warpgroups in the S-CTA (tiles are dealt round-robin)
constexpr int NTHREADS = 128 + NEPI * 128;   // 4 control warps + 12 epilogue warps
constexpr int CHUNK1_BYTES = BN * KC * 2;   // 16384
constexpr float LOG2E = 1.4426950408889634f;

// ---- shared memory map (same for both roles; 1024-byte aligned base) ----
// [0, 1024)                 barriers, tmem base, small staging
// S-CTA: [1024, +D/64*16K)  P tile (K-major SW128 chunks)      then NS1 x 16 KB W chunk ring
// O-CTA: [1024, +96K)       P~ buffers 3 x 32 KB               then NS2 x (D/64 * 4 KB) W stage ring
constexpr int OFF_DATA = 1024;
constexpr int PT_BYTES = BM * BN * 2;       // 32768 per P~ buffer

struct Bars {   // all in the first 1024 bytes
  uint64_t p_full;
  uint64_t w_full[NS1], w_empty[NS1];
  uint64_t s_full[NSB], s_empty[NSB];
  uint64_t pt_empty[NPB];          // S-CTA side: P~ buffer b may be overwritten (signalled by the peer's tcgen05.commit)
  uint64_t w2_full[8], w2_empty[8];
  uint64_t pt_full[NPB];           // O-CTA side: P~ buffer b is complete (st.async complete_tx, 32 KB per tile)
  uint64_t o_full;
  uint32_t tmem_base;
  uint32_t pad;
};
static_assert(sizeof(Bars) <= 512, "barrier block too large");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // pairs with a remote release.cluster arrive
  const uint32_t addr = smem_u32(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAITC_LOOP:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAITC_DONE;\n"
      "bra WAITC_LOOP;\n"
      "WAITC_DONE:\n"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// asynchronous 16-byte store into the peer CTA's shared memory; completion is counted (complete_tx) on the
// peer's mbarrier, so the writer needs no fence / drain before the consumer may be released
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint32_t mbar, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr), "r"(a), "r"(b),
               "r"(c), "r"(d), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// tcgen05.commit that arrives on the barrier at the same offset in the CTAs of `cta_mask`
__device__ __forceinline__ void tc_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {   // long waits: back off instead of spinning
  const uint32_t addr = smem_u32(bar);
  uint32_t done = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    asm volatile("nanosleep.u32 2000;");
  }
}

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): SWIZZLE_128B, version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
  d |= (uint64_t)2 << 61;   // layout_type = SWIZZLE_128B
  return d;
}
// instruction descriptor (InstrDescriptor): bf16 x bf16 -> f32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

struct Sm100Params {
  int n_rows, D;
  int64_t n_cols;
  const int32_t* n_cols_dev;
  const int32_t* tcol;
  const uint32_t* cmask;
  const float* thr;
  const uint8_t* is_out;
  float a2, b2;       // p~ = 2^(a2 * z - b2)
  int sv, k;
  int n_chunks, tiles_per_chunk, ns2;
  int debug;   // bottleneck isolation, results are WRONG except 0 and 6 (FFC_SM100_DEBUG, see profiles/r1_bottleneck_isolation.md):
               // 1 = O-CTA skips TMA+MMA, 2 = epilogue skips tcgen05.ld/exp, 3 = S-CTA skips TMA+MMA, 6 = plain remote stores
  float* l_part;
  float* o_part;
  float* topv_part;
  int32_t* topi_part;
};

template <bool SV>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
    ffc_head_sweep_sm100_kernel(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_w1,
                                const __grid_constant__ CUtensorMap map_w2, const Sm100Params prm) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // SWIZZLE_128B operands need a 1024-byte aligned base; both CTAs of the pair compute the same offset
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars& bars = *reinterpret_cast<Bars*>(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int item = blockIdx.x >> 1;
  const int n_row_tiles = (prm.n_rows + BM - 1) / BM;
  const int rt = item % n_row_tiles, chunk = item / n_row_tiles;
  const int row0 = rt * BM;
  const int D = prm.D, nkc = D / KC;
  const int64_t n_cols = prm.n_cols_dev ? (int64_t)*prm.n_cols_dev : prm.n_cols;
  const int n_tiles_total = (int)((n_cols + BN - 1) / BN);
  const int t_begin = chunk * prm.tiles_per_chunk;
  int t_end = t_begin + prm.tiles_per_chunk;
  if (t_end > n_tiles_total) t_end = n_tiles_total;
  const int n_tiles = t_end > t_begin ? t_end - t_begin : 0;
  const int ns2 = prm.ns2;
  const int stage2_bytes = nkc * JB * KC * 2;   // D/64 boxes of 4 KB
  const int n_nhalf = (D + 255) / 256;          // GEMM-2 instructions per K step (N <= 256 each)
  const int n2 = D < 256 ? D : 256;
  // P~ hand-off flavour: st.async (per-store complete_tx on the peer's mbarrier) or plain remote stores + one
  // fence/arrive per warp and tile (FFC_SM100_DEBUG=6)
  const bool plain_st = prm.debug == 6;

  unsigned char* sP = smem + OFF_DATA;                         // S-CTA
  unsigned char* sW1 = sP + nkc * CHUNK1_BYTES;                // S-CTA
  unsigned char* sPt = smem + OFF_DATA;                        // O-CTA
  unsigned char* sW2 = sPt + NPB * PT_BYTES;                   // O-CTA

  if (threadIdx.x == 0) {
    mbar_init(&bars.p_full, 1);
    for (int i = 0; i < NS1; ++i) {
      mbar_init(&bars.w_full[i], 1);
      mbar_init(&bars.w_empty[i], 1);
    }
    for (int i = 0; i < NSB; ++i) {
      mbar_init(&bars.s_full[i], 1);
      mbar_init(&bars.s_empty[i], 4);
    }
    for (int i = 0; i < NPB; ++i) {
      mbar_init(&bars.pt_empty[i], 1);
      mbar_init(&bars.pt_full[i], plain_st ? 4 : 1);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&bars.w2_full[i], 1);
      mbar_init(&bars.w2_empty[i], 1);
    }
    mbar_init(&bars.o_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    // S-CTA: NSB x 128 columns of S; O-CTA: D columns of O (power of two >= 32)
    const uint32_t ncols = rank == 0 ? (uint32_t)(NSB * BN) : (uint32_t)(D < 32 ? 32 : D);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&bars.tmem_base);
  if (rank == 0) {
    // =========================================== S-CTA ===========================================
    if (warp == 0) {
      if (lane == 0 && n_tiles > 0 && prm.debug != 3) {
        mbar_expect_tx(&bars.p_full, (uint32_t)(nkc * CHUNK1_BYTES));
        for (int kc = 0; kc < nkc; ++kc) tma_load_2d(&map_p, &bars.p_full, sP + kc * CHUNK1_BYTES, kc * KC, row0);
        int stage = 0;
        uint32_t ph = 0;
        for (int t = t_begin; t < t_end; ++t) {
          for (int kc = 0; kc < nkc; ++kc) {
            mbar_wait(&bars.w_empty[stage], ph ^ 1);
            mbar_expect_tx(&bars.w_full[stage], CHUNK1_BYTES);
            tma_load_2d(&map_w1, &bars.w_full[stage], sW1 + stage * CHUNK1_BYTES, kc * KC, t * BN);
            if (++stage == NS1) {
              stage = 0;
              ph ^= 1;
            }
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0 && n_tiles > 0) {
        constexpr uint32_t idesc = make_idesc(BM, BN, 0, 0);
        if (prm.debug != 3) mbar_wait(&bars.p_full, 0);
        tc_fence_after();
        int stage = 0;
        uint32_t ph = 0;
        for (int i = 0; i < n_tiles; ++i) {
          const int sb = i & (NSB - 1);
          const uint32_t use = (uint32_t)(i / NSB);
          mbar_wait(&bars.s_empty[sb], (use & 1) ^ 1);
          tc_fence_after();
          if (prm.debug == 3) {
            mbar_arrive(&bars.s_full[sb]);
            continue;
          }
          const uint32_t tmem_s = tmem_base + (uint32_t)(sb * BN);
          for (int kc = 0; kc < nkc; ++kc) {
            mbar_wait(&bars.w_full[stage], ph);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(sP + kc * CHUNK1_BYTES);
            const uint32_t b_addr = smem_u32(sW1 + stage * CHUNK1_BYTES);
#pragma unroll
            for (int k = 0; k < KC / 16; ++k) {
              tc_mma(tmem_s, make_desc(a_addr + k * 32, 16, 1024), make_desc(b_addr + k * 32, 16, 1024), idesc, (kc | k) ? 1u : 0u);
            }
            tc_commit(&bars.w_empty[stage]);
            if (++stage == NS1) {
              stage = 0;
              ph ^= 1;
            }
          }
          tc_commit(&bars.s_full[sb]);
        }
      }
    } else if (warp >= 4) {
      // ---- epilogue: warpgroup g takes the tiles of parity g (S buffer i%4, P~ buffer i%3) ----
      const int g = (warp - 4) >> 2;
      const int q4 = warp & 3;                    // TMEM lane quarter
      const int r_local = q4 * 32 + lane;         // row within the item
      const int row = row0 + r_local;
      const bool row_ok = row < prm.n_rows;
      const int32_t tcol = row_ok ? prm.tcol[row] : -1;
      const bool outl = row_ok && prm.is_out[row];
      const bool warp_out = __any_sync(0xffffffffu, outl);
      float thr = INFINITY;
      if (SV && row_ok && prm.thr) thr = prm.thr[row];
      const float a2 = prm.a2, b2 = prm.b2;
      const int k = prm.k;
      float lsum = 0.f;
      // Hard-negative top-k (ffc.py:86-92) as a branch-free register list of integer keys: only POSITIVE cosines can
      // contribute (clip(.., 0) zeroes the rest and their gradient), positive floats order like their int32 bit
      // patterns, and the low 4 mantissa bits carry the column's index inside its 16-column chunk (value error 2^-19).
      // tk[] descending keys (0 = empty), tc[] first column of the chunk the key came from.
      int tk[KMAX], tc[KMAX];
#pragma unroll
      for (int q = 0; q < KMAX; ++q) {
        tk[q] = 0;
        tc[q] = -1;
      }
      int kth = 0;
      const uint32_t pt_remote0 = map_to_rank(smem_u32(smem + OFF_DATA), 1);
      const uint32_t ptfull_remote0 = map_to_rank(smem_u32(&bars.pt_full[0]), 1);
      const uint32_t sw = (uint32_t)(r_local & 7);
      // tile i uses S buffer i % NSB and P~ buffer i % NPB; NEPI == NPB, so this warpgroup always writes P~ buffer g
      static_assert(NEPI == NPB, "epilogue warpgroups and P~ buffers are paired");
      const int pb = g;
      uint32_t pt_use = 0;
      for (int i = g; i < n_tiles; i += NEPI, ++pt_use) {
        const int sb = i & (NSB - 1);
        const int j0 = (t_begin + i) * BN;
        // exclusion words of the four 32-column chunks of this tile (one 16-byte load, issued before the waits)
        uint4 cm = make_uint4(0u, 0u, 0u, 0u);
        if (prm.cmask && (int64_t)j0 < n_cols) cm = __ldg(reinterpret_cast<const uint4*>(prm.cmask + (j0 >> 5)));
        mbar_wait(&bars.s_full[sb], (uint32_t)(i / NSB) & 1);
        tc_fence_after();
        mbar_wait(&bars.pt_empty[pb], (pt_use & 1) ^ 1);
        const uint32_t tmem_s = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(sb * BN);
        const uint32_t pt_remote = pt_remote0 + (uint32_t)(pb * PT_BYTES);
        const uint32_t ptfull_remote = ptfull_remote0 + (uint32_t)(pb * 8);
        // A tile is "clean" for this warp when no column is excluded (no `ones` column, no row's target, not the
        // ragged tail) and no row takes part in the top-k: then the chunk loop is pure ld -> ex2 -> pack -> st.async.
        const bool clean = !warp_out && !__any_sync(0xffffffffu, (cm.x | cm.y | cm.z | cm.w) != 0u || (unsigned)(tcol - j0) < (unsigned)BN) &&
                           (int64_t)j0 + BN <= n_cols;
        if (clean) {
          float l0 = 0.f, l1 = 0.f;
#pragma unroll 1
          for (int cc = 0; cc < BN / 32; ++cc) {
            uint32_t v[32];
            if (prm.debug == 2) {
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = 0u;
            } else {
              tc_ld32(tmem_s + cc * 32, v);
            }
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              if (prm.debug == 2) {
                pk[c >> 1] = 0u;
                continue;
              }
              const float x0 = __uint_as_float(v[c]), x1 = __uint_as_float(v[c + 1]);
              float p0, p1, g0, g1;
              if (SV) {
                const bool m0 = x0 > thr, m1 = x1 > thr;
                p0 = ex2f(fmaf(m0 ? fmaf(SV_T, x0, SV_T - 1.f) : x0, a2, -b2));
                p1 = ex2f(fmaf(m1 ? fmaf(SV_T, x1, SV_T - 1.f) : x1, a2, -b2));
                g0 = m0 ? p0 * SV_T : p0;
                g1 = m1 ? p1 * SV_T : p1;
              } else {
                p0 = g0 = ex2f(fmaf(x0, a2, -b2));
                p1 = g1 = ex2f(fmaf(x1, a2, -b2));
              }
              l0 += p0;
              l1 += p1;
              pk[c >> 1] = pack_bf16(g0, g1);
            }
            const uint32_t base = pt_remote + (uint32_t)((cc >> 1) * (BM * 128)) + (uint32_t)(r_local * 128);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t chunk16 = (uint32_t)((cc & 1) * 4 + q);
              if (plain_st)
                st_cluster_v4(base + ((chunk16 ^ sw) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
              else
                st_async_v4(base + ((chunk16 ^ sw) << 4), ptfull_remote, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            }
          }
          lsum += l0 + l1;
        } else
#pragma unroll 1
        for (int cc = 0; cc < BN / 16; ++cc) {
          uint32_t v[16];
          tc_ld16(tmem_s + cc * 16, v);
          const int col0 = j0 + cc * 16;
          const uint32_t cw = (cc >> 1) == 0 ? cm.x : (cc >> 1) == 1 ? cm.y : (cc >> 1) == 2 ? cm.z : cm.w;
          uint32_t excl = (cw >> ((cc & 1) * 16)) & 0xffffu;
          const int trel = tcol - col0;
          if ((unsigned)trel < 16u) excl |= 1u << trel;
          if ((int64_t)col0 + 16 > n_cols) {
            const int nv = (int)(n_cols - col0);   // valid columns in this chunk (may be <= 0)
            excl |= nv <= 0 ? 0xffffu : (0xffffu << nv) & 0xffffu;
          }
          const bool slow = __any_sync(0xffffffffu, excl != 0u);
          // hard-negative top-k on the raw cosines of outlier rows: all lanes run the same code; a lane with a
          // candidate (key above its k-th) extracts its largest remaining key per round until no lane has any left
          if (warp_out) {
            int key[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              key[c] = (int)((v[c] & 0xfffffff0u) | (uint32_t)c);
              if (slow && ((excl >> c) & 1u)) key[c] = 0;
            }
            int bound = 0x7fffffff;      // keys >= bound were already extracted in this chunk
            while (true) {
              int mx = 0;
#pragma unroll
              for (int c = 0; c < 16; ++c) mx = max(mx, key[c] < bound ? key[c] : 0);
              const bool has = outl && mx > kth;
              if (!__any_sync(0xffffffffu, has)) break;
              if (has) {
                int xk = mx, xc = col0;
#pragma unroll
                for (int r = 0; r < KMAX; ++r) {
                  if (r < k) {
                    const bool pgt = xk > tk[r];
                    const int nk = pgt ? xk : tk[r], nc = pgt ? xc : tc[r];
                    xk = pgt ? tk[r] : xk;
                    xc = pgt ? tc[r] : xc;
                    tk[r] = nk;
                    tc[r] = nc;
                  }
                }
#pragma unroll
                for (int r = 0; r < KMAX; ++r)
                  if (r == k - 1) kth = tk[r];
                bound = mx;
              } else {
                bound = 0;               // this lane is done with the chunk
              }
            }
          }
          uint32_t pk[8];
#pragma unroll
          for (int c = 0; c < 16; c += 2) {
            float x0 = __uint_as_float(v[c]), x1 = __uint_as_float(v[c + 1]);
            float p0, p1, g0, g1;
            if (SV) {
              const bool m0 = x0 > thr, m1 = x1 > thr;
              p0 = ex2f(fmaf(m0 ? fmaf(SV_T, x0, SV_T - 1.f) : x0, a2, -b2));
              p1 = ex2f(fmaf(m1 ? fmaf(SV_T, x1, SV_T - 1.f) : x1, a2, -b2));
              g0 = m0 ? p0 * SV_T : p0;
              g1 = m1 ? p1 * SV_T : p1;
            } else {
              p0 = g0 = ex2f(fmaf(x0, a2, -b2));
              p1 = g1 = ex2f(fmaf(x1, a2, -b2));
            }
            if (slow) {   // warp-uniform: only chunks that contain an excluded column pay for the per-element test
              if ((excl >> c) & 1u) p0 = g0 = 0.f;
              if ((excl >> (c + 1)) & 1u) p1 = g1 = 0.f;
            }
            lsum += p0 + p1;
            pk[c >> 1] = pack_bf16(g0, g1);
          }
          // P~[r_local][cc*16 .. +16) as 2 x 16-byte chunks, K-major SWIZZLE_128B: 64-column sub-tiles of 16 KB.
          // st.async: each 16-byte store counts itself on the peer's pt_full[pb] (32 KB per tile in total).
          const uint32_t base = pt_remote + (uint32_t)((cc >> 2) * (BM * 128)) + (uint32_t)(r_local * 128);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const uint32_t chunk16 = (uint32_t)((cc & 3) * 2 + q);
            if (plain_st)
              st_cluster_v4(base + ((chunk16 ^ sw) << 4), pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            else
              st_async_v4(base + ((chunk16 ^ sw) << 4), ptfull_remote, pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        }
        // S buffer sb may be overwritten by a later tile's MMA
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.s_empty[sb]);
        if (plain_st) {   // publish this warp's rows of P~ (generic-proxy remote writes -> async-proxy reads in the peer)
          asm volatile("fence.proxy.async;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_remote(ptfull_remote);
        }
      }
      // ---- per-row partials: combine the two warpgroups through shared memory ----
      float* stage_v = reinterpret_cast<float*>(sW1);                             // [2][128][KMAX] (W ring is idle now)
      int32_t* stage_i = reinterpret_cast<int32_t*>(sW1 + 2 * BM * KMAX * 4);
      float* stage_l = reinterpret_cast<float*>(sW1 + 4 * BM * KMAX * 4);         // [2][128]
      // every MMA that read the W ring has completed once the last s_full was observed by its epilogue warpgroup;
      // synchronise the 12 epilogue warps (named barrier 1) before reusing the ring as staging
      asm volatile("bar.sync 1, 384;" ::: "memory");
      float tv[KMAX];
      int32_t ti[KMAX];
#pragma unroll
      for (int q = 0; q < KMAX; ++q) {
        const bool live = tk[q] > 0;
        tv[q] = live ? __int_as_float(tk[q] & (int)0xfffffff0) : -INFINITY;
        ti[q] = live ? tc[q] + (tk[q] & 15) : -1;
      }
      if (g >= 1) {
        stage_l[(g - 1) * BM + r_local] = lsum;
#pragma unroll
        for (int q = 0; q < KMAX; ++q) {
          stage_v[((g - 1) * BM + r_local) * KMAX + q] = tv[q];
          stage_i[((g - 1) * BM + r_local) * KMAX + q] = ti[q];
        }
      }
      asm volatile("bar.sync 1, 384;" ::: "memory");
      if (g == 0 && row_ok) {
        float kthv = -INFINITY;
#pragma unroll
        for (int q = 0; q < KMAX; ++q)
          if (q == k - 1) kthv = tv[q];
#pragma unroll 1
        for (int og = 0; og < NEPI - 1; ++og) {
          lsum += stage_l[og * BM + r_local];
          if (outl) {
#pragma unroll 1
            for (int q = 0; q < k; ++q) {
              const float x = stage_v[(og * BM + r_local) * KMAX + q];
              if (x > kthv) {
                topk_insert<KMAX>(tv, ti, k, x, stage_i[(og * BM + r_local) * KMAX + q]);
#pragma unroll
                for (int qq = 0; qq < KMAX; ++qq)
                  if (qq == k - 1) kthv = tv[qq];
              }
            }
          }
        }
        const int64_t pr = (int64_t)chunk * prm.n_rows + row;
        prm.l_part[pr] = lsum;
#pragma unroll
        for (int q = 0; q < KMAX; ++q) {
          if (q < k) {
            prm.topv_part[pr * k + q] = tv[q];
            prm.topi_part[pr * k + q] = ti[q];
          }
        }
      }
    }
  } else {
    // =========================================== O-CTA ===========================================
    if (warp == 0) {
      if (lane == 0 && n_tiles > 0 && prm.debug != 1) {
        int stage = 0;
        uint32_t ph = 0;
        for (int t = t_begin; t < t_end; ++t) {
          for (int jb = 0; jb < BN / JB; ++jb) {
            mbar_wait(&bars.w2_empty[stage], ph ^ 1);
            mbar_expect_tx(&bars.w2_full[stage], (uint32_t)stage2_bytes);
            unsigned char* dst = sW2 + stage * stage2_bytes;
            for (int kc = 0; kc < nkc; ++kc) tma_load_2d(&map_w2, &bars.w2_full[stage], dst + kc * (JB * 128), kc * KC, t * BN + jb * JB);
            if (++stage == ns2) {
              stage = 0;
              ph ^= 1;
            }
          }
        }
      }
    } else if (warp == 1) {
      if (lane == 0 && n_tiles > 0) {
        const uint32_t idesc = make_idesc(BM, n2, 0, 1);
        int stage = 0;
        uint32_t ph = 0;
        for (int i = 0; i < n_tiles; ++i) {
          const int pb = i % NPB;
          const uint32_t use = (uint32_t)(i / NPB);
          // this thread is the single arriver of pt_full[pb]; the 32 KB of P~ arrive as st.async complete_tx bytes
          if (!plain_st) mbar_expect_tx(&bars.pt_full[pb], PT_BYTES);
          mbar_wait_cluster(&bars.pt_full[pb], use & 1);
          asm volatile("fence.proxy.async;" ::: "memory");
          tc_fence_after();
          if (prm.debug == 1) {
            mbar_arrive_remote(map_to_rank(smem_u32(&bars.pt_empty[pb]), 0));
            continue;
          }
          const uint32_t a_base = smem_u32(sPt + pb * PT_BYTES);
          for (int jb = 0; jb < BN / JB; ++jb) {
            mbar_wait(&bars.w2_full[stage], ph);
            tc_fence_after();
            const uint32_t b_base = smem_u32(sW2 + stage * stage2_bytes);
#pragma unroll
            for (int k16 = 0; k16 < JB / 16; ++k16) {
              const int kk = jb * JB + k16 * 16;                                    // queue row within the tile
              const uint64_t adesc = make_desc(a_base + (uint32_t)((kk >> 6) * (BM * 128)) + (uint32_t)(((kk & 63) >> 4) * 32), 16, 1024);
              for (int nh = 0; nh < n_nhalf; ++nh) {
                // B: MN-major (N = feature dim), 64-wide atoms JB*128 bytes apart (LBO), 8-row groups 1024 bytes apart (SBO)
                const uint64_t bdesc = make_desc(b_base + (uint32_t)(nh * 4 * (JB * 128)) + (uint32_t)(k16 * 16 * 128), JB * 128, 1024);
                tc_mma(tmem_base + (uint32_t)(nh * 256), adesc, bdesc, idesc, (i | jb | k16) ? 1u : 0u);
              }
            }
            tc_commit(&bars.w2_empty[stage]);
            if (++stage == ns2) {
              stage = 0;
              ph ^= 1;
            }
          }
          tc_commit_mcast(&bars.pt_empty[pb], (uint16_t)1);   // frees P~ buffer pb: arrives in the S-CTA (cluster rank 0)
        }
        if (prm.debug == 1) mbar_arrive(&bars.o_full); else tc_commit(&bars.o_full);
        mbar_wait(&bars.o_full, 0);     // one polling thread; the 8 epilogue warps block on a hardware barrier instead
      }
      __syncwarp();
      asm volatile("bar.sync 2, 288;" ::: "memory");
    } else if (warp >= 4 && warp < 12) {
      // ---- O epilogue: TMEM -> global partial [chunk][row][D]; warpgroup g takes half of the columns ----
      const int g = (warp - 4) >> 2;
      const int q4 = warp & 3;
      const int r_local = q4 * 32 + lane;
      const int row = row0 + r_local;
      const bool row_ok = row < prm.n_rows;
      asm volatile("bar.sync 2, 288;" ::: "memory");   // released by the MMA warp once every tcgen05.mma of the item has completed
      tc_fence_after();
      const int half = D / 2;                      // D is a multiple of 64
      float* dst = prm.o_part + ((int64_t)chunk * prm.n_rows + (row_ok ? row : 0)) * D + g * half;
      for (int c0 = 0; c0 < half; c0 += 32) {
        uint32_t v[32];
        if (n_tiles > 0) {
          tc_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(g * half + c0), v);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) v[c] = 0u;
        }
        if (row_ok) {
#pragma unroll
          for (int c = 0; c < 32; c += 4) {
            *reinterpret_cast<uint4*>(dst + c0 + c) = make_uint4(v[c], v[c + 1], v[c + 2], v[c + 3]);
          }
        }
      }
      tc_fence_before();
    }
  }
  // S-CTA with no tiles still owes l / top-k partials (zeros / -inf): handled here for uniformity
  if (rank == 0 && n_tiles == 0 && warp >= 4 && warp < 8) {
    const int r_local = (warp & 3) * 32 + lane;
    const int row = row0 + r_local;
    if (row < prm.n_rows) {
      const int64_t pr = (int64_t)chunk * prm.n_rows + row;
      prm.l_part[pr] = 0.f;
      for (int q = 0; q < prm.k; ++q) {
        prm.topv_part[pr * prm.k + q] = -INFINITY;
        prm.topi_part[pr * prm.k + q] = -1;
      }
    }
  }
  __syncthreads();
  cluster_sync_all();     // the peer may still signal into this CTA's shared memory until here
  if (warp == 2) {
    const uint32_t ncols = rank == 0 ? (uint32_t)(NSB * BN) : (uint32_t)(D < 32 ? 32 : D);
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct MapKey {
  const void* ptr;
  int64_t rows;
  int D, box_rows;
  bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && D == o.D && box_rows == o.box_rows; }
};
struct Sm100Cache {
  PFN_encodeTiled encode = nullptr;
  std::vector<std::pair<MapKey, CUtensorMap>> maps;
  bool attr_set = false;
};

Sm100Cache* sm100_cache_create() { return new Sm100Cache(); }
void sm100_cache_destroy(Sm100Cache* c) { delete c; }

static int get_map(Sm100Cache* c, const void* ptr, int64_t rows, int D, int box_rows, CUtensorMap* out) {
  MapKey key{ptr, rows, D, box_rows};
  for (auto& e : c->maps)
    if (e.first == key) {
      *out = e.second;
      return FFC_OK;
    }
  if (!c->encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    FFC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    FFC_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
    c->encode = (PFN_encodeTiled)fn;
  }
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)D * 2};
  cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = c->encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with %d (ptr=%p rows=%lld D=%d box_rows=%d)", (int)r, ptr, (long long)rows, D, box_rows);
    return FFC_ERR_CUDA;
  }
  if (c->maps.size() > 64) c->maps.clear();
  c->maps.emplace_back(key, m);
  *out = m;
  return FFC_OK;
}

int sm100_pick_chunks(int n_rows, int64_t n_cols, int D) {
  (void)D;
  // items = row_tiles x chunks run as CTA pairs, 74 pairs at a time: pick the chunk count whose last wave is fullest
  // (ties -> fewer chunks: fewer partials and fewer P loads / O write-outs)
  const int row_tiles = std::max(1, (n_rows + BM - 1) / BM);
  const int64_t n_tiles = std::max<int64_t>(1, ceil_div64(n_cols, BN));
  const int max_c = (int)std::min<int64_t>(n_tiles, 40);
  int best = 1;
  double best_eff = 0.0;
  for (int c = 1; c <= max_c; ++c) {
    const int items = row_tiles * c;
    const int waves = (items + 73) / 74;
    // every item pays a fixed cost (prologue, P load, O write-out) worth about 12 tile-times
    const double tiles_per_item = (double)n_tiles / c;
    const double eff = ((double)items / (waves * 74.0)) * (tiles_per_item / (tiles_per_item + 12.0));
    if (eff > best_eff * 1.005) {
      best_eff = eff;
      best = c;
    }
  }
  return best;
}

static int ns2_for(int D) {
  const int stage = (D / KC) * JB * KC * 2;
  int ns = (int)((225 * 1024 - NPB * PT_BYTES) / stage);
  return ns > 8 ? 8 : (ns < 2 ? 2 : ns);
}

int launch_sweep_sm100(Sm100Cache* cache, int cache_slot, const SweepArgs& a, cudaStream_t s) {
  (void)cache_slot;
  FFC_REQUIRE(a.D % 64 == 0 && a.D >= 64 && a.D <= 512 && (a.D & (a.D - 1)) == 0, "tcgen05 sweep: D=%d must be 64, 128, 256 or 512", a.D);
  FFC_REQUIRE(a.W_bf16 && a.P_bf16, "tcgen05 sweep: bf16 operands missing");
  CUtensorMap mp, mw1, mw2;
  int rc;
  if ((rc = get_map(cache, a.P_bf16, a.n_rows, a.D, BM, &mp))) return rc;
  if ((rc = get_map(cache, a.W_bf16, a.n_cols, a.D, BN, &mw1))) return rc;
  if ((rc = get_map(cache, a.W_bf16, a.n_cols, a.D, JB, &mw2))) return rc;
  Sm100Params p;
  p.n_rows = a.n_rows;
  p.D = a.D;
  p.n_cols = a.n_cols;
  p.n_cols_dev = a.n_cols_dev;
  p.tcol = a.tcol;
  p.cmask = a.cmask;
  p.thr = a.thr;
  p.is_out = a.is_out;
  p.a2 = a.scale * LOG2E;
  p.b2 = a.fixed_max * LOG2E;
  p.sv = a.sv;
  p.k = a.k;
  p.n_chunks = a.n_chunks;
  const int64_t n_tiles = std::max<int64_t>(1, ceil_div64(a.n_cols, BN));
  p.tiles_per_chunk = (int)ceil_div64(n_tiles, a.n_chunks);
  p.ns2 = ns2_for(a.D);
  {
    const char* dbg = getenv("FFC_SM100_DEBUG");
    p.debug = dbg ? atoi(dbg) : 0;
  }
  p.l_part = a.l_part;
  p.o_part = a.o_part;
  p.topv_part = a.topv_part;
  p.topi_part = a.topi_part;
  const int nkc = a.D / KC;
  const size_t smem_s = OFF_DATA + (size_t)nkc * CHUNK1_BYTES + (size_t)NS1 * CHUNK1_BYTES;
  const size_t smem_o = OFF_DATA + NPB * (size_t)PT_BYTES + (size_t)p.ns2 * nkc * JB * KC * 2;
  const size_t smem = std::max(smem_s, smem_o) + 1024;   // slack for the 1024-byte alignment of the dynamic base
  FFC_REQUIRE(smem <= 227 * 1024, "tcgen05 sweep: shared memory budget exceeded (%zu bytes)", smem);
  if (!cache->attr_set) {
    FFC_CUDA(cudaFuncSetAttribute(ffc_head_sweep_sm100_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    FFC_CUDA(cudaFuncSetAttribute(ffc_head_sweep_sm100_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cache->attr_set = true;
  }
  const int row_tiles = (a.n_rows + BM - 1) / BM;
  const int n_items = row_tiles * a.n_chunks;
  dim3 grid(2 * n_items), block(NTHREADS);
  if (a.sv)
    ffc_head_sweep_sm100_kernel<true><<<grid, block, smem, s>>>(mp, mw1, mw2, p);
  else
    ffc_head_sweep_sm100_kernel<false><<<grid, block, smem, s>>>(mp, mw1, mw2, p);
  FFC_LAUNCH_CHECK();
  return FFC_OK;
}

}  // namespace ffc
