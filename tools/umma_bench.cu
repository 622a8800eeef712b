// Micro-benchmark: cycles per tcgen05.mma (bf16, M=128 per CTA) for the operand placements the FFC sweep can choose
// between.  One CTA (or CTA pair) per SM issues NITER back-to-back MMAs on whatever is in shared memory / TMEM and
// reports (t1 - t0) / NITER.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_bench.bin tools/umma_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc),
               "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc),
               "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_ss2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc),
               "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mma_ts2(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc),
               "r"(acc)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  asm volatile("{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WD;\nbra WL;\nWD:\n}\n" ::"r"(addr), "r"(parity) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n.reg .pred px;\nelect.sync _|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}\n" : "=r"(pred));
  return pred != 0;
}

struct Res {
  long long cyc;
};

// Lean issue loop: everything but the start addresses is a compile-time constant, one elected lane issues 32 MMAs per
// iteration (8 K-chunks x 4 k16 steps, like one [128 x N x 512] tile of the FFC sweep), D alternates between two
// accumulators every DPER MMAs.
template <int CG, int TS, int N, int BMN, int DPER>
__global__ void __launch_bounds__(128, 1) umma_bench(int niter, Res* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(&tmem_base_s);
  if (warp == 1 && rank == 0) {
    constexpr int M = CG == 2 ? 256 : 128;
    constexpr uint32_t idesc = make_idesc(M, N, 0, BMN);
    constexpr int nloc = CG == 2 ? N / 2 : N;
    constexpr uint32_t dstride = TS ? (N <= 128 ? (uint32_t)N : 0u) : (N == 256 ? 256u : (uint32_t)N);
    const uint64_t a0 = make_desc(smem_u32(smem), 16, 1024, 2);
    const uint64_t b0 = BMN ? make_desc(smem_u32(smem) + 64 * 1024, 32 * 128, 1024, 2) : make_desc(smem_u32(smem) + 64 * 1024, 16, 1024, 2);
    long long t0 = clock64();
    if (elect_one()) {
#pragma unroll 1
      for (int i = 0; i < niter; i += 32) {
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            constexpr int dummy = 0;
            const int j = kc * 4 + k;
            const uint32_t dd = tmem_base + (uint32_t)((j / DPER) & 1) * dstride;
            const uint64_t bdesc = b0 + (uint64_t)(BMN ? (((kc & 3) * 32768 + (k & 1) * 16 * 128) >> 4) : (((kc & 3) * (nloc * 128) + k * 32) >> 4));
            if (TS) {
              const uint32_t a = tmem_base + 256u + (uint32_t)(j * 8);
              if (CG == 1) mma_ts(dd, a, bdesc, idesc, 1u); else mma_ts2(dd, a, bdesc, idesc, 1u);
            } else {
              const uint64_t adesc = a0 + (uint64_t)((kc * 16384 + k * 32) >> 4);
              if (CG == 1) mma_ss(dd, adesc, bdesc, idesc, 1u); else mma_ss2(dd, adesc, bdesc, idesc, 1u);
            }
            (void)dummy;
          }
        }
      }
      if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                     "h"((uint16_t)1)
                     : "memory");
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[blockIdx.x / CG].cyc = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) {
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

static Res* d_res;
static const size_t SMEM = 201 * 1024 + 1024;

template <int CG, int TS, int N, int BMN, int DPER>
void run(const char* note) {
  const int niter = 4096;
  auto kern = umma_bench<CG, TS, N, BMN, DPER>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(d_res, 0, 148 * sizeof(Res));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = SMEM;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, niter, d_res);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("CUDA error %s\n", cudaGetErrorString(e));
      exit(1);
    }
  }
  Res h[148];
  cudaMemcpy(h, d_res, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1ll << 60;
  for (int i = 0; i < 148 / CG; ++i) {
    mx = h[i].cyc > mx ? h[i].cyc : mx;
    mn = h[i].cyc < mn ? h[i].cyc : mn;
  }
  const double cyc = (double)mx / niter;
  printf("cta_group::%d %s N=%3d B %s-major D switch/%-4d: %7.1f cyc/MMA (min %.1f) -> %5.0f MAC/clk/SM (floor %d) %s\n", CG, TS ? "TS" : "SS", N,
         BMN ? "MN" : "K ", DPER, cyc, (double)mn / niter, 128.0 * N * 16 / cyc, N / 2, note);
}

int main() {
  cudaMalloc(&d_res, 148 * sizeof(Res));
  run<1, 0, 64, 0, 64>("");
  run<1, 0, 128, 0, 64>("GEMM-1 as is");
  run<1, 0, 128, 0, 1>("");
  run<1, 0, 256, 0, 64>("");
  run<1, 0, 256, 1, 64>("");
  run<1, 0, 256, 1, 1>("GEMM-2 as is (two N=256 halves alternate)");
  run<1, 0, 256, 1, 4>("");
  run<1, 1, 64, 0, 64>("");
  run<1, 1, 128, 0, 64>("GEMM-1 with P in TMEM");
  run<1, 1, 128, 0, 1>("");
  run<1, 1, 256, 0, 64>("");
  run<1, 1, 256, 1, 64>("");
  run<2, 0, 128, 0, 64>("");
  run<2, 0, 256, 0, 64>("");
  run<2, 0, 256, 1, 64>("");
  run<2, 0, 256, 1, 1>("");
  run<2, 1, 128, 0, 64>("");
  run<2, 1, 256, 0, 64>("");
  run<2, 1, 256, 1, 64>("");
  return 0;
}
