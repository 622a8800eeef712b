// Shared helpers for libffc_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/ffc_b200.h"

namespace ffc {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define FFC_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ffc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return FFC_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define FFC_REQUIRE(cond, ...)                                                              \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      ffc::set_error(__VA_ARGS__);                                                          \
      return FFC_ERR_INVALID;                                                               \
    }                                                                                       \
  } while (0)

#define FFC_LAUNCH_CHECK()                                                                  \
  do {                                                                                      \
    ffc::count_launch();                                                                    \
    FFC_CUDA(cudaGetLastError());                                                           \
  } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t next_pow2(int64_t v) {
  int64_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

constexpr int64_t KEY_EMPTY = INT64_MIN;      // never-used hash cell
constexpr int64_t KEY_TOMB = INT64_MIN + 1;   // deleted hash cell

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

}  // namespace ffc
