// Shared host/device helpers for the dvg_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/dvg_b200.h"

namespace dvg {

void set_error(const char* fmt, ...);

#define DVG_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      dvg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return DVG_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define DVG_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      dvg::set_error(__VA_ARGS__);      \
      return DVG_ERR_ARG;               \
    }                                   \
  } while (0)

#define DVG_LAUNCH_CHECK() DVG_CUDA(cudaGetLastError())

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---- tile geometry of the packed bf16 operand images (tensor-core variants) -----------------------
// An operand "k-block" is a [rows x 64] bf16 tile stored exactly as tcgen05 wants it in shared memory:
// K-major, 128-byte rows, SWIZZLE_128B (16-byte chunk index XOR (row & 7)), 8-row groups 1024 B apart.
// A (activation) images have 128 rows (16 KB); B (weight) images have N_TILE rows.
// Global layout: [tile][k_block][part: hi, lo][rows * 128 B]  -> one cp.async.bulk per (tile, k_block).
constexpr int TC_ROWS = 128;            // UMMA M
constexpr int TC_KBLK = 64;             // bf16 elements per k-block (= 128 B swizzle span)
constexpr int TC_A_IMG = TC_ROWS * 128; // bytes of one A image part

__host__ __device__ inline uint32_t sw128_offset(uint32_t row, uint32_t chunk16) {
  return row * 128u + ((chunk16 ^ (row & 7u)) << 4);
}

#ifdef __CUDACC__
// Exact-ish fp32 activations.  MUFU ex2/rcp based (rel. err ~2 ulp), no fast-math flags needed.
__device__ __forceinline__ float sigmoid_f(float x) {
  return __fdividef(1.0f, 1.0f + __expf(-x));
}
__device__ __forceinline__ float tanh_f(float x) {
  // |x| >= 0.25: tanh(x) = 1 - 2/(exp(2x)+1) (abs err ~1e-7, saturates cleanly);
  // |x| <  0.25: odd Taylor polynomial to x^9 (truncation < 3e-9) so small values keep relative accuracy.
  float e = __expf(2.0f * x);
  float big = 1.0f - __fdividef(2.0f, e + 1.0f);
  float x2 = x * x;
  float p = fmaf(x2, 62.0f / 2835.0f, -17.0f / 315.0f);
  p = fmaf(x2, p, 2.0f / 15.0f);
  p = fmaf(x2, p, -1.0f / 3.0f);
  float small = fmaf(x * x2, p, x);
  return fabsf(x) < 0.25f ? small : big;
}

// a ~= hi + lo with both bf16 (round-to-nearest): relative residual <= 2^-17.
__device__ __forceinline__ void split_bf16(float a, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(a);
  lo = __float2bfloat16_rn(a - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack2_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
#endif

}  // namespace dvg
