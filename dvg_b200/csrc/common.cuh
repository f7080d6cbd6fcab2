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

// ---- fast gate math for the tensor-core epilogues --------------------------------------------------------
// MUFU-only building blocks (1 instruction each, no range fix-ups).
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;
// 2^x on the FMA pipe (Cody-Waite + degree-5 minimax, max rel. err 1.9e-7 ~ MUFU.EX2): B200's MUFU pipe
// sustains only ~8 results/clk/SM, which made the LSTM epilogue MUFU-bound; x must be <= 64.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;        // 1.5 * 2^23: the integer part lands in the low mantissa bits
  const float n = t - 12582912.f;
  const float f = x - n;                 // [-0.5, 0.5]
  float p = 0.001326472731307149f;
  p = fmaf(p, f, 0.009671512991189957f);
  p = fmaf(p, f, 0.05550733581185341f);
  p = fmaf(p, f, 0.24022242426872253f);
  p = fmaf(p, f, 0.6931470036506653f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));   // * 2^n
}
// LSTM cell from gate pre-activations.  a* are the raw accumulators, b* the biases PRE-SCALED by -log2(e)
// (i, f, o) and -2 log2(e) (g), so every exponent argument is one FFMA:
//   E_i = e^-z_i, E_f = e^-z_f, E_o = e^-z_o, E_g = e^-2 z_g;  sigmoid(z) = 1/(1+E),  tanh(z) = (1-E)/(1+E)
//   c' = sig(f) c + sig(i) tanh(g) = [c (1+E_i)(1+E_g) + (1-E_g)(1+E_f)] / [(1+E_f)(1+E_i)(1+E_g)]   (ONE reciprocal)
//   h' = sig(o) tanh(c')          = (1-E_c) / [(1+E_o)(1+E_c)],  E_c = e^-2 c'                     (ONE reciprocal)
// Exponent arguments are clamped at 40 (E <= 2^40, triple products stay finite; sigmoid/tanh are saturated to
// fp32 precision well before).  E_i, E_f, E_o use the FMA-pipe polynomial, E_g, E_c the MUFU: 4 MUFU + ~57 FP32
// ops per hidden unit, balanced between the two pipes.  Abs. error ~3e-7.
// NPOLY of the three sigmoid exponentials use the FMA-pipe polynomial, the rest the MUFU (the rolled epilogue of
// lstm_step.cu is issue bound, not MUFU bound: ~80 issue slots per cell at NPOLY = 3 against 5 XU results at 16
// lanes/clk/SM; all-MUFU, NPOLY = 0, was fastest; the fully unrolled per-GEMM epilogues keep 3).
template <int NPOLY = 3>
__device__ __forceinline__ void lstm_cell_fast(float ai, float af, float ag, float ao, float bi, float bf, float bg,
                                               float bo, float c_prev, float& h_new, float& c_new) {
  const float Ei = NPOLY >= 1 ? ex2_poly(fminf(fmaf(ai, -kLog2e, bi), 40.f)) : ex2_ftz(fminf(fmaf(ai, -kLog2e, bi), 40.f));
  const float Ef = NPOLY >= 2 ? ex2_poly(fminf(fmaf(af, -kLog2e, bf), 40.f)) : ex2_ftz(fminf(fmaf(af, -kLog2e, bf), 40.f));
  const float Eo = NPOLY >= 3 ? ex2_poly(fminf(fmaf(ao, -kLog2e, bo), 40.f)) : ex2_ftz(fminf(fmaf(ao, -kLog2e, bo), 40.f));
  const float Eg = ex2_ftz(fminf(fmaf(ag, -2.f * kLog2e, bg), 40.f));
  const float pi = 1.f + Ei, pf = 1.f + Ef, pg = 1.f + Eg;
  const float P = pi * pg;
  const float num = fmaf(c_prev, P, (1.f - Eg) * pf);
  c_new = num * rcp_ftz(P * pf);
  const float Ec = ex2_ftz(fminf(c_new * (-2.f * kLog2e), 40.f));
  h_new = (1.f - Ec) * rcp_ftz((1.f + Eo) * (1.f + Ec));
}
// tanh(z + b) with the bias pre-scaled by -2 log2(e): 2 MUFU.
__device__ __forceinline__ float tanh_fast_prescaled(float a, float b) {
  const float E = ex2_ftz(fminf(fmaf(a, -2.f * kLog2e, b), 40.f));
  return (1.f - E) * rcp_ftz(1.f + E);
}

// Same with the exponential on the FMA pipe (ex2_poly): 1 MUFU.  The head epilogue of lstm_step.cu is XU bound (only
// transcendentals, no other math), so it alternates the two forms.
__device__ __forceinline__ float tanh_fast_prescaled_poly(float a, float b) {
  const float E = ex2_poly(fminf(fmaf(a, -2.f * kLog2e, b), 40.f));
  return (1.f - E) * rcp_ftz(1.f + E);
}

// a ~= hi + lo with both bf16 (round-to-nearest): relative residual <= 2^-17.
__device__ __forceinline__ void split_bf16(float a, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(a);
  lo = __float2bfloat16_rn(a - __bfloat162float(hi));
}
// Two values at once with the packed converter (cvt.rn.bf16x2.f32): returns hi pair / lo pair as bf16x2 words
// (element 0 in the low half), same RN/RN arithmetic as split_bf16.
__device__ __forceinline__ void split2_bf16(float a0, float a1, uint32_t& hi2, uint32_t& lo2) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(a1), "f"(a0));
  const float h0 = __uint_as_float(hi2 << 16), h1 = __uint_as_float(hi2 & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(a1 - h1), "f"(a0 - h0));
}
__device__ __forceinline__ uint32_t pack2_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
#endif

}  // namespace dvg
