// Device-side pieces of the GP variance trigger shared by the stand-alone trigger kernel (gp.cu) and the fused
// rollout-step kernel (lstm_tc.cu).  See gp.cu for the math and the reference line citations.
#pragma once
#include "common.cuh"

namespace dvg {

constexpr int MAX_WINDOW = 128;

__device__ __forceinline__ float np_pairwise_sum(const float* a, int n) {
  // numpy's pairwise_sum for n <= 128 (float32 add.reduce of a contiguous vector)
  if (n < 8) {
    float r = 0.f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
    return r;
  }
  float r[8];
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, a[i]);
  return res;
}


// |mat k|^2 for one (rollout, dim) task.  mat is Linv (lower triangular, upper == false -> |v|^2) or L_q^T (upper
// triangular, upper == true -> |w|^2) of the dim, staged in shared memory; only the non-zero 4-wide blocks of each
// row are visited (bounds are compile-time for MREG > 0).  k_m = s exp(-0.5 t^2) uses the 1-instruction MUFU
// exp2 (rel. err 2^-22, far inside the 1e-4 variance bar): the precise expf was a third of the instructions.
template <int MREG>
__device__ __forceinline__ float gp_trig_partial(float xv, float sc, float inv_ell, int MP, const float* mat,
                                                 const float* s_z, bool upper) {
  float part = 0.f;
  if (MREG > 0) {
    float k[MREG > 0 ? MREG : 1];
#pragma unroll
    for (int m = 0; m < MREG; ++m) {
      const float t = (xv - s_z[m]) * inv_ell;
      k[m] = sc * ex2_ftz(t * t * (-0.5f * kLog2e));
    }
    if (!upper) {
#pragma unroll
      for (int j = 0; j < MREG; ++j) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int m = 0; m <= (j / 4) * 4; m += 4) {
          const float4 l4 = *reinterpret_cast<const float4*>(mat + j * MREG + m);
          a0 = fmaf(l4.x, k[m], a0); a1 = fmaf(l4.y, k[m + 1], a1); a0 = fmaf(l4.z, k[m + 2], a0); a1 = fmaf(l4.w, k[m + 3], a1);
        }
        const float a = a0 + a1;
        part = fmaf(a, a, part);
      }
    } else {
#pragma unroll
      for (int j = 0; j < MREG; ++j) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int m = (j / 4) * 4; m < MREG; m += 4) {
          const float4 l4 = *reinterpret_cast<const float4*>(mat + j * MREG + m);
          a0 = fmaf(l4.x, k[m], a0); a1 = fmaf(l4.y, k[m + 1], a1); a0 = fmaf(l4.z, k[m + 2], a0); a1 = fmaf(l4.w, k[m + 3], a1);
        }
        const float a = a0 + a1;
        part = fmaf(a, a, part);
      }
    }
  } else {
    for (int j = 0; j < MP; ++j) {
      float a0 = 0.f, a1 = 0.f;
      const int m_lo = upper ? (j / 4) * 4 : 0, m_hi = upper ? MP : (j / 4) * 4 + 4;
      for (int m = m_lo; m < m_hi; m += 4) {
        const float4 l4 = *reinterpret_cast<const float4*>(mat + j * MP + m);
        float kq[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float t = (xv - s_z[m + e]) * inv_ell;
          kq[e] = sc * ex2_ftz(t * t * (-0.5f * kLog2e));
        }
        a0 = fmaf(l4.x, kq[0], a0); a1 = fmaf(l4.y, kq[1], a1); a0 = fmaf(l4.z, kq[2], a0); a1 = fmaf(l4.w, kq[3], a1);
      }
      const float a = a0 + a1;
      part = fmaf(a, a, part);
    }
  }
  return part;
}


// Compact-code variant of gp_trig_partial<M> for single-warp callers (the step kernel's auxiliary warp): rows are
// visited in two ROLLED loops whose bodies cover m in [0, M/2) / [0, M) (lower) or [0, M) / [M/2, M) (upper).  The extra
// terms are exact zeros of the triangular factor, added in the same a0/a1 order, so the result is bit-identical to
// the fully unrolled version while the code is ~150 instructions instead of ~1700 (a lone warp running straight-line
// code is instruction-fetch bound).  mat may live in global memory (warp-uniform addresses, L1 broadcast).
template <int M, int M_LO, int M_HI>
__device__ __forceinline__ float gp_trig_rows(const float* __restrict__ mat, const float (&k)[M], int j_lo, int j_hi,
                                              float part) {
#pragma unroll 1
  for (int j = j_lo; j < j_hi; ++j) {
    const float4* row = reinterpret_cast<const float4*>(mat + j * M);
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int m = M_LO; m < M_HI; m += 4) {
      const float4 l4 = row[m >> 2];
      a0 = fmaf(l4.x, k[m], a0); a1 = fmaf(l4.y, k[m + 1], a1); a0 = fmaf(l4.z, k[m + 2], a0); a1 = fmaf(l4.w, k[m + 3], a1);
    }
    const float a = a0 + a1;
    part = fmaf(a, a, part);
  }
  return part;
}
// k_m = s exp(-0.5 ((x - z_m) / ell)^2) for all M inducing points, in registers.
template <int M>
__device__ __forceinline__ void gp_trig_kvec(float xv, float sc, float inv_ell, const float* __restrict__ z, float (&k)[M]) {
#pragma unroll
  for (int m = 0; m < M; m += 4) {
    const float4 z4 = *reinterpret_cast<const float4*>(z + m);
    const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float t = (xv - zz[e]) * inv_ell;
      k[m + e] = sc * ex2_ftz(t * t * (-0.5f * kLog2e));
    }
  }
}
template <int M>
__device__ __forceinline__ void gp_trig_partial_rolled(float xv, float sc, float inv_ell, const float* __restrict__ linv,
                                                       const float* __restrict__ lqt, const float* __restrict__ z,
                                                       float& pv, float& pw) {
  static_assert(M % 8 == 0, "M/2 must be a multiple of 4");
  float k[M];
#pragma unroll
  for (int m = 0; m < M; m += 4) {
    const float4 z4 = *reinterpret_cast<const float4*>(z + m);
    const float zz[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float t = (xv - zz[e]) * inv_ell;
      k[m + e] = sc * ex2_ftz(t * t * (-0.5f * kLog2e));
    }
  }
  pv = 0.f; pw = 0.f;
  if (linv != nullptr) {
    pv = gp_trig_rows<M, 0, M / 2>(linv, k, 0, M / 2, 0.f);
    pv = gp_trig_rows<M, 0, M>(linv, k, M / 2, M, pv);
  }
  if (lqt != nullptr) {
    pw = gp_trig_rows<M, 0, M>(lqt, k, 0, M / 2, 0.f);
    pw = gp_trig_rows<M, M / 2, M>(lqt, k, M / 2, M, pw);
  }
}

// Decision for rollout s from the transposed variance scratch var_rows[D][S] (one thread per rollout):
// fp32 norm over d in numpy's sequential order (generate_frames.py:230), window slide (:231), threshold (:288),
// strict '>' (:289); fired rollouts are appended to trig_list.
__device__ __forceinline__ void gp_trig_finalize_rollout(int s, int S, int D, const float* var_rows, float* window,
                                                         int W, int cnt, int warmup, float factor, float* value,
                                                         float* thr, uint8_t* mask, int* trig_list, int* trig_count) {
  float acc = 0.f;
  for (int d0 = 0; d0 < D; d0 += 48) {      // register batches: all loads of a batch are in flight together
    float v[48];
#pragma unroll
    for (int u = 0; u < 48; ++u) v[u] = d0 + u < D ? __ldcg(var_rows + (size_t)(d0 + u) * S + s) : 0.f;
#pragma unroll
    for (int u = 0; u < 48; ++u)
      if (d0 + u < D) acc = __fadd_rn(acc, __fmul_rn(v[u], v[u]));
  }
  const float val = sqrtf(acc);
  float* wdw = window + (size_t)s * W;
  int fired = 0;
  if (value) value[s] = val;
  if (warmup) {
    if (cnt < W) wdw[cnt] = val;
    else {
      for (int q = 0; q + 1 < W; ++q) wdw[q] = wdw[q + 1];
      wdw[W - 1] = val;
    }
    if (thr) thr[s] = nanf("");
  } else {
    float loc[MAX_WINDOW];
    for (int q = 0; q + 1 < W; ++q) loc[q] = wdw[q + 1];
    loc[W - 1] = val;
    for (int q = 0; q < W; ++q) wdw[q] = loc[q];
    const float mean = __fdiv_rn(np_pairwise_sum(loc, W), (float)W);
    for (int q = 0; q < W; ++q) {
      const float dlt = __fsub_rn(loc[q], mean);
      loc[q] = __fmul_rn(dlt, dlt);
    }
    const float sd = sqrtf(__fdiv_rn(np_pairwise_sum(loc, W), (float)W));
    const float t = __fadd_rn(mean, __fmul_rn(factor, sd));
    fired = val > t ? 1 : 0;
    if (thr) thr[s] = t;
  }
  if (mask) mask[s] = (uint8_t)fired;
  if (fired) trig_list[atomicAdd(trig_count, 1)] = s;
}


// Register-resident variant for short windows (W <= 16; the reference uses 12): every loop is unrolled over the 16
// slots with q < W guards, so the window never touches local memory.  Same arithmetic, same order.
__device__ __forceinline__ float np_pairwise_sum16(const float (&a)[16], int n) {
  if (n < 8) {
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (i < n) r = __fadd_rn(r, a[i]);
    return r;
  }
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  if (n == 16) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[8 + j]);
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  if (n != 16) {
#pragma unroll
    for (int i = 8; i < 16; ++i)
      if (i < n) res = __fadd_rn(res, a[i]);
  }
  return res;
}

__device__ __forceinline__ void gp_trig_finalize_rollout16(int s, int S, int D, const float* var_rows, float* window,
                                                           int W, int cnt, int warmup, float factor, float* value,
                                                           float* thr, uint8_t* mask, int* trig_list, int* trig_count) {
  float* wdw = window + (size_t)s * W;
  float loc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) loc[q] = (!warmup && q + 1 < W) ? wdw[q + 1] : 0.f;   // issued before the norm loads return
  float acc = 0.f;
  for (int d0 = 0; d0 < D; d0 += 48) {
    float v[48];
#pragma unroll
    for (int u = 0; u < 48; ++u) v[u] = d0 + u < D ? __ldcg(var_rows + (size_t)(d0 + u) * S + s) : 0.f;
#pragma unroll
    for (int u = 0; u < 48; ++u)
      if (d0 + u < D) acc = __fadd_rn(acc, __fmul_rn(v[u], v[u]));
  }
  const float val = sqrtf(acc);
  int fired = 0;
  if (value) value[s] = val;
  if (warmup) {
    if (cnt < W) wdw[cnt] = val;
    else {
      for (int q = 0; q + 1 < W; ++q) wdw[q] = wdw[q + 1];
      wdw[W - 1] = val;
    }
    if (thr) thr[s] = nanf("");
  } else {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      if (q == W - 1) loc[q] = val;
      if (q < W) wdw[q] = loc[q];
    }
    const float mean = __fdiv_rn(np_pairwise_sum16(loc, W), (float)W);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float dlt = __fsub_rn(loc[q], mean);
      loc[q] = q < W ? __fmul_rn(dlt, dlt) : 0.f;
    }
    const float sd = sqrtf(__fdiv_rn(np_pairwise_sum16(loc, W), (float)W));
    const float t = __fadd_rn(mean, __fmul_rn(factor, sd));
    fired = val > t ? 1 : 0;
    if (thr) thr[s] = t;
  }
  if (mask) mask[s] = (uint8_t)fired;
  if (fired) trig_list[atomicAdd(trig_count, 1)] = s;
}

}  // namespace dvg
