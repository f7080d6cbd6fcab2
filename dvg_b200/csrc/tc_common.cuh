// Device helpers shared by the tensor-core LSTM kernels (lstm_tc.cu: one launch per GEMM; lstm_step.cu: the
// persistent whole-step kernel).
#pragma once
#include "internal.cuh"
#include "ptx.cuh"

namespace dvg {

constexpr int TRACE_SLOTS = 128;
#ifdef DVG_TRACE
#define TRACE(slot)                                                                   \
  do {                                                                                \
    if (p.trace) {                                                                    \
      unsigned long long _t;                                                          \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t) :: "memory");                          \
      p.trace[(size_t)blockIdx.x * TRACE_SLOTS + (slot)] = _t;                                 \
    }                                                                                 \
  } while (0)
#else
#define TRACE(slot) do {} while (0)
#endif
constexpr int EPI_WARPS = 8;
constexpr int TC_THREADS = 64 + EPI_WARPS * 32;
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;  // TMEM columns per accumulator buffer

// Thread-per-row accesses touch 32 different 128-byte lines per warp instruction, and the L1TEX cost is per
// line touched, not per byte: use the 256-bit LDG/STG of sm_100 so each line is visited as rarely as possible.
__device__ __forceinline__ void ld256(const float* p, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void st256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void st256u(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void load16(const float* p, float (&v)[16]) {
  ld256(p, v);
  ld256(p + 8, v + 8);
}
__device__ __forceinline__ void store16(float* p, const float (&v)[16]) {
  st256(p, v);
  st256(p + 8, v + 8);
}

__device__ __forceinline__ void store_split16(uint8_t* img_hi, uint32_t r_in_tile, uint32_t chunk0, const float (&v)[16]) {
  // 16 consecutive K elements of one row -> two 16-byte chunks in the hi image and two in the lo image.
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split2_bf16(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
  uint8_t* img_lo = img_hi + TC_A_IMG;
  // chunk0 is even: the swizzled positions of chunks {chunk0, chunk0+1} form one aligned 32-byte sector,
  // in swapped order when bit 0 of (row & 7) is set -> one 256-bit store per image.
  const uint32_t o0 = sw128_offset(r_in_tile, chunk0), o1 = sw128_offset(r_in_tile, chunk0 + 1);
  const bool swap = o1 < o0;
  const uint32_t ob = swap ? o1 : o0;
  uint32_t th[8], tl[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    th[i] = swap ? hi[4 + i] : hi[i];
    th[4 + i] = swap ? hi[i] : hi[4 + i];
    tl[i] = swap ? lo[4 + i] : lo[i];
    tl[4 + i] = swap ? lo[i] : lo[4 + i];
  }
  st256u(img_hi + ob, th);
  st256u(img_lo + ob, tl);
}

// LSTM pointwise math for 16 hidden units of one row (i,f,g,o pre-activations in r[0..63]).
__device__ __forceinline__ void lstm_pointwise16(const uint32_t (&r)[64], const float* sb, int cb, const float (&cp)[16],
                                                 float (&hn)[16], float (&cn)[16]) {
  // sb holds the biases pre-scaled by -log2e (i, f, o) / -2 log2e (g): see lstm_cell_fast
#pragma unroll
  for (int i = 0; i < 16; ++i)
    lstm_cell_fast(__uint_as_float(r[i]), __uint_as_float(r[16 + i]), __uint_as_float(r[32 + i]),
                   __uint_as_float(r[48 + i]), sb[cb + i], sb[64 + cb + i], sb[128 + cb + i], sb[192 + cb + i], cp[i],
                   hn[i], cn[i]);
}
}  // namespace dvg
