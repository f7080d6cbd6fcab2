// Bookkeeping kernels for the N-diverse-futures rollout.
//
// dvg_rollout_score: per (rollout s, sequence b) mean squared error of the generated latent sequence against a
// target latent sequence -- the device-side scoring pass of the best-of-N selection (the reference scores on
// the host after a D2H copy of every frame, generate_frames.py:175-178,185-190).  One pass over the [T, S*B, G]
// tensor: HBM bound, algorithmic bytes = T*S*B*G*4 (+ the small target).
#include "internal.cuh"

namespace dvg {

// Generic fallback: one warp per row (s*B + b), lanes stride over g, loop over t; 8 rows per CTA.  Used when the time
// slices are not 16-byte aligned ((S*B*G) % 4 != 0) or G > 128.
__global__ void __launch_bounds__(256) rollout_score_generic_kernel(int T, int S, int B, int G,
                                                                    const float* __restrict__ out,
                                                                    const float* __restrict__ target,
                                                                    float* __restrict__ scores) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  const int R = S * B;
  if (row >= R) return;
  const int b = row % B;
  float acc = 0.f;
  for (int t = 0; t < T; ++t) {
    const float* o = out + ((size_t)t * R + row) * G;
    const float* g = target + ((size_t)t * B + b) * G;
    for (int i = lane; i < G; i += 32) {
      const float dlt = __ldg(o + i) - __ldg(g + i);
      acc = fmaf(dlt, dlt, acc);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) scores[row] = acc / (float)((size_t)T * G);
}

// Streaming version (G % 4 == 0 or not, any G with 8*G % 4 == 0): a CTA owns 8 consecutive rows = 8*G contiguous floats of
// every time slice, read as float4 (thread j always reads elements 4j .. 4j+3 of the block, which lie in at most two
// rows, so it keeps two running sums in registers); 13 independent 16-byte loads per thread are in flight per batch.
// Deterministic: fixed per-thread order over t, then a fixed-order shared-memory reduction per row.
// (History: one warp per row with scalar loads ran at 1.6 TB/s, 44 us for the 70 MB of kth_s100.)
__global__ void __launch_bounds__(256) rollout_score_kernel(int T, int S, int B, int G, const float* __restrict__ out,
                                                            const float* __restrict__ target,
                                                            float* __restrict__ scores) {
  __shared__ float s_sum[8][65];            // [row in block][contributing thread slot]
  const int R = S * B;
  const int row0 = blockIdx.x * 8;
  const int n4 = 2 * G;                     // float4 per block per time slice (8 * G / 4)
  const int j = threadIdx.x;
  float acc0 = 0.f, acc1 = 0.f;
  int ra = 0, rb = 0, split = 4;            // elements [0, split) of this thread's float4 belong to row ra, the rest to rb
  if (j < n4) {
    const int e0 = 4 * j;
    ra = e0 / G;
    rb = (e0 + 3) / G;
    split = rb == ra ? 4 : (ra + 1) * G - e0;
  }
  const bool live = j < n4 && row0 + ra < R;
  constexpr int TB = 13;
  if (live) {
    const size_t blk = (size_t)row0 * G + 4 * (size_t)j;
    const int ga = (4 * j) % G;             // column of element 0 within row ra
    // per-element target offsets are the same for every time slice: computed once (the first version re-derived
    // row % B and the column for every element of every slice and was instruction bound)
    int toff[4];
    bool ok[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool first = e < split;
      const int row = row0 + (first ? ra : rb);
      ok[e] = row < R;
      toff[e] = ok[e] ? (row % B) * G + (first ? ga + e : e - split) : 0;
    }
    const bool whole = row0 + rb < R;       // the float4 stays inside the valid rows of the tensor
    const size_t slice_o = (size_t)R * G, slice_t = (size_t)B * G;
    for (int t0 = 0; t0 < T; t0 += TB) {
      float4 v[TB];
#pragma unroll
      for (int u = 0; u < TB; ++u)
        if (t0 + u < T) {
          const float* src = out + (size_t)(t0 + u) * slice_o + blk;
          if (whole) {
            v[u] = __ldcs(reinterpret_cast<const float4*>(src));
          } else {            // would run past the last valid row: element-wise
            float x[4] = {0.f, 0.f, 0.f, 0.f};
            for (int e = 0; e < split; ++e) x[e] = __ldg(src + e);
            v[u] = make_float4(x[0], x[1], x[2], x[3]);
          }
        }
#pragma unroll
      for (int u = 0; u < TB; ++u) {
        if (t0 + u >= T) continue;
        const float x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        const float* tg = target + (size_t)(t0 + u) * slice_t;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (ok[e]) {
            const float dlt = x[e] - __ldg(tg + toff[e]);
            if (e < split) acc0 = fmaf(dlt, dlt, acc0);
            else acc1 = fmaf(dlt, dlt, acc1);
          }
        }
      }
    }
  }
  // per-row reduction: thread j contributes acc0 to row ra and acc1 to row rb; a row is covered by <= G/4 + 2 threads
  for (int i = threadIdx.x; i < 8 * 65; i += 256) (&s_sum[0][0])[i] = 0.f;
  __syncthreads();
  const int per_row = G / 4 + 2;            // thread slots per row
  if (j < n4 && per_row <= 64) {
    // slot of this thread within its row(s): threads are consecutive, the first thread touching row r is floor(r*G/4)
    s_sum[ra][j - (ra * G) / 4] = acc0;
    if (rb != ra && rb < 8) s_sum[rb][64] = acc1;     // at most one thread straddles into row rb from below
  }
  __syncthreads();
  if (threadIdx.x < 8 && row0 + threadIdx.x < R) {
    float tot = s_sum[threadIdx.x][64];
    for (int i = 0; i < 64; ++i) tot += s_sum[threadIdx.x][i];
    scores[row0 + threadIdx.x] = tot / (float)((size_t)T * G);
  }
}

// ---------------------------------------------------------------------------------------------------
// Frame metrics on the device: per (frame t, sample s, sequence b) the channel-mean SSIM and PSNR of a generated frame
// against the ground truth, without the per-frame D2H copy of generate_frames.py:175-178.  Two variants:
//   EV_FINN     utils.finn_eval_seq (utils.py:237-301): 11x11 Gaussian window (sigma 1.5), 'valid' region, K1=.01,
//               K2=.03, L=1, population moments, NaN -> -1; PSNR = 10 log10(1 / mse).
//   EV_SKIMAGE  utils.eval_seq (utils.py:220-234) = legacy skimage.measure.compare_ssim / compare_psnr defaults, the
//               metric make_gifs actually selects on (generate_frames.py:178): 7x7 uniform window, sample covariance
//               (x 49/48), K1=.01, K2=.03, data_range 2 for float images, mean over the interior cropped by 3 pixels;
//               PSNR = 10 log10(R^2 / mse) with R = 1 when min(ground truth) >= 0 else 2.
// Both windows are separable: horizontal pass into shared memory, then vertical pass + SSIM map + mean, in blocks of
// EV_RB output rows so 128x128 frames fit.  One CTA per (t, s, b).
// ---------------------------------------------------------------------------------------------------
constexpr int EV_RB = 16;     // output rows per block
enum { EV_FINN = 0, EV_SKIMAGE = 1 };

template <int MODE>
__global__ void __launch_bounds__(256) eval_seq_kernel(int T, int S, int B, int C, int H, int W,
                                                       const float* __restrict__ gt, const float* __restrict__ gen,
                                                       float* __restrict__ ssim, float* __restrict__ psnr) {
  constexpr int WIN = MODE == EV_FINN ? 11 : 7;
  extern __shared__ __align__(16) float sm[];
  __shared__ float s_win[WIN];
  __shared__ float s_red[3][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.x % B, s = (blockIdx.x / B) % S, t = blockIdx.x / (B * S);
  const int Wo = W - (WIN - 1), Ho = H - (WIN - 1);
  float* s_a = sm;                                   // [EV_RB+WIN-1][W]
  float* s_b = s_a + (EV_RB + WIN - 1) * W;          // [EV_RB+WIN-1][W]
  float* s_h = s_b + (EV_RB + WIN - 1) * W;          // [5][EV_RB+WIN-1][Wo]
  if (tid < WIN) {
    if (MODE == EV_FINN) {
      float g[WIN], tot = 0.f;
      for (int i = 0; i < WIN; ++i) { const float d = (float)(i - WIN / 2); g[i] = expf(-(d * d) / (2.f * 1.5f * 1.5f)); tot += g[i]; }
      s_win[tid] = g[tid] / tot;
    } else {
      s_win[tid] = 1.0f / (float)WIN;
    }
  }
  __syncthreads();
  const float c1 = MODE == EV_FINN ? 1e-4f : 4e-4f, c2 = MODE == EV_FINN ? 9e-4f : 3.6e-3f;   // (K1 R)^2, (K2 R)^2
  const float cov_norm = MODE == EV_FINN ? 1.0f : (float)(WIN * WIN) / (float)(WIN * WIN - 1);
  float ssim_c_sum = 0.f, psnr_c_sum = 0.f;
  for (int c = 0; c < C; ++c) {
    const float* A = gt + (((size_t)t * B + b) * C + c) * H * W;
    const float* G = gen + ((((size_t)t * S + s) * B + b) * C + c) * H * W;
    float sq = 0.f, acc = 0.f, amin = 3.0e38f;
    for (int e = tid; e < H * W; e += 256) {
      const float av = __ldg(A + e);
      const float d = av - __ldg(G + e);
      sq = fmaf(d, d, sq);
      amin = fminf(amin, av);
    }
    for (int r0 = 0; r0 < Ho; r0 += EV_RB) {
      const int rows_out = Ho - r0 < EV_RB ? Ho - r0 : EV_RB;
      const int rows_in = rows_out + WIN - 1;
      __syncthreads();
      for (int e = tid; e < rows_in * W; e += 256) {
        s_a[e] = __ldg(A + (size_t)r0 * W + e);
        s_b[e] = __ldg(G + (size_t)r0 * W + e);
      }
      __syncthreads();
      for (int e = tid; e < rows_in * Wo; e += 256) {     // horizontal pass: 5 moments
        const int i = e / Wo, j = e - i * Wo;
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < WIN; ++k) {
          const float w = s_win[k], x = s_a[i * W + j + k], y = s_b[i * W + j + k];
          m1 = fmaf(w, x, m1); m2 = fmaf(w, y, m2);
          e11 = fmaf(w * x, x, e11); e22 = fmaf(w * y, y, e22); e12 = fmaf(w * x, y, e12);
        }
        const int o = i * Wo + j, st = (EV_RB + WIN - 1) * Wo;
        s_h[o] = m1; s_h[st + o] = m2; s_h[2 * st + o] = e11; s_h[3 * st + o] = e22; s_h[4 * st + o] = e12;
      }
      __syncthreads();
      for (int e = tid; e < rows_out * Wo; e += 256) {    // vertical pass + SSIM map
        const int i = e / Wo, j = e - i * Wo;
        const int st = (EV_RB + WIN - 1) * Wo;
        float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < WIN; ++k) {
          const float w = s_win[k];
          const int o = (i + k) * Wo + j;
          m1 = fmaf(w, s_h[o], m1); m2 = fmaf(w, s_h[st + o], m2);
          e11 = fmaf(w, s_h[2 * st + o], e11); e22 = fmaf(w, s_h[3 * st + o], e22); e12 = fmaf(w, s_h[4 * st + o], e12);
        }
        const float s11 = cov_norm * (e11 - m1 * m1), s22 = cov_norm * (e22 - m2 * m2), s12 = cov_norm * (e12 - m1 * m2);
        acc += ((2.f * m1 * m2 + c1) * (2.f * s12 + c2)) / ((m1 * m1 + m2 * m2 + c1) * (s11 + s22 + c2));
      }
    }
    // block reduction of (acc, sq, amin)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      acc += __shfl_xor_sync(0xffffffffu, acc, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
      amin = fminf(amin, __shfl_xor_sync(0xffffffffu, amin, o));
    }
    __syncthreads();
    if (lane == 0) { s_red[0][warp] = acc; s_red[1][warp] = sq; s_red[2][warp] = amin; }
    __syncthreads();
    if (tid == 0) {
      float a = 0.f, q = 0.f, mn = 3.0e38f;
      for (int w8 = 0; w8 < 8; ++w8) { a += s_red[0][w8]; q += s_red[1][w8]; mn = fminf(mn, s_red[2][w8]); }
      float sv = a / (float)(Ho * Wo);
      const float mse = q / (float)(H * W);
      if (MODE == EV_FINN) {
        if (sv != sv) sv = -1.f;                                   // utils.py:247-248
        psnr_c_sum += 10.f * log10f(1.f / mse);                    // utils.py:259-261
      } else {
        const float R = mn >= 0.f ? 1.f : 2.f;                     // skimage compare_psnr, float images
        psnr_c_sum += 10.f * log10f(R * R / mse);
      }
      ssim_c_sum += sv;
    }
  }
  if (tid == 0) {
    const size_t o = ((size_t)s * B + b) * T + t;
    ssim[o] = ssim_c_sum / (float)C;
    psnr[o] = psnr_c_sum / (float)C;
  }
}

template <int MODE>
static int eval_seq_launch_t(int T, int S, int B, int C, int H, int W, const float* gt, const float* gen, float* ssim,
                             float* psnr, cudaStream_t stream) {
  constexpr int WIN = MODE == EV_FINN ? 11 : 7;
  DVG_REQUIRE(H >= WIN && W >= WIN && W <= 256, "frame size %dx%d unsupported (need %d <= H, %d <= W <= 256)", H, W, WIN, WIN);
  const size_t smem = sizeof(float) * ((size_t)2 * (EV_RB + WIN - 1) * W + 5 * (size_t)(EV_RB + WIN - 1) * (W - WIN + 1));
  static bool configured = false;
  if (!configured) {
    DVG_CUDA(cudaFuncSetAttribute(eval_seq_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  eval_seq_kernel<MODE><<<T * S * B, 256, smem, stream>>>(T, S, B, C, H, W, gt, gen, ssim, psnr);
  DVG_LAUNCH_CHECK();
  return DVG_OK;
}

int eval_seq_finn_launch(int T, int S, int B, int C, int H, int W, const float* gt, const float* gen, float* ssim,
                         float* psnr, cudaStream_t stream) {
  return eval_seq_launch_t<EV_FINN>(T, S, B, C, H, W, gt, gen, ssim, psnr, stream);
}
int eval_seq_skimage_launch(int T, int S, int B, int C, int H, int W, const float* gt, const float* gen, float* ssim,
                            float* psnr, cudaStream_t stream) {
  return eval_seq_launch_t<EV_SKIMAGE>(T, S, B, C, H, W, gt, gen, ssim, psnr, stream);
}

int rollout_score_launch(int T, int S, int B, int G, const float* out, const float* target, float* scores,
                         cudaStream_t stream) {
  const bool vec = G >= 4 && G <= 128 && ((size_t)S * B * G) % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (vec) rollout_score_kernel<<<ceil_div(S * B, 8), 256, 0, stream>>>(T, S, B, G, out, target, scores);
  else rollout_score_generic_kernel<<<ceil_div(S * B, 8), 256, 0, stream>>>(T, S, B, G, out, target, scores);
  DVG_LAUNCH_CHECK();
  return DVG_OK;
}

}  // namespace dvg
