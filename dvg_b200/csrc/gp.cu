// Variational-GP predictive, variance trigger and rsample for the DVG rollout (fp32 CUDA-core kernels
// with shared-memory staging; factor preparation in fp64).
//
// Replaces gp_layer(x) / likelihood(...) of models/gp_models.py:10-24 (gpytorch WhitenedVariationalStrategy
// eval branch + GaussianLikelihood), the host-side trigger arithmetic of generate_frames.py:227-232,275,
// 283-289 and .rsample() of generate_frames.py:171,292 / train.py:284.
//
// Per latent dimension d (x = column d of the [N,D] latent):
//   k_m   = s exp(-0.5 ((x - z_m)/ell)^2)                         (M exps)
//   v     = L_ZZ^-1 k          (triangular mat-vec with the explicit inverse Linv)                 [Linv hoisted]
//   mean  = c + v . beta                                           beta = L_ZZ^-1 (m_q - c)        [hoisted]
//   w     = L_q^T k
//   var   = s - |v|^2 + |w|^2 + noise
// K_ZZ, its Cholesky factor, Linv and alpha are constant in eval mode; the reference recomputes them on
// every call, here gp_prepare_kernel builds them once per weight load in fp64.
#include "gp_rsample.cuh"
#include "gp_trigger.cuh"
#include "internal.cuh"

namespace dvg {

// ---------------------------------------------------------------------------------------------------
// prepare: one CTA per latent dimension, fp64, M <= 160 (K_ZZ / L in dynamic shared memory)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double softplus_d(double x) { return x > 30.0 ? x : log1p(exp(x)); }

__global__ void __launch_bounds__(128) gp_prepare_kernel(int D, int M, int mp, double jitter, double noise_lb,
                                                         const float* __restrict__ inducing,
                                                         const float* __restrict__ var_mean,
                                                         const float* __restrict__ chol_var,
                                                         const float* __restrict__ mean_const,
                                                         const float* __restrict__ raw_os,
                                                         const float* __restrict__ raw_ls,
                                                         const float* __restrict__ raw_noise, float* __restrict__ z_out,
                                                         float* __restrict__ linv_out, float* __restrict__ lqt_out,
                                                         float* __restrict__ linvT_out, float* __restrict__ lq_out,
                                                         float* __restrict__ alpha_out, float* __restrict__ hyp_out,
                                                         double* __restrict__ work) {
  extern __shared__ double sm[];
  const int d = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const int ld = M + 1;
  double* A = sm;              // [M][M+1]  K_ZZ -> L (lower)
  double* zs = A + M * ld;     // [M]
  double* tv = zs + M;         // [M]
  double* X = work + (size_t)d * M * M;  // [M][M] Linv in fp64 (row-major, column c owned by thread c)

  const double ell = softplus_d((double)raw_ls[d]);
  const double s = softplus_d((double)raw_os[d]);
  const double c = (double)mean_const[d];
  const double noise = softplus_d((double)raw_noise[d]) + noise_lb;
  if (tid == 0) {
    hyp_out[d * 4 + 0] = (float)ell;
    hyp_out[d * 4 + 1] = (float)s;
    hyp_out[d * 4 + 2] = (float)c;
    hyp_out[d * 4 + 3] = (float)noise;
  }
  for (int m = tid; m < mp; m += nt) {
    const float zf = m < M ? inducing[(size_t)d * M + m] : 0.f;
    if (m < M) zs[m] = (double)zf;
    z_out[(size_t)d * mp + m] = zf;
  }
  __syncthreads();
  for (int e = tid; e < M * M; e += nt) {
    const int i = e / M, j = e % M;
    const double t = (zs[i] - zs[j]) / ell;
    A[i * ld + j] = s * exp(-0.5 * t * t) + (i == j ? jitter : 0.0);
  }
  __syncthreads();
  // left-looking Cholesky, column by column
  for (int j = 0; j < M; ++j) {
    double sum = 0.0;
    const int i = j + tid;  // rows j.. handled by threads (loop if M - j > nt)
    for (int ii = i; ii < M; ii += nt) {
      sum = A[ii * ld + j];
      for (int k = 0; k < j; ++k) sum -= A[ii * ld + k] * A[j * ld + k];
      tv[ii] = sum;
    }
    __syncthreads();
    const double diag = sqrt(tv[j]);
    for (int ii = i; ii < M; ii += nt) A[ii * ld + j] = (ii == j) ? diag : tv[ii] / diag;
    __syncthreads();
  }
  // X = L^-1 by forward substitution, one column per thread
  for (int col = tid; col < M; col += nt) {
    for (int i = 0; i < M; ++i) {
      double v = 0.0;
      if (i >= col) {
        v = (i == col) ? 1.0 : 0.0;
        for (int k = col; k < i; ++k) v -= A[i * ld + k] * X[(size_t)k * M + col];
        v /= A[i * ld + i];
      }
      X[(size_t)i * M + col] = v;
    }
  }
  __syncthreads();
  // beta = L^-1 (m_q - c);  mean = c + (L^-1 k) . beta  (conditioned like L, not like K_ZZ)
  for (int i = tid; i < mp; i += nt) {
    double t = 0.0;
    if (i < M)
      for (int k = 0; k <= i; ++k) t += X[(size_t)i * M + k] * ((double)var_mean[(size_t)d * M + k] - c);
    alpha_out[(size_t)d * mp + i] = (float)t;
  }
  // fp32 outputs: Linv (lower) and L_q^T (upper), zero padded to mp
  for (int e = tid; e < mp * mp; e += nt) {
    const int r = e / mp, q = e % mp;
    float lv = 0.f, qv = 0.f;
    if (r < M && q < M) {
      if (q <= r) lv = (float)X[(size_t)r * M + q];
      if (q >= r) qv = chol_var[((size_t)d * M + q) * M + r];  // lqt[r][q] = L_q[q][r], q >= r
    }
    linv_out[((size_t)d * mp + r) * mp + q] = lv;
    lqt_out[((size_t)d * mp + r) * mp + q] = qv;
    linvT_out[((size_t)d * mp + q) * mp + r] = lv;   // transposed copies: lane-per-row kernels read [m][j]
    lq_out[((size_t)d * mp + q) * mp + r] = qv;
  }
}

// ---------------------------------------------------------------------------------------------------
// predictive mean / variance: grid (row tiles, D), 128 threads = 128 rows, one latent dim per CTA.
// Linv and L_q^T of that dim are staged in shared memory and read as warp-broadcast float4; the kernel
// row k[] of each thread lives in registers (MREG = padded M known at compile time) or shared memory.
// ---------------------------------------------------------------------------------------------------
// Stage one latent dimension's factors into shared memory (all 128 threads).
__device__ __forceinline__ void gp_stage_dim(int d, int MP, int tid, const float* __restrict__ zall,
                                             const float* __restrict__ linv_all, const float* __restrict__ lqt_all,
                                             const float* __restrict__ alpha_all, float* s_linv, float* s_lqt,
                                             float* s_z, float* s_alpha) {
  const float4* g1 = reinterpret_cast<const float4*>(linv_all + (size_t)d * MP * MP);
  const float4* g2 = reinterpret_cast<const float4*>(lqt_all + (size_t)d * MP * MP);
  for (int e = tid; e < MP * MP / 4; e += 128) {
    reinterpret_cast<float4*>(s_linv)[e] = __ldg(g1 + e);
    reinterpret_cast<float4*>(s_lqt)[e] = __ldg(g2 + e);
  }
  for (int e = tid; e < MP; e += 128) {
    s_z[e] = zall[(size_t)d * MP + e];
    s_alpha[e] = alpha_all[(size_t)d * MP + e];
  }
}

// One (row, dim) evaluation: mu = (Linv k).beta, vv = |Linv k|^2, ww = |L_q^T k|^2.
// Two independent accumulators per dot product halve the dependent-FMA chains (the kernel is latency bound
// when only a few rows are evaluated, e.g. the trigger's one row per rollout).
template <int MREG>
__device__ __forceinline__ void gp_row_eval(float xv, float s, float inv_ell, int MP, int tid, const float* s_linv,
                                            const float* s_lqt, const float* s_z, const float* s_alpha, float* s_k,
                                            float& mu, float& vv, float& ww) {
  mu = 0.f; vv = 0.f; ww = 0.f;
  if (MREG > 0) {
    float k[MREG > 0 ? MREG : 1];
#pragma unroll
    for (int m = 0; m < MREG; ++m) {
      const float t = (xv - s_z[m]) * inv_ell;
      k[m] = s * expf(-0.5f * t * t);   // padded z entries only ever meet zero factors
    }
#pragma unroll
    for (int j = 0; j < MREG; ++j) {
      float v0 = 0.f, v1 = 0.f, w0 = 0.f, w1 = 0.f;
      const int jm = (j / 4) * 4;
#pragma unroll
      for (int m = 0; m <= jm; m += 4) {
        const float4 l4 = *reinterpret_cast<const float4*>(s_linv + j * MREG + m);
        v0 = fmaf(l4.x, k[m], v0); v1 = fmaf(l4.y, k[m + 1], v1); v0 = fmaf(l4.z, k[m + 2], v0); v1 = fmaf(l4.w, k[m + 3], v1);
      }
#pragma unroll
      for (int m = jm; m < MREG; m += 4) {
        const float4 q4 = *reinterpret_cast<const float4*>(s_lqt + j * MREG + m);
        w0 = fmaf(q4.x, k[m], w0); w1 = fmaf(q4.y, k[m + 1], w1); w0 = fmaf(q4.z, k[m + 2], w0); w1 = fmaf(q4.w, k[m + 3], w1);
      }
      const float v = v0 + v1, w = w0 + w1;
      vv = fmaf(v, v, vv);
      ww = fmaf(w, w, ww);
      mu = fmaf(v, s_alpha[j], mu);
    }
  } else {
    for (int m = 0; m < MP; ++m) {
      const float t = (xv - s_z[m]) * inv_ell;
      s_k[m * 128 + tid] = s * expf(-0.5f * t * t);
    }
    for (int j = 0; j < MP; ++j) {
      float v0 = 0.f, v1 = 0.f, w0 = 0.f, w1 = 0.f;
      const int jm = (j / 4) * 4;
      for (int m = 0; m <= jm; m += 4) {
        const float4 l4 = *reinterpret_cast<const float4*>(s_linv + j * MP + m);
        v0 = fmaf(l4.x, s_k[m * 128 + tid], v0); v1 = fmaf(l4.y, s_k[(m + 1) * 128 + tid], v1);
        v0 = fmaf(l4.z, s_k[(m + 2) * 128 + tid], v0); v1 = fmaf(l4.w, s_k[(m + 3) * 128 + tid], v1);
      }
      for (int m = jm; m < MP; m += 4) {
        const float4 q4 = *reinterpret_cast<const float4*>(s_lqt + j * MP + m);
        w0 = fmaf(q4.x, s_k[m * 128 + tid], w0); w1 = fmaf(q4.y, s_k[(m + 1) * 128 + tid], w1);
        w0 = fmaf(q4.z, s_k[(m + 2) * 128 + tid], w0); w1 = fmaf(q4.w, s_k[(m + 3) * 128 + tid], w1);
      }
      const float v = v0 + v1, w = w0 + w1;
      vv = fmaf(v, v, vv);
      ww = fmaf(w, w, ww);
      mu = fmaf(v, s_alpha[j], mu);
    }
  }
}

template <int MREG>
__global__ void __launch_bounds__(128) gp_predict_kernel(int n_rows, int D, int mp, const float* __restrict__ x, int ldx,
                                                         const int32_t* __restrict__ row_index,
                                                         const float* __restrict__ zall,
                                                         const float* __restrict__ linv_all,
                                                         const float* __restrict__ lqt_all,
                                                         const float* __restrict__ alpha_all,
                                                         const float* __restrict__ hyp, float* __restrict__ mean,
                                                         int ldm, float* __restrict__ var, int ldv) {
  extern __shared__ __align__(16) float smf[];
  const int d = blockIdx.y, tid = threadIdx.x;
  const int MP = MREG > 0 ? MREG : mp;
  float* s_linv = smf;                  // [MP][MP]
  float* s_lqt = s_linv + MP * MP;      // [MP][MP]
  float* s_z = s_lqt + MP * MP;         // [MP]
  float* s_alpha = s_z + MP;            // [MP]
  float* s_k = s_alpha + MP;            // [MP][128] (generic path only)
  gp_stage_dim(d, MP, tid, zall, linv_all, lqt_all, alpha_all, s_linv, s_lqt, s_z, s_alpha);
  __syncthreads();
  const float ell = hyp[d * 4 + 0], s = hyp[d * 4 + 1], c = hyp[d * 4 + 2], noise = hyp[d * 4 + 3];
  const int i = blockIdx.x * 128 + tid;
  if (i >= n_rows) return;
  const int row = row_index ? row_index[i] : i;
  const float xv = __ldg(x + (size_t)row * ldx + d);
  float mu, vv, ww;
  gp_row_eval<MREG>(xv, s, 1.0f / ell, MP, tid, s_linv, s_lqt, s_z, s_alpha, s_k, mu, vv, ww);
  if (mean) mean[(size_t)i * ldm + d] = c + mu;
  if (var) var[(size_t)i * ldv + d] = (s - vv) + ww + noise;
}

// ---------------------------------------------------------------------------------------------------
// trigger finalize: one thread per rollout.  numpy-order float32 arithmetic (see oracle/trigger_ref.py).
// ---------------------------------------------------------------------------------------------------


// Fused trigger: grid (ceil(S/128), D), 256 threads.  Phase 1: two threads per (rollout, dim) task -- thread h=0
// accumulates |Linv k|^2, thread h=1 accumulates |L_q^T k|^2 -- with the factors of the CTA's dim staged in
// shared memory (warp-broadcast float4 reads) and k[] in registers; var_rows is written TRANSPOSED [D][S] so both
// this store and the finalize loads are coalesced.  Phase 2 (the last CTA to finish, atomic ticket): one thread
// per rollout loads its D variances in register batches of 32 (independent loads, one L2 round trip per batch),
// sums them in numpy's sequential fp32 order, updates the window, thresholds, decides, and appends fired
// rollouts to the compacted list used by the fused step kernel.  (History: thread-per-task without the v/w split
// was 10 us + 2 more launches; a warp-per-task variant was issue-bound at 33 us.)
template <int MREG>
__global__ void __launch_bounds__(256) gp_trigger_kernel(int S, int D, int mp, const float* __restrict__ x, int ldx,
                                                         const int32_t* __restrict__ stat_rows,
                                                         const float* __restrict__ zall,
                                                         const float* __restrict__ linv_all,
                                                         const float* __restrict__ lqt_all,
                                                         const float* __restrict__ hyp, float* var_rows,
                                                         unsigned int* ticket, float* window, int W, int32_t* count,
                                                         int warmup, float factor, float* value, float* thr,
                                                         uint8_t* mask, int* trig_list, int* trig_count) {
  extern __shared__ __align__(16) float smf[];
  __shared__ int s_last, s_cnt;
  const int d = blockIdx.y, tid = threadIdx.x;
  const int MP = MREG > 0 ? MREG : mp;
  float* s_linv = smf;
  float* s_lqt = s_linv + MP * MP;
  float* s_z = s_lqt + MP * MP;
  const int half = tid >> 7;                 // 0: v = Linv k,  1: w = L_q^T k
  const int li = tid & 127;
  const int i = blockIdx.x * 128 + li;
  float xv = 0.f;
  if (i < S) xv = __ldg(x + (size_t)stat_rows[i] * ldx + d);     // issue the (possibly DRAM) load before staging
  {
    const float4* g1 = reinterpret_cast<const float4*>(linv_all + (size_t)d * MP * MP);
    const float4* g2 = reinterpret_cast<const float4*>(lqt_all + (size_t)d * MP * MP);
    for (int e = tid; e < MP * MP / 4; e += 256) {
      reinterpret_cast<float4*>(s_linv)[e] = __ldg(g1 + e);
      reinterpret_cast<float4*>(s_lqt)[e] = __ldg(g2 + e);
    }
    for (int e = tid; e < MP; e += 256) s_z[e] = zall[(size_t)d * MP + e];
  }
  __syncthreads();
  const float ell = hyp[d * 4 + 0], sc = hyp[d * 4 + 1], noise = hyp[d * 4 + 3];
  float part = 0.f;                           // |v|^2 (half 0) or |w|^2 (half 1)
  if (i < S) part = gp_trig_partial<MREG>(xv, sc, 1.0f / ell, MP, half == 0 ? s_linv : s_lqt, s_z, half != 0);
  // combine the two halves through shared memory (smf is dead after the barrier)
  __syncthreads();
  if (half == 1) smf[li] = part;
  __syncthreads();
  if (half == 0 && i < S) var_rows[(size_t)d * S + i] = (sc - part) + smf[li] + noise;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x * gridDim.y - 1 ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid == 0) s_cnt = count[0];               // read once, before thread 0 may advance it below
  __syncthreads();
  const int cnt = s_cnt;
  for (int s = tid; s < S; s += 256)
    gp_trig_finalize_rollout(s, S, D, var_rows, window, W, cnt, warmup, factor, value, thr, mask, trig_list, trig_count);
  if (tid == 0) {
    *ticket = 0;                                   // ready for the next launch
    if (warmup && cnt < W) count[0] = cnt + 1;
  }
}

// grid (S, D): one CTA per (rollout, dim); CTAs of unmasked rollouts exit.
__global__ void __launch_bounds__(RS_THREADS) gp_rsample_kernel(int S, int N, int D, int mp, const float* __restrict__ x, int ldx,
                                                         const float* __restrict__ eps,
                                                         const uint8_t* __restrict__ mask,
                                                         const float* __restrict__ zall,
                                                         const float* __restrict__ linv_all,
                                                         const float* __restrict__ lqt_all,
                                                         const float* __restrict__ alpha_all,
                                                         const float* __restrict__ hyp, float* __restrict__ out, int ldo) {
  if (mask != nullptr && mask[blockIdx.x] == 0) return;
  extern __shared__ __align__(16) float smf[];
  gp_rsample_body<RS_THREADS>(smf, (int)threadIdx.x, [] { __syncthreads(); }, blockIdx.x, blockIdx.y, N, D, mp, x, ldx, eps, zall, linv_all,
                  lqt_all, alpha_all, hyp, out, ldo);
}

// grid (D, splits): CTA (d, j) scans the mask and solves dim d of every masked rollout s with s % splits == j, so a step
// in which nothing fired costs D*splits near-empty CTAs instead of S*D launches with 200 KB of shared memory each.
// The mask CONTENTS decide (an earlier version replayed the fired list of the last trigger call whenever the same
// mask pointer came back; a caller that edits or zeroes that buffer, or an allocator that hands the address to another
// mask, then resampled a stale list).
__global__ void __launch_bounds__(RS_THREADS) gp_rsample_scan_kernel(int S, int N, int D, int mp, const float* __restrict__ x, int ldx,
                                                              const float* __restrict__ eps,
                                                              const uint8_t* __restrict__ mask,
                                                              const float* __restrict__ zall,
                                                              const float* __restrict__ linv_all,
                                                              const float* __restrict__ lqt_all,
                                                              const float* __restrict__ alpha_all,
                                                              const float* __restrict__ hyp, float* __restrict__ out,
                                                              int ldo) {
  extern __shared__ __align__(16) float smf[];
  for (int s = blockIdx.y; s < S; s += gridDim.y) {
    if (mask[s] == 0) continue;                      // CTA-uniform
    gp_rsample_body<RS_THREADS>(smf, (int)threadIdx.x, [] { __syncthreads(); }, s, blockIdx.x, N, D, mp, x, ldx, eps, zall,
                    linv_all, lqt_all, alpha_all, hyp, out, ldo);
    __syncthreads();   // shared memory is reused by the next rollout
  }
}

// ---------------------------------------------------------------------------------------------------
// host launchers (called from capi.cu)
// ---------------------------------------------------------------------------------------------------
int gp_prepare_launch(dvg_gp_s* h, const float* inducing, const float* var_mean, const float* chol_var,
                      const float* mean_const, const float* raw_os, const float* raw_ls, const float* raw_noise,
                      cudaStream_t stream) {
  const int D = h->dims.num_dims, M = h->dims.num_inducing;
  const size_t smem = sizeof(double) * ((size_t)M * (M + 1) + 2 * M);
  DVG_REQUIRE(smem <= 227 * 1024, "num_inducing=%d too large for the on-device fp64 factorisation (max 160)", M);
  static bool configured = false;
  if (!configured) {
    DVG_CUDA(cudaFuncSetAttribute(gp_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  gp_prepare_kernel<<<D, 128, smem, stream>>>(D, M, h->mp, (double)h->dims.jitter, (double)h->dims.noise_lower_bound,
                                              inducing, var_mean, chol_var, mean_const, raw_os, raw_ls, raw_noise,
                                              h->z, h->linv, h->lqt, h->linvT, h->lq, h->alpha, h->hyp, h->work);
  DVG_LAUNCH_CHECK();
  return DVG_OK;
}

int gp_predict_launch(dvg_gp_s* h, int n_rows, const float* x, int ldx, const int32_t* row_index, float* mean, int ldm,
                      float* var, int ldv, cudaStream_t stream) {
  if (n_rows <= 0) return DVG_OK;
  const int D = h->dims.num_dims, mp = h->mp;
  dim3 grid(ceil_div(n_rows, 128), D);
  static bool configured = false;
  if (!configured) {
    DVG_CUDA(cudaFuncSetAttribute(gp_predict_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  if (mp == 40) {
    const size_t smem = sizeof(float) * (2 * 40 * 40 + 2 * 40);
    gp_predict_kernel<40><<<grid, 128, smem, stream>>>(n_rows, D, mp, x, ldx, row_index, h->z, h->linv, h->lqt, h->alpha,
                                                       h->hyp, mean, ldm, var, ldv);
  } else {
    const size_t smem = sizeof(float) * ((size_t)2 * mp * mp + 2 * mp + (size_t)mp * 128);
    DVG_REQUIRE(smem <= 227 * 1024, "num_inducing=%d too large for the shared-memory predictive kernel", mp);
    gp_predict_kernel<0><<<grid, 128, smem, stream>>>(n_rows, D, mp, x, ldx, row_index, h->z, h->linv, h->lqt, h->alpha,
                                                      h->hyp, mean, ldm, var, ldv);
  }
  DVG_LAUNCH_CHECK();
  return DVG_OK;
}

int gp_trigger_launch(dvg_gp_s* h, int S, const float* x, int ldx, const int32_t* stat_rows, float* window, int W,
                      int32_t* count, int warmup, float factor, float* value, float* thr, uint8_t* mask,
                      cudaStream_t stream) {
  DVG_REQUIRE(W >= 1 && W <= MAX_WINDOW, "window_len must be in [1,%d]", MAX_WINDOW);
  DVG_REQUIRE(S <= h->var_rows_cap, "n_rollouts=%d exceeds the reserved trigger scratch (%d)", S, h->var_rows_cap);
  const int D = h->dims.num_dims, mp = h->mp;
  // the fired list is rebuilt by every call: reset its counter first (tiny memset node)
  DVG_CUDA(cudaMemsetAsync(h->trig_count, 0, sizeof(int), stream));
  dim3 grid(ceil_div(S, 128), D);
  size_t smem = sizeof(float) * ((size_t)2 * mp * mp + mp);
  if (smem < 128 * sizeof(float)) smem = 128 * sizeof(float);
  DVG_REQUIRE(smem <= 226 * 1024, "num_inducing=%d too large for the shared-memory trigger kernel", mp);
  static bool configured = false;
  if (!configured) {
    DVG_CUDA(cudaFuncSetAttribute(gp_trigger_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    configured = true;
  }
  if (mp == 40)
    gp_trigger_kernel<40><<<grid, 256, smem, stream>>>(S, D, mp, x, ldx, stat_rows, h->z, h->linv, h->lqt, h->hyp,
                                                       h->var_rows, h->ticket, window, W, count, warmup, factor, value,
                                                       thr, mask, h->trig_list, h->trig_count);
  else
    gp_trigger_kernel<0><<<grid, 256, smem, stream>>>(S, D, mp, x, ldx, stat_rows, h->z, h->linv, h->lqt, h->hyp,
                                                      h->var_rows, h->ticket, window, W, count, warmup, factor, value,
                                                      thr, mask, h->trig_list, h->trig_count);
  DVG_LAUNCH_CHECK();
  return DVG_OK;
}

int gp_rsample_launch(dvg_gp_s* h, int S, int N, const float* x, int ldx, const float* eps, const uint8_t* mask,
                      float* out, int ldo, cudaStream_t stream) {
  const int D = h->dims.num_dims, mp = h->mp;
  const size_t smem = sizeof(float) * gp_rsample_smem_floats(N, mp);
  DVG_REQUIRE(smem <= 227 * 1024, "rsample needs %zu B of shared memory for N=%d, M=%d (max 227 KB)", smem, N, mp);
  static bool configured = false;
  if (!configured) {
    DVG_CUDA(cudaFuncSetAttribute(gp_rsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DVG_CUDA(cudaFuncSetAttribute(gp_rsample_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  if (mask != nullptr && S > 8) {
    gp_rsample_scan_kernel<<<dim3(D, 4), RS_THREADS, smem, stream>>>(S, N, D, mp, x, ldx, eps, mask, h->z, h->linv, h->lqt,
                                                               h->alpha, h->hyp, out, ldo);
  } else {
    gp_rsample_kernel<<<dim3(S, D), RS_THREADS, smem, stream>>>(S, N, D, mp, x, ldx, eps, mask, h->z, h->linv, h->lqt, h->alpha,
                                                         h->hyp, out, ldo);
  }
  DVG_LAUNCH_CHECK();
  return DVG_OK;
}

}  // namespace dvg
