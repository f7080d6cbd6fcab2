// GP predictive mean / variance and variance trigger for LARGE inducing sets (M > 128: the per-dimension factors no
// longer fit shared memory; BASELINE configs[4] sweeps M up to 4096).  Same math as gp.cu
// (models/gp_models.py:10-24 + gpytorch WhitenedVariationalStrategy eval branch, SURVEY 8c eqs. 1-6):
//   k = k(Z, x) [M],  v = Linv k,  w = L_q^T k,  mean = c + v . beta,  var = s - |v|^2 + |w|^2 + noise
// but organised as a tiled FP32 GEMM per latent dimension: V = Linv (M x M, lower) * K (M x N), W = L_q^T (upper) * K.
// A CTA owns a 64 (matrix rows j) x 64 (points n) tile of V and W, walks the m tiles that are non-zero for its row
// block (m <= j for Linv, m >= j for L_q^T), builds the K tile on the fly (exp per element, never stored in HBM) and
// reduces its rows to three partial sums per point (|v|^2, |w|^2, v . beta).  A second tiny kernel adds the row-block
// partials in a fixed order (deterministic) and writes mean / variance with caller-chosen strides, so the trigger can
// ask for the transposed [D][S] scratch directly.
// Bound: FP32 FMA pipe (2 M^2 FMA per (point, dim) against 12 B of I/O); factors stream from L2/HBM once per 64 points.
#include "gp_trigger.cuh"
#include "internal.cuh"

namespace dvg {

constexpr int GB_T = 64;        // tile edge
constexpr int GB_LD = 68;       // smem row stride (floats): multiple of 4 for float4 reads, != 64 against conflicts

__global__ void __launch_bounds__(256) gp_big_partial_kernel(int n_rows, int Mp, const float* __restrict__ x, int ldx,
                                                             const int32_t* __restrict__ row_index,
                                                             const float* __restrict__ zall,
                                                             const float* __restrict__ linv_all,
                                                             const float* __restrict__ lqt_all,
                                                             const float* __restrict__ beta_all,
                                                             const float* __restrict__ hyp, float* __restrict__ partial,
                                                             int n_pad) {
  extern __shared__ __align__(16) float smf[];
  float* Ks = smf;                       // [64 m][GB_LD]  K tile, m-major
  float* Lt = Ks + GB_T * GB_LD;         // [64 m][GB_LD]  Linv tile transposed: Lt[m][j]
  float* Qt = Lt + GB_T * GB_LD;         // [64 m][GB_LD]  L_q^T tile transposed
  float* xs = Qt + GB_T * GB_LD;         // [64]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int d = blockIdx.z, jb = blockIdx.y, n0 = blockIdx.x * GB_T, JB = gridDim.y;
  const int MT = Mp / GB_T, j0 = jb * GB_T;
  const float ell = hyp[d * 4 + 0], sc = hyp[d * 4 + 1];
  const float inv_ell = 1.0f / ell;
  const float* linv = linv_all + (size_t)d * Mp * Mp;
  const float* lqt = lqt_all + (size_t)d * Mp * Mp;
  const float* z = zall + (size_t)d * Mp;
  if (tid < GB_T) {
    const int n = n0 + tid;
    float v = 0.f;
    if (n < n_rows) {
      const int row = row_index ? row_index[n] : n;
      v = __ldg(x + (size_t)row * ldx + d);
    }
    xs[tid] = v;
  }
  float accV[4][4], accW[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) { accV[a][b] = 0.f; accW[a][b] = 0.f; }
  __syncthreads();
  for (int mt = 0; mt < MT; ++mt) {
    const int m0 = mt * GB_T;
    const bool doV = mt <= jb, doW = mt >= jb;
    // factor tiles: [64 j][64 m] row-major in global (float4 along m) -> transposed in shared memory
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + i * 256;
      const int jj = e >> 4, m4 = (e & 15) * 4;
      if (doV) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(linv + (size_t)(j0 + jj) * Mp + m0 + m4));
        Lt[(m4 + 0) * GB_LD + jj] = v.x; Lt[(m4 + 1) * GB_LD + jj] = v.y;
        Lt[(m4 + 2) * GB_LD + jj] = v.z; Lt[(m4 + 3) * GB_LD + jj] = v.w;
      }
      if (doW) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(lqt + (size_t)(j0 + jj) * Mp + m0 + m4));
        Qt[(m4 + 0) * GB_LD + jj] = q.x; Qt[(m4 + 1) * GB_LD + jj] = q.y;
        Qt[(m4 + 2) * GB_LD + jj] = q.z; Qt[(m4 + 3) * GB_LD + jj] = q.w;
      }
    }
    // K tile: k(z_m, x_n) = s exp(-0.5 ((x - z) / ell)^2)   (padded z entries only ever meet zero factors)
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int e = tid + i * 256;
      const int mm = e >> 6, n = e & 63;
      const float t = (xs[n] - __ldg(z + m0 + mm)) * inv_ell;
      Ks[mm * GB_LD + n] = sc * expf(-0.5f * t * t);
    }
    __syncthreads();
    if (doV && doW) {
#pragma unroll 8
      for (int kk = 0; kk < GB_T; ++kk) {
        const float4 b4 = *reinterpret_cast<const float4*>(Ks + kk * GB_LD + tx * 4);
        const float4 a4 = *reinterpret_cast<const float4*>(Lt + kk * GB_LD + ty * 4);
        const float4 q4 = *reinterpret_cast<const float4*>(Qt + kk * GB_LD + ty * 4);
        const float bv[4] = {b4.x, b4.y, b4.z, b4.w}, av[4] = {a4.x, a4.y, a4.z, a4.w}, qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            accV[a][b] = fmaf(av[a], bv[b], accV[a][b]);
            accW[a][b] = fmaf(qv[a], bv[b], accW[a][b]);
          }
      }
    } else {
      const float* Ft = doV ? Lt : Qt;
      float (&acc)[4][4] = doV ? accV : accW;
#pragma unroll 8
      for (int kk = 0; kk < GB_T; ++kk) {
        const float4 b4 = *reinterpret_cast<const float4*>(Ks + kk * GB_LD + tx * 4);
        const float4 a4 = *reinterpret_cast<const float4*>(Ft + kk * GB_LD + ty * 4);
        const float bv[4] = {b4.x, b4.y, b4.z, b4.w}, av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
      }
    }
    __syncthreads();
  }
  // row reduction of this 64-row block: per point |v|^2, |w|^2, v . beta  (fixed order: rows within a thread, then ty)
  float* red = smf;                      // [16 ty][64 n][3] aliases the tiles (all reads of them are done)
  float bt[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) bt[a] = __ldg(beta_all + (size_t)d * Mp + j0 + ty * 4 + a);
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    float pv = 0.f, pw = 0.f, pm = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      pv = fmaf(accV[a][b], accV[a][b], pv);
      pw = fmaf(accW[a][b], accW[a][b], pw);
      pm = fmaf(accV[a][b], bt[a], pm);
    }
    float* r = red + ((size_t)ty * GB_T + tx * 4 + b) * 3;
    r[0] = pv; r[1] = pw; r[2] = pm;
  }
  __syncthreads();
  if (tid < 3 * GB_T) {
    const int n = tid / 3, q = tid % 3;
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < 16; ++t) s += red[((size_t)t * GB_T + n) * 3 + q];
    if (n0 + n < n_pad) partial[(((size_t)d * JB + jb) * n_pad + n0 + n) * 3 + q] = s;
  }
}

// mean / var of point n, dim d from the JB row-block partials (summed in row-block order).
__global__ void __launch_bounds__(128) gp_big_finalize_kernel(int n_rows, int D, int JB, int n_pad,
                                                              const float* __restrict__ partial,
                                                              const float* __restrict__ hyp, float* __restrict__ mean,
                                                              long long mean_sn, long long mean_sd,
                                                              float* __restrict__ var, long long var_sn, long long var_sd) {
  const int n = blockIdx.x * 128 + threadIdx.x, d = blockIdx.y;
  if (n >= n_rows) return;
  float pv = 0.f, pw = 0.f, pm = 0.f;
  for (int jb = 0; jb < JB; ++jb) {
    const float* r = partial + (((size_t)d * JB + jb) * n_pad + n) * 3;
    pv += r[0]; pw += r[1]; pm += r[2];
  }
  const float sc = hyp[d * 4 + 1], c = hyp[d * 4 + 2], noise = hyp[d * 4 + 3];
  if (mean) mean[(size_t)n * mean_sn + (size_t)d * mean_sd] = c + pm;
  if (var) var[(size_t)n * var_sn + (size_t)d * var_sd] = (sc - pv) + pw + noise;
}

// Window / threshold / decision of every rollout from var_rows [D][S] (generate_frames.py:230-231, 283-289): one CTA.
__global__ void __launch_bounds__(1024) gp_trigger_finalize_kernel(int S, int D, const float* var_rows, float* window, int W,
                                                                   int32_t* count, int warmup, float factor, float* value,
                                                                   float* thr, uint8_t* mask, int* trig_list,
                                                                   int* trig_count) {
  const int cnt = count[0];
  __syncthreads();
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    if (W <= 16)
      gp_trig_finalize_rollout16(s, S, D, var_rows, window, W, cnt, warmup, factor, value, thr, mask, trig_list, trig_count);
    else
      gp_trig_finalize_rollout(s, S, D, var_rows, window, W, cnt, warmup, factor, value, thr, mask, trig_list, trig_count);
  }
  __syncthreads();
  if (threadIdx.x == 0 && warmup && cnt < W) count[0] = cnt + 1;
}

// [D][M][M] row-major factors -> zero-padded [D][Mp][Mp]; transpose = 1 writes dst[d][c][r] = src[d][r][c].
__global__ void gp_big_pad_kernel(int M, int Mp, const float* __restrict__ src, float* __restrict__ dst, int transpose) {
  const int d = blockIdx.z;
  const int r = blockIdx.y * 16 + threadIdx.y, c = blockIdx.x * 16 + threadIdx.x;
  if (r >= Mp || c >= Mp) return;
  float v = 0.f;
  if (r < M && c < M) v = transpose ? src[((size_t)d * M + c) * M + r] : src[((size_t)d * M + r) * M + c];
  dst[((size_t)d * Mp + r) * Mp + c] = v;
}
__global__ void gp_big_pad_vec_kernel(int M, int Mp, const float* __restrict__ src, float* __restrict__ dst) {
  const int d = blockIdx.y, m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m < Mp) dst[(size_t)d * Mp + m] = m < M ? src[(size_t)d * M + m] : 0.f;
}

int gp_big_load_factors(dvg_gp_s* h, const float* inducing, const float* linv, const float* lq, const float* beta,
                        const float* hyp, cudaStream_t stream) {
  const int D = h->dims.num_dims, M = h->dims.num_inducing, Mp = h->mp;
  dim3 blk(16, 16), grd(ceil_div(Mp, 16), ceil_div(Mp, 16), D);
  gp_big_pad_kernel<<<grd, blk, 0, stream>>>(M, Mp, linv, h->linv, 0);
  DVG_LAUNCH_CHECK();
  gp_big_pad_kernel<<<grd, blk, 0, stream>>>(M, Mp, lq, h->lqt, 1);      // lqt[r][c] = L_q[c][r]
  DVG_LAUNCH_CHECK();
  dim3 vg(ceil_div(Mp, 128), D);
  gp_big_pad_vec_kernel<<<vg, 128, 0, stream>>>(M, Mp, inducing, h->z);
  DVG_LAUNCH_CHECK();
  gp_big_pad_vec_kernel<<<vg, 128, 0, stream>>>(M, Mp, beta, h->alpha);
  DVG_LAUNCH_CHECK();
  DVG_CUDA(cudaMemcpyAsync(h->hyp, hyp, sizeof(float) * D * 4, cudaMemcpyDeviceToDevice, stream));
  return gp_tc_pack(h, stream);          // bf16 hi/lo k-block images for the tensor-core path (gp_tc.cu)
}

static int gp_big_reserve_partial(dvg_gp_s* h, size_t floats, cudaStream_t stream) {
  if (floats <= h->partial_cap) return DVG_OK;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &st);
  DVG_REQUIRE(st == cudaStreamCaptureStatusNone, "GP scratch must be grown before stream capture (call once eagerly)");
  DVG_CUDA(cudaDeviceSynchronize());
  if (h->partial) h->retired.push_back(h->partial);     // CUDA graphs captured earlier may still reference it
  h->partial = nullptr; h->partial_cap = 0;
  DVG_CUDA(cudaMalloc(&h->partial, sizeof(float) * floats));
  h->partial_cap = floats;
  return DVG_OK;
}

// mean[n * mean_sn + d * mean_sd], var likewise (either may be null)
int gp_big_predict_launch(dvg_gp_s* h, int n_rows, const float* x, int ldx, const int32_t* row_index, float* mean,
                          long long mean_sn, long long mean_sd, float* var, long long var_sn, long long var_sd,
                          cudaStream_t stream) {
  if (n_rows <= 0) return DVG_OK;
  const int D = h->dims.num_dims, Mp = h->mp, JB = Mp / GB_T;
  static bool configured = false;
  const size_t smem = sizeof(float) * (3 * GB_T * GB_LD + GB_T);
  if (!configured) {
    DVG_CUDA(cudaFuncSetAttribute(gp_big_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  // bound the partial-sum scratch: process the points in chunks
  const int chunk_max = 4096;
  const bool tc = gp_tc_enabled() && h->tc_img_v != nullptr;
  for (int c0 = 0; tc && c0 < n_rows; c0 += chunk_max) {
    // tensor-core path: 128-point x 256-row tiles, same scratch layout and finalize kernel (JB = row tiles of 256)
    const int nc = n_rows - c0 < chunk_max ? n_rows - c0 : chunk_max;
    const int n_pad = ceil_div(nc, 128) * 128;
    int rc = gp_big_reserve_partial(h, (size_t)D * h->tc_JT * n_pad * 3, stream);
    if (rc) return rc;
    const float* xc = row_index ? x : x + (size_t)c0 * ldx;
    const int32_t* ric = row_index ? row_index + c0 : nullptr;
    if ((rc = gp_tc_partial_launch(h, nc, n_pad, xc, ldx, ric, mean != nullptr ? 1 : 0, stream))) return rc;
    dim3 g2(ceil_div(nc, 128), D);
    gp_big_finalize_kernel<<<g2, 128, 0, stream>>>(nc, D, h->tc_JT, n_pad, h->partial, h->hyp,
                                                   mean ? mean + (size_t)c0 * mean_sn : nullptr, mean_sn, mean_sd,
                                                   var ? var + (size_t)c0 * var_sn : nullptr, var_sn, var_sd);
    DVG_LAUNCH_CHECK();
  }
  if (tc) return DVG_OK;
  for (int c0 = 0; c0 < n_rows; c0 += chunk_max) {
    const int nc = n_rows - c0 < chunk_max ? n_rows - c0 : chunk_max;
    const int n_pad = ceil_div(nc, GB_T) * GB_T;
    int rc = gp_big_reserve_partial(h, (size_t)D * JB * n_pad * 3, stream);
    if (rc) return rc;
    const float* xc = row_index ? x : x + (size_t)c0 * ldx;
    const int32_t* ric = row_index ? row_index + c0 : nullptr;
    dim3 grid(n_pad / GB_T, JB, D);
    gp_big_partial_kernel<<<grid, 256, smem, stream>>>(nc, Mp, xc, ldx, ric, h->z, h->linv, h->lqt, h->alpha, h->hyp,
                                                       h->partial, n_pad);
    DVG_LAUNCH_CHECK();
    dim3 g2(ceil_div(nc, 128), D);
    gp_big_finalize_kernel<<<g2, 128, 0, stream>>>(nc, D, JB, n_pad, h->partial, h->hyp,
                                                   mean ? mean + (size_t)c0 * mean_sn : nullptr, mean_sn, mean_sd,
                                                   var ? var + (size_t)c0 * var_sn : nullptr, var_sn, var_sd);
    DVG_LAUNCH_CHECK();
  }
  return DVG_OK;
}

// ---------------------------------------------------------------------------------------------------
// .rsample() for large inducing sets (generate_frames.py:171,292 at BASELINE configs[4] sizes): one CTA per
// (rollout, latent dim), masked rollouts only.  The [N,N] predictive covariance of the rollout's N <= 128 points
//   Sigma_y = K_xx + noise I + sum_jb ( W_jb^T W_jb - V_jb^T V_jb ),   V = Linv K_zx,  W = L_q^T K_zx  (row blocks jb of 64)
// is accumulated row block by row block: the 64 x N tiles of V and W are produced exactly like gp_big_partial_kernel
// does (K tile rebuilt on the fly, only the non-zero triangular m tiles visited), parked in shared memory, and folded
// into register-resident 4 x 4 blocks of the lower triangle of the Gram difference; mean = c + V^T beta on the way.
// Then an in-place Cholesky in shared memory and  out = mean + L eps.  Everything in a fixed order: deterministic.
// Work per CTA: M^2 N FMAs (M = 4096, N = 50: 0.84 G) -- a rare event (fired rollouts only), D CTAs per fired rollout.
// ---------------------------------------------------------------------------------------------------
constexpr int GR_NMAX = 128;
constexpr int GR_LDV = 132;      // row stride of the parked V / W tiles and of Sigma (floats)

__global__ void __launch_bounds__(256) gp_big_rsample_kernel(int S, int N, int D, int Mp, const float* __restrict__ x, int ldx,
                                                             const float* __restrict__ eps, const uint8_t* __restrict__ mask,
                                                             const float* __restrict__ zall, const float* __restrict__ linv_all,
                                                             const float* __restrict__ lqt_all, const float* __restrict__ beta_all,
                                                             const float* __restrict__ hyp, float* __restrict__ out, int ldo) {
  const int sidx = blockIdx.x, d = blockIdx.y;
  if (mask != nullptr && mask[sidx] == 0) return;
  extern __shared__ __align__(16) float smf[];
  float* Ks = smf;                           // [64 m][GB_LD]
  float* Lt = Ks + GB_T * GB_LD;             // [64 m][GB_LD]  Linv tile transposed
  float* Qt = Lt + GB_T * GB_LD;             // [64 m][GB_LD]  L_q^T tile transposed
  float* Vs = Qt + GB_T * GB_LD;             // [64 j][GR_LDV] V row block, all N points
  float* Ws = Vs + GB_T * GR_LDV;            // [64 j][GR_LDV]
  float* Sg = Ws + GB_T * GR_LDV;            // [GR_NMAX][GR_LDV] Sigma_y (lower triangle) / Cholesky factor
  float* xs = Sg + GR_NMAX * GR_LDV;         // [GR_NMAX] the rollout's latents of this dim
  float* mn = xs + GR_NMAX;                  // [GR_NMAX] predictive mean
  float* bt = mn + GR_NMAX;                  // [64] beta of the current row block
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int MT = Mp / GB_T, NT = (N + GB_T - 1) / GB_T;
  const float ell = hyp[d * 4 + 0], sc = hyp[d * 4 + 1], cst = hyp[d * 4 + 2], noise = hyp[d * 4 + 3];
  const float inv_ell = 1.0f / ell;
  const float* linv = linv_all + (size_t)d * Mp * Mp;
  const float* lqt = lqt_all + (size_t)d * Mp * Mp;
  const float* z = zall + (size_t)d * Mp;
  if (tid < GR_NMAX) {
    xs[tid] = tid < N ? __ldg(x + (size_t)(sidx * N + tid) * ldx + d) : 0.f;
    mn[tid] = 0.f;
  }
  // this thread's 4 x 4 blocks of the lower triangle of G = sum_jb (W^T W - V^T V): block (bi, bj), bj <= bi, NB = ceil(N/4)
  const int NB = (N + 3) / 4;
  const int n_blocks = NB * (NB + 1) / 2;
  constexpr int MAXB = 3;                    // 528 blocks at N = 128 over 256 threads
  float g[MAXB][4][4];
  int gbi[MAXB], gbj[MAXB];
#pragma unroll
  for (int q = 0; q < MAXB; ++q) {
    const int b = tid + q * 256;
    int bi = 0;
    while ((bi + 1) * (bi + 2) / 2 <= b) ++bi;
    gbi[q] = bi; gbj[q] = b - bi * (bi + 1) / 2;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c2 = 0; c2 < 4; ++c2) g[q][a][c2] = 0.f;
  }
  __syncthreads();
  for (int jb = 0; jb < MT; ++jb) {
    const int j0 = jb * GB_T;
    if (tid < GB_T) bt[tid] = __ldg(beta_all + (size_t)d * Mp + j0 + tid);
    for (int nt = 0; nt < NT; ++nt) {
      float accV[4][4], accW[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { accV[a][b] = 0.f; accW[a][b] = 0.f; }
      for (int mt = 0; mt < MT; ++mt) {
        const int m0 = mt * GB_T;
        const bool doV = mt <= jb, doW = mt >= jb;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int e = tid + i * 256;
          const int jj = e >> 4, m4 = (e & 15) * 4;
          if (doV) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(linv + (size_t)(j0 + jj) * Mp + m0 + m4));
            Lt[(m4 + 0) * GB_LD + jj] = v.x; Lt[(m4 + 1) * GB_LD + jj] = v.y;
            Lt[(m4 + 2) * GB_LD + jj] = v.z; Lt[(m4 + 3) * GB_LD + jj] = v.w;
          }
          if (doW) {
            const float4 q4 = __ldg(reinterpret_cast<const float4*>(lqt + (size_t)(j0 + jj) * Mp + m0 + m4));
            Qt[(m4 + 0) * GB_LD + jj] = q4.x; Qt[(m4 + 1) * GB_LD + jj] = q4.y;
            Qt[(m4 + 2) * GB_LD + jj] = q4.z; Qt[(m4 + 3) * GB_LD + jj] = q4.w;
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int e = tid + i * 256;
          const int mm = e >> 6, n = e & 63;
          const float t = (xs[nt * GB_T + n] - __ldg(z + m0 + mm)) * inv_ell;
          Ks[mm * GB_LD + n] = sc * expf(-0.5f * t * t);
        }
        __syncthreads();
#pragma unroll 8
        for (int kk = 0; kk < GB_T; ++kk) {
          const float4 b4 = *reinterpret_cast<const float4*>(Ks + kk * GB_LD + tx * 4);
          const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
          if (doV) {
            const float4 a4 = *reinterpret_cast<const float4*>(Lt + kk * GB_LD + ty * 4);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b) accV[a][b] = fmaf(av[a], bv[b], accV[a][b]);
          }
          if (doW) {
            const float4 q4 = *reinterpret_cast<const float4*>(Qt + kk * GB_LD + ty * 4);
            const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int b = 0; b < 4; ++b) accW[a][b] = fmaf(qv[a], bv[b], accW[a][b]);
          }
        }
        __syncthreads();
      }
      // park the 64 x 64 tiles: rows ty*4 + a, points nt*64 + tx*4 + b
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        *reinterpret_cast<float4*>(Vs + (ty * 4 + a) * GR_LDV + nt * GB_T + tx * 4) = make_float4(accV[a][0], accV[a][1], accV[a][2], accV[a][3]);
        *reinterpret_cast<float4*>(Ws + (ty * 4 + a) * GR_LDV + nt * GB_T + tx * 4) = make_float4(accW[a][0], accW[a][1], accW[a][2], accW[a][3]);
      }
    }
    __syncthreads();
    // mean += V_jb^T beta_jb  (rows in order: deterministic)
    if (tid < N) {
      float m = mn[tid];
      for (int r = 0; r < GB_T; ++r) m = fmaf(Vs[r * GR_LDV + tid], bt[r], m);
      mn[tid] = m;
    }
    // G += W^T W - V^T V on this thread's blocks
#pragma unroll
    for (int q = 0; q < MAXB; ++q) {
      if (tid + q * 256 >= n_blocks) continue;
      const int i0 = gbi[q] * 4, c0 = gbj[q] * 4;
      for (int r = 0; r < GB_T; ++r) {
        const float4 vi = *reinterpret_cast<const float4*>(Vs + r * GR_LDV + i0), vj = *reinterpret_cast<const float4*>(Vs + r * GR_LDV + c0);
        const float4 wi = *reinterpret_cast<const float4*>(Ws + r * GR_LDV + i0), wj = *reinterpret_cast<const float4*>(Ws + r * GR_LDV + c0);
        const float via[4] = {vi.x, vi.y, vi.z, vi.w}, vja[4] = {vj.x, vj.y, vj.z, vj.w};
        const float wia[4] = {wi.x, wi.y, wi.z, wi.w}, wja[4] = {wj.x, wj.y, wj.z, wj.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c2 = 0; c2 < 4; ++c2) g[q][a][c2] = fmaf(wia[a], wja[c2], fmaf(-via[a], vja[c2], g[q][a][c2]));
      }
    }
    __syncthreads();
  }
  // Sigma_y (lower triangle) = K_xx + noise I + G
#pragma unroll
  for (int q = 0; q < MAXB; ++q) {
    if (tid + q * 256 >= n_blocks) continue;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c2 = 0; c2 < 4; ++c2) {
        const int i = gbi[q] * 4 + a, j = gbj[q] * 4 + c2;
        if (i < N && j <= i) {
          const float t = (xs[i] - xs[j]) * inv_ell;
          Sg[i * GR_LDV + j] = (i == j ? sc + noise : sc * expf(-0.5f * t * t)) + g[q][a][c2];
        }
      }
  }
  __syncthreads();
  // right-looking Cholesky, column by column (N <= 128)
  for (int k = 0; k < N; ++k) {
    if (tid == 0) Sg[k * GR_LDV + k] = sqrtf(Sg[k * GR_LDV + k]);
    __syncthreads();
    const float inv = 1.0f / Sg[k * GR_LDV + k];
    for (int i = k + 1 + tid; i < N; i += 256) Sg[i * GR_LDV + k] *= inv;
    __syncthreads();
    const int rem = N - k - 1;
    for (int e = tid; e < rem * rem; e += 256) {
      const int i = k + 1 + e / rem, j = k + 1 + e % rem;
      if (j <= i) Sg[i * GR_LDV + j] = fmaf(-Sg[i * GR_LDV + k], Sg[j * GR_LDV + k], Sg[i * GR_LDV + j]);
    }
    __syncthreads();
  }
  if (tid < N) {
    const float* e = eps + ((size_t)sidx * D + d) * N;
    float acc = cst + mn[tid];
    for (int j = 0; j <= tid; ++j) acc = fmaf(Sg[tid * GR_LDV + j], __ldg(e + j), acc);
    out[(size_t)(sidx * N + tid) * ldo + d] = acc;
  }
}

int gp_big_rsample_launch(dvg_gp_s* h, int S, int N, const float* x, int ldx, const float* eps, const uint8_t* mask,
                          float* out, int ldo, cudaStream_t stream) {
  DVG_REQUIRE(N >= 1 && N <= GR_NMAX, "rsample correlates at most %d points per call (got %d)", GR_NMAX, N);
  const size_t smem = sizeof(float) * ((size_t)3 * GB_T * GB_LD + 2 * GB_T * GR_LDV + (size_t)GR_NMAX * GR_LDV + 2 * GR_NMAX + GB_T);
  static bool configured = false;
  if (!configured) {
    DVG_CUDA(cudaFuncSetAttribute(gp_big_rsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  gp_big_rsample_kernel<<<dim3(S, h->dims.num_dims), 256, smem, stream>>>(S, N, h->dims.num_dims, h->mp, x, ldx, eps, mask, h->z,
                                                                          h->linv, h->lqt, h->alpha, h->hyp, out, ldo);
  DVG_LAUNCH_CHECK();
  return DVG_OK;
}

int gp_big_trigger_launch(dvg_gp_s* h, int S, const float* x, int ldx, const int32_t* stat_rows, float* window, int W,
                          int32_t* count, int warmup, float factor, float* value, float* thr, uint8_t* mask,
                          cudaStream_t stream) {
  DVG_REQUIRE(W >= 1 && W <= MAX_WINDOW, "window_len must be in [1,%d]", MAX_WINDOW);
  DVG_REQUIRE(S <= h->var_rows_cap, "n_rollouts=%d exceeds the reserved trigger scratch (%d)", S, h->var_rows_cap);
  DVG_CUDA(cudaMemsetAsync(h->trig_count, 0, sizeof(int), stream));
  // variance of the statistic row of every rollout, written transposed: var_rows[d][s]
  int rc = gp_big_predict_launch(h, S, x, ldx, stat_rows, nullptr, 0, 0, h->var_rows, 1, S, stream);
  if (rc) return rc;
  gp_trigger_finalize_kernel<<<1, 1024, 0, stream>>>(S, h->dims.num_dims, h->var_rows, window, W, count, warmup, factor,
                                                     value, thr, mask, h->trig_list, h->trig_count);
  DVG_LAUNCH_CHECK();
  return DVG_OK;
}

}  // namespace dvg
