// Eval-mode constants of the variational GP for LARGE inducing sets (M > 64, BASELINE configs[4]), once per weight load:
//   K = k(Z,Z) + jitter I        L = chol(K)        Linv = L^-1        beta = Linv (m_q - c)
// (models/gp_models.py:10-24 + gpytorch's WhitenedVariationalStrategy recompute chol(K_ZZ) on every call; here it is
// hoisted.)  For M <= 128 gp_prepare_kernel does this in shared memory; this file is the blocked fp64 version for any M:
// right-looking Cholesky on 64 x 64 blocks -- the diagonal block is factorised AND inverted in shared memory, the
// panel below it is multiplied by that inverse, the trailing matrix is updated by a tiled fp64 GEMM -- followed by the
// block-column recursion  Linv[j+1.., j] = -Linv[j+1.., j+1..] L[j+1.., j] L_jj^-1  (last block column first).  All of it is FP64 FMA work (B200 has the
// full-rate FP64 pipe; tensor cores do not apply), batched over the latent dims that fit the workspace.
#include <vector>

#include "internal.cuh"

namespace dvg {

namespace {

constexpr int FZ_NB = 64;          // block size of the factorisation
constexpr int FZ_KC = 16;          // k chunk of the GEMM tiles

__device__ __forceinline__ double fz_softplus(double x) { return x > 30.0 ? x : log1p(exp(x)); }

// A[d][i][j] = s exp(-0.5 ((z_i - z_j) / ell)^2) + jitter [i == j]; rows / columns >= M: identity (keeps the padded
// matrix positive definite; the padding never mixes with the real block)
__global__ void fz_build_kernel(int M, int Mp, double jitter, const float* __restrict__ inducing,
                                const float* __restrict__ raw_os, const float* __restrict__ raw_ls, int d0, double* A) {
  const int d = blockIdx.z;
  const int i = blockIdx.y * 16 + threadIdx.y, j = blockIdx.x * 16 + threadIdx.x;
  const int gd = d0 + d;
  const double ell = fz_softplus((double)raw_ls[gd]), s = fz_softplus((double)raw_os[gd]);
  double v = i == j ? 1.0 : 0.0;
  if (i < M && j < M) {
    const double t = ((double)inducing[(size_t)gd * M + i] - (double)inducing[(size_t)gd * M + j]) / ell;
    v = s * exp(-0.5 * t * t) + (i == j ? jitter : 0.0);
  }
  A[((size_t)d * Mp + i) * Mp + j] = v;
}

// Diagonal block kb: in-place Cholesky in shared memory, L_kk back to A (upper part zeroed), L_kk^-1 to dinv[d][kb].
// One CTA of 256 threads per latent dim.  `bad` is set when a pivot is not positive.
__global__ void __launch_bounds__(256) fz_diag_kernel(int Mp, int kb, int nb, double* A, double* dinv, int* bad) {
  extern __shared__ double fz_smem[];                      // 2 x 64 x 65 doubles (65 KB: dynamic)
  double (*s_a)[FZ_NB + 1] = reinterpret_cast<double (*)[FZ_NB + 1]>(fz_smem);
  double (*s_x)[FZ_NB + 1] = reinterpret_cast<double (*)[FZ_NB + 1]>(fz_smem + FZ_NB * (FZ_NB + 1));
  const int d = blockIdx.x, tid = threadIdx.x;
  double* blk = A + ((size_t)d * Mp + (size_t)kb * FZ_NB) * Mp + (size_t)kb * FZ_NB;
  for (int e = tid; e < FZ_NB * FZ_NB; e += 256) s_a[e / FZ_NB][e % FZ_NB] = blk[(size_t)(e / FZ_NB) * Mp + (e % FZ_NB)];
  __syncthreads();
  for (int j = 0; j < FZ_NB; ++j) {
    if (tid == 0) {
      const double p = s_a[j][j];
      if (!(p > 0.0)) *bad = 1;
      s_a[j][j] = sqrt(p > 0.0 ? p : 1.0);
    }
    __syncthreads();
    if (tid > j && tid < FZ_NB) s_a[tid][j] /= s_a[j][j];
    __syncthreads();
    // trailing update of the lower triangle: (r, c) with j < c <= r < NB
    const int n = FZ_NB - 1 - j;                 // rows / columns left
    for (int e = tid; e < n * n; e += 256) {
      const int r = j + 1 + e / n, c = j + 1 + e % n;
      if (c <= r) s_a[r][c] -= s_a[r][j] * s_a[c][j];
    }
    __syncthreads();
  }
  // X = L^-1 (lower): thread c solves column c by forward substitution
  if (tid < FZ_NB) {
    const int c = tid;
    for (int r = 0; r < c; ++r) s_x[r][c] = 0.0;
    s_x[c][c] = 1.0 / s_a[c][c];
    for (int r = c + 1; r < FZ_NB; ++r) {
      double acc = 0.0;
      for (int k = c; k < r; ++k) acc = fma(s_a[r][k], s_x[k][c], acc);
      s_x[r][c] = -acc / s_a[r][r];
    }
  }
  __syncthreads();
  double* xo = dinv + ((size_t)d * nb + kb) * FZ_NB * FZ_NB;
  for (int e = tid; e < FZ_NB * FZ_NB; e += 256) {
    const int r = e / FZ_NB, c = e % FZ_NB;
    blk[(size_t)r * Mp + c] = c <= r ? s_a[r][c] : 0.0;
    xo[e] = s_x[r][c];
  }
}

// C[m x n] = alpha * A[m x k] * op(B) + beta * C, fp64, row-major with leading dimensions; op(B) = B^T with B [n x k]
// (BT) or B [k x n].  n a multiple of 64, k a multiple of 16, m any multiple of 64 (rows past m are masked).  Grid
// (n / 64, ceil(m / 128), batch), 256 threads, each thread 8 rows x 4 (strided) columns of a 128 x 64 tile of C: 12 shared-memory
// loads per 32 FMAs (the 64 x 64 / 4 x 4 version was shared-memory bound at a third of the FP64 pipe).
//   lower_only  C square: tiles entirely above the diagonal are skipped (trailing update of the Cholesky)
//   tri_b       B [k x n] lower triangular in 64-blocks: the k blocks above output column block x are zero, skipped
//   tri_a       A [m x k] lower triangular in 64-blocks: k stops at the last block row of the tile
// A CTA reads everything it needs of A and B before it writes C, so C may alias A's own tile rows (the in-place panel
// solve).
constexpr int FZ_BM = 128;
template <bool BT>
__global__ void __launch_bounds__(256) fz_gemm_kernel(int m, int k, double alpha, const double* A, int lda, long long sa,
                                                      const double* B, int ldb, long long sb, double beta, double* C,
                                                      int ldc, long long sc, int lower_only, int tri_b, int tri_a) {
  if (lower_only && (int)blockIdx.x * FZ_NB > (int)blockIdx.y * FZ_BM + FZ_BM - 1) return;
  __shared__ double s_A[FZ_KC][FZ_BM + 2];      // [kk][row]
  __shared__ double s_B[FZ_KC][FZ_NB + 2];      // [kk][col]
  const int tid = threadIdx.x;
  const int tr = tid / 16, tc = tid % 16;       // 16 x 16 threads, 8 x 4 outputs each
  const int row0 = (int)blockIdx.y * FZ_BM;
  const double* Ab = A + (size_t)blockIdx.z * sa + (size_t)row0 * lda;
  const double* Bb = B + (size_t)blockIdx.z * sb;
  double* Cb = C + (size_t)blockIdx.z * sc + (size_t)row0 * ldc + (size_t)blockIdx.x * FZ_NB;
  double acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  const int k_begin = tri_b ? (int)blockIdx.x * FZ_NB : 0;
  int k_end = k;
  if (tri_a && row0 + FZ_BM < k_end) k_end = row0 + FZ_BM;
  for (int k0 = k_begin; k0 < k_end; k0 += FZ_KC) {
    // A tile: 128 rows x 16 k (2048 doubles, 8 per thread)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = tid + q * 256;
      const int r = e / FZ_KC, kk = e % FZ_KC;
      s_A[kk][r] = row0 + r < m ? Ab[(size_t)r * lda + k0 + kk] : 0.0;
    }
    if (BT) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {             // B is [n x k]: rows = output columns
        const int e = tid + q * 256;
        const int c = e / FZ_KC, kk = e % FZ_KC;
        s_B[kk][c] = Bb[(size_t)(blockIdx.x * FZ_NB + c) * ldb + k0 + kk];
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {             // B is [k x n]
        const int e = tid + q * 256;
        const int kk = e / FZ_NB, c = e % FZ_NB;
        s_B[kk][c] = Bb[(size_t)(k0 + kk) * ldb + blockIdx.x * FZ_NB + c];
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < FZ_KC; ++kk) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = s_A[kk][tr * 8 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = s_B[kk][tc + 16 * j];      // columns tc, tc + 16, ...: conflict-free, coalesced stores
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (row0 + tr * 8 + i >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double* cp = Cb + (size_t)(tr * 8 + i) * ldc + tc + 16 * j;
      *cp = beta == 0.0 ? alpha * acc[i][j] : alpha * acc[i][j] + beta * *cp;
    }
  }
}

// Linv64[d][i-block][i-block] = dinv[d][i]
__global__ void fz_place_diag_kernel(int Mp, int nb, int i, const double* dinv, double* X) {
  const int d = blockIdx.x;
  const double* src = dinv + ((size_t)d * nb + i) * FZ_NB * FZ_NB;
  double* dst = X + ((size_t)d * Mp + (size_t)i * FZ_NB) * Mp + (size_t)i * FZ_NB;
  for (int e = threadIdx.x; e < FZ_NB * FZ_NB; e += blockDim.x) dst[(size_t)(e / FZ_NB) * Mp + e % FZ_NB] = src[e];
}

// fp32 outputs: linv [D][M][M] (lower, zeros above) and beta = Linv (m_q - c), one CTA per (row, dim)
__global__ void __launch_bounds__(128) fz_finish_kernel(int M, int Mp, int d0, const double* X, const float* __restrict__ var_mean,
                                                        const float* __restrict__ mean_const, float* linv, float* beta) {
  __shared__ double s_red[128];
  const int d = blockIdx.y, r = blockIdx.x, gd = d0 + d;
  const double* row = X + ((size_t)d * Mp + r) * Mp;
  float* out = linv + ((size_t)gd * M + r) * M;
  const double c = (double)mean_const[gd];
  double acc = 0.0;
  for (int j = threadIdx.x; j < M; j += 128) {
    const double v = j <= r ? row[j] : 0.0;
    out[j] = (float)v;
    acc = fma(v, (double)var_mean[(size_t)gd * M + j] - c, acc);
  }
  s_red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 64; s > 0; s >>= 1) {
    if (threadIdx.x < s) s_red[threadIdx.x] += s_red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) beta[(size_t)gd * M + r] = (float)s_red[0];
}

}  // namespace

// Workspace of gp_factorize for `batch` latent dims per pass: an error flag, then per dim two Mp x Mp fp64 matrices
// (K / L in place, Linv), the inverted diagonal blocks and one 64-row scratch panel.
static size_t fz_dim_bytes(int M) {
  const size_t Mp = align_up((size_t)M, FZ_NB), nb = Mp / FZ_NB;
  return sizeof(double) * (2 * Mp * Mp + nb * FZ_NB * FZ_NB + FZ_NB * Mp);
}
size_t gp_factorize_workspace(int M, int batch) { return 256 + (size_t)(batch < 1 ? 1 : batch) * fz_dim_bytes(M); }

// Host driver.  The caller owns the workspace (device memory, any size >= gp_factorize_workspace(M, 1): as many dims per
// pass as fit).  Synchronises the stream at the end to read the "not positive definite" flag back: once per weight load.
int gp_factorize(int D, int M, double jitter, const float* inducing, const float* var_mean, const float* mean_const,
                 const float* raw_os, const float* raw_ls, float* linv, float* beta, void* workspace,
                 size_t workspace_bytes, cudaStream_t stream) {
  const int Mp = (int)align_up((size_t)M, FZ_NB), nb = Mp / FZ_NB;
  const size_t mat = (size_t)Mp * Mp;
  DVG_REQUIRE(workspace != nullptr && workspace_bytes >= gp_factorize_workspace(M, 1) &&
                  (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
              "factorisation workspace too small or misaligned (%zu bytes, need >= %zu, 256-byte aligned)", workspace_bytes,
              gp_factorize_workspace(M, 1));
  size_t b = (workspace_bytes - 256) / fz_dim_bytes(M);
  if (b > (size_t)D) b = (size_t)D;
  int* bad = reinterpret_cast<int*>(workspace);
  double* A = reinterpret_cast<double*>(reinterpret_cast<uint8_t*>(workspace) + 256);
  double* X = A + b * mat;
  double* dinv = X + b * mat;
  double* T = dinv + b * nb * FZ_NB * FZ_NB;
  cudaError_t e = cudaMemsetAsync(bad, 0, sizeof(int), stream);
  if (e != cudaSuccess) {
    set_error("GP factorisation: %s", cudaGetErrorString(e));
    return DVG_ERR_CUDA;
  }
  constexpr size_t diag_smem = sizeof(double) * 2 * FZ_NB * (FZ_NB + 1);
  static bool configured = false;
  if (!configured) {
    e = cudaFuncSetAttribute(fz_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)diag_smem);
    if (e != cudaSuccess) {
      set_error("GP factorisation: %s", cudaGetErrorString(e));
      return DVG_ERR_CUDA;
    }
    configured = true;
  }
  int rc = DVG_OK;
  auto fail = [&](cudaError_t err) {
    if (err != cudaSuccess && rc == DVG_OK) {
      set_error("GP factorisation: %s", cudaGetErrorString(err));
      rc = DVG_ERR_CUDA;
    }
    return err != cudaSuccess;
  };
  for (int d0 = 0; d0 < D && rc == DVG_OK; d0 += (int)b) {
    const int nbat = D - d0 < (int)b ? D - d0 : (int)b;
    fz_build_kernel<<<dim3(Mp / 16, Mp / 16, nbat), dim3(16, 16), 0, stream>>>(M, Mp, jitter, inducing, raw_os, raw_ls, d0, A);
    if (fail(cudaGetLastError())) break;
    // ---- blocked Cholesky, in place on A (lower part) ----
    for (int kb = 0; kb < nb; ++kb) {
      fz_diag_kernel<<<nbat, 256, diag_smem, stream>>>(Mp, kb, nb, A, dinv, bad);
      const int rem = nb - kb - 1;
      if (rem == 0) break;
      double* panel = A + (size_t)(kb + 1) * FZ_NB * Mp + (size_t)kb * FZ_NB;       // rows below the diagonal block
      // panel <- panel * (L_kk^-1)^T
      const int mrows = rem * FZ_NB, mt = (mrows + FZ_BM - 1) / FZ_BM;
      fz_gemm_kernel<true><<<dim3(1, mt, nbat), 256, 0, stream>>>(mrows, FZ_NB, 1.0, panel, Mp, (long long)mat,
                                                                dinv + (size_t)kb * FZ_NB * FZ_NB, FZ_NB,
                                                                (long long)nb * FZ_NB * FZ_NB, 0.0, panel, Mp, (long long)mat, 0, 0, 0);
      // trailing <- trailing - panel * panel^T   (tiles entirely above the diagonal skipped)
      double* trail = A + (size_t)(kb + 1) * FZ_NB * Mp + (size_t)(kb + 1) * FZ_NB;
      fz_gemm_kernel<true><<<dim3(rem, mt, nbat), 256, 0, stream>>>(mrows, FZ_NB, -1.0, panel, Mp, (long long)mat, panel, Mp,
                                                                  (long long)mat, 1.0, trail, Mp, (long long)mat, 1, 0, 0);
    }
    if (fail(cudaGetLastError())) break;
    // ---- Linv = L^-1 by block columns, last to first:  X[j+1.., j] = -X[j+1.., j+1..] L[j+1.., j] L_jj^-1 ----
    if (fail(cudaMemsetAsync(X, 0, sizeof(double) * nbat * mat, stream))) break;
    for (int j = nb - 1; j >= 0; --j) {
      fz_place_diag_kernel<<<nbat, 256, 0, stream>>>(Mp, nb, j, dinv, X);
      const int rem = nb - 1 - j;
      if (rem == 0) continue;
      const int mrows = rem * FZ_NB, mt = (mrows + FZ_BM - 1) / FZ_BM;
      const double* xt = X + (size_t)(j + 1) * FZ_NB * Mp + (size_t)(j + 1) * FZ_NB;     // trailing inverse (lower triangular)
      const double* lp = A + (size_t)(j + 1) * FZ_NB * Mp + (size_t)j * FZ_NB;           // panel of L below block j
      // T[m x 64] = X_trail * L_panel          (A operand triangular: k stops at the tile's last block row)
      fz_gemm_kernel<false><<<dim3(1, mt, nbat), 256, 0, stream>>>(mrows, mrows, 1.0, xt, Mp, (long long)mat, lp, Mp,
                                                                 (long long)mat, 0.0, T, FZ_NB, (long long)FZ_NB * Mp, 0, 0, 1);
      // X[j+1.., j] = -T * L_jj^-1
      fz_gemm_kernel<false><<<dim3(1, mt, nbat), 256, 0, stream>>>(mrows, FZ_NB, -1.0, T, FZ_NB, (long long)FZ_NB * Mp,
                                                                 dinv + (size_t)j * FZ_NB * FZ_NB, FZ_NB,
                                                                 (long long)nb * FZ_NB * FZ_NB, 0.0,
                                                                 X + (size_t)(j + 1) * FZ_NB * Mp + (size_t)j * FZ_NB, Mp,
                                                                 (long long)mat, 0, 0, 0);
    }
    if (fail(cudaGetLastError())) break;
    fz_finish_kernel<<<dim3(M, nbat), 128, 0, stream>>>(M, Mp, d0, X, var_mean, mean_const, linv, beta);
    if (fail(cudaGetLastError())) break;
  }
  int hbad = 0;
  if (rc == DVG_OK) {
    if (!fail(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, stream))) fail(cudaStreamSynchronize(stream));
  } else {
    cudaStreamSynchronize(stream);
  }
  if (rc == DVG_OK && hbad) {
    set_error("K_ZZ + jitter I is not positive definite in fp64 (non-positive Cholesky pivot)");
    return DVG_ERR_ARG;
  }
  return rc;
}

}  // namespace dvg
