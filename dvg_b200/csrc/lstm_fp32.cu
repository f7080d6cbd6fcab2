// DVG_FP32 variant of the LSTM hot path: CUDA-core FFMA GEMMs with exact fp32 products and the LSTM
// pointwise math fused into the epilogue.  One generic 64x64x16 smem-tiled kernel, four epilogues:
//   EPI_BIAS  embed              (models/lstm.py:66)
//   EPI_LSTM  LSTMCell           (models/lstm.py:69)   gates -> c', h'
//   EPI_TANH  output Linear+Tanh (models/lstm.py:72)
//   EPI_GAUSS mu/logvar heads + reparameterize (models/lstm.py:161-164,172-174)
// This variant is the numerics anchor for the tensor-core variants (lstm_tc.cu) and the path for
// sizes they do not cover (hidden_size % 64 != 0).
#include "internal.cuh"

namespace dvg {

enum { EPI_BIAS = 0, EPI_LSTM = 1, EPI_TANH = 2, EPI_GAUSS = 3 };

struct FfmaArgs {
  int rows;
  const float* a0; int lda0; int k0;
  const float* a1; int lda1; int k1;
  const float* wt; int ldw;
  const float* bias;
  int n;  // valid output columns (packed order)
  float* out; int ldo;
  const float* c_in; const float* h_in; float* h_out; float* c_out; int H;
  const uint8_t* hold; int rows_per_flag;
  const float* eps; float* z; float* mu; float* logvar; int Z;
};

constexpr int BM = 64, BN = 64, BK = 16;

template <int EPI>
__global__ void __launch_bounds__(256) ffma_gemm_kernel(FfmaArgs p) {
  __shared__ float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ar = tid >> 2, akq = (tid & 3) * 4;  // A loader: row ar, k akq..akq+3
  const int bk = tid >> 4, bnq = (tid & 15) * 4;  // B loader: k bk, cols bnq..bnq+3

  for (int src = 0; src < 2; ++src) {
    const float* a = src == 0 ? p.a0 : p.a1;
    const int lda = src == 0 ? p.lda0 : p.lda1;
    const int K = src == 0 ? p.k0 : p.k1;
    const int koff = src == 0 ? 0 : p.k0;
    if (a == nullptr || K == 0) continue;
    for (int kb = 0; kb < K; kb += BK) {
      {
        const int r = row0 + ar;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k = kb + akq + i;
          As[akq + i][ar] = (r < p.rows && k < K) ? __ldg(a + (size_t)r * lda + k) : 0.f;
        }
        const int k = kb + bk;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K) w = __ldg(reinterpret_cast<const float4*>(p.wt + (size_t)(koff + k) * p.ldw + n0 + bnq));
        *reinterpret_cast<float4*>(&Bs[bk][bnq]) = w;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float av[4], bv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = As[k][ty * 4 + i];
        const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  const int nc = n0 + tx * 4;
  float b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = (nc + j < p.n) ? __ldg(p.bias + nc + j) : 0.f;

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r >= p.rows) continue;
    if (EPI == EPI_BIAS || EPI == EPI_TANH) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (nc + j < p.n) {
          float v = acc[i][j] + b[j];
          if (EPI == EPI_TANH) v = tanh_f(v);
          p.out[(size_t)r * p.ldo + nc + j] = v;
        }
      }
    } else if (EPI == EPI_LSTM) {
      const int u = nc >> 2;
      if (u < p.H) {
        const size_t idx = (size_t)r * p.H + u;
        const float c_prev = p.c_in[idx];
        const bool held = p.hold != nullptr && p.hold[r / p.rows_per_flag] != 0;
        if (held) {
          p.c_out[idx] = c_prev;
          p.h_out[idx] = p.h_in[idx];
        } else {
          const float gi = sigmoid_f(acc[i][0] + b[0]);
          const float gf = sigmoid_f(acc[i][1] + b[1]);
          const float gg = tanh_f(acc[i][2] + b[2]);
          const float go = sigmoid_f(acc[i][3] + b[3]);
          const float c_new = fmaf(gf, c_prev, gi * gg);
          p.c_out[idx] = c_new;
          p.h_out[idx] = go * tanh_f(c_new);
        }
      }
    } else {  // EPI_GAUSS: columns (2z, 2z+1) = (mu_z, logvar_z)
#pragma unroll
      for (int j = 0; j < 4; j += 2) {
        const int zi = (nc + j) >> 1;
        if (zi < p.Z) {
          const float m = acc[i][j] + b[j];
          const float lv = acc[i][j + 1] + b[j + 1];
          const size_t idx = (size_t)r * p.Z + zi;
          p.mu[idx] = m;
          p.logvar[idx] = lv;
          p.z[idx] = fmaf(p.eps[idx], expf(0.5f * lv), m);   // models/lstm.py:162-164
        }
      }
    }
  }
}

// dst[k][col(n)] = src[row(n)][k]; see modes in lstm_fp32_pack.
__global__ void pack_wt_kernel(float* dst, int ldw, const float* src, int K, int n_cols, int mode, int H) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (n >= n_cols || k >= K) return;
  int row, col;
  if (mode == 0) { row = n; col = n; }                          // identity
  else if (mode == 1) { row = (n & 3) * H + (n >> 2); col = n; }  // gate-interleaved: col = unit*4 + gate
  else if (mode == 2) { row = n; col = 2 * n; }                 // mu -> even columns
  else { row = n; col = 2 * n + 1; }                            // logvar -> odd columns
  dst[(size_t)k * ldw + col] = src[(size_t)row * K + k];
}
__global__ void pack_bias_kernel(float* dst, const float* b0, const float* b1, int n_cols, int mode, int H) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_cols) return;
  int row, col;
  if (mode == 0) { row = n; col = n; }
  else if (mode == 1) { row = (n & 3) * H + (n >> 2); col = n; }
  else if (mode == 2) { row = n; col = 2 * n; }
  else { row = n; col = 2 * n + 1; }
  dst[col] = b0[row] + (b1 ? b1[row] : 0.f);
}

static int pack_wt(float* dst, int ldw, const float* src, int K, int n_cols, int mode, int H, cudaStream_t s) {
  dim3 grid(ceil_div(n_cols, 128), K);
  pack_wt_kernel<<<grid, 128, 0, s>>>(dst, ldw, src, K, n_cols, mode, H);
  DVG_LAUNCH_CHECK();
  return DVG_OK;
}
static int pack_bias(float* dst, const float* b0, const float* b1, int n_cols, int mode, int H, cudaStream_t s) {
  pack_bias_kernel<<<ceil_div(n_cols, 128), 128, 0, s>>>(dst, b0, b1, n_cols, mode, H);
  DVG_LAUNCH_CHECK();
  return DVG_OK;
}

int lstm_fp32_pack(dvg_lstm_s* h, const float* embed_w, const float* embed_b, const float* const* w_ih,
                   const float* const* w_hh, const float* const* b_ih, const float* const* b_hh,
                   const float* head0_w, const float* head0_b, const float* head1_w, const float* head1_b,
                   cudaStream_t stream) {
  const int G = h->dims.input_size, H = h->dims.hidden_size, L = h->dims.n_layers;
  const bool gauss = h->dims.kind == DVG_GAUSSIAN_LSTM;
  h->hp = (int)align_up(H, 64);
  h->n_head = gauss ? 2 * h->dims.output_size : h->dims.output_size;
  h->np_head = (int)align_up(h->n_head, 64);
  const int n4 = 4 * H;  // H % 16 == 0 is required so 4H % 64 == 0
  if (!h->f_embed_wt) {
    DVG_CUDA(cudaMalloc(&h->f_embed_wt, sizeof(float) * G * h->hp));
    DVG_CUDA(cudaMalloc(&h->f_embed_b, sizeof(float) * h->hp));
    for (int l = 0; l < L; ++l) {
      DVG_CUDA(cudaMalloc(&h->f_layer_wt[l], sizeof(float) * 2 * H * n4));
      DVG_CUDA(cudaMalloc(&h->f_layer_b[l], sizeof(float) * n4));
    }
    DVG_CUDA(cudaMalloc(&h->f_head_wt, sizeof(float) * H * h->np_head));
    DVG_CUDA(cudaMalloc(&h->f_head_b, sizeof(float) * h->np_head));
  }
  DVG_CUDA(cudaMemsetAsync(h->f_embed_wt, 0, sizeof(float) * G * h->hp, stream));
  DVG_CUDA(cudaMemsetAsync(h->f_embed_b, 0, sizeof(float) * h->hp, stream));
  DVG_CUDA(cudaMemsetAsync(h->f_head_wt, 0, sizeof(float) * H * h->np_head, stream));
  DVG_CUDA(cudaMemsetAsync(h->f_head_b, 0, sizeof(float) * h->np_head, stream));
  int rc;
  if ((rc = pack_wt(h->f_embed_wt, h->hp, embed_w, G, H, 0, H, stream))) return rc;
  if ((rc = pack_bias(h->f_embed_b, embed_b, nullptr, H, 0, H, stream))) return rc;
  for (int l = 0; l < L; ++l) {
    if ((rc = pack_wt(h->f_layer_wt[l], n4, w_ih[l], H, n4, 1, H, stream))) return rc;
    if ((rc = pack_wt(h->f_layer_wt[l] + (size_t)H * n4, n4, w_hh[l], H, n4, 1, H, stream))) return rc;
    if ((rc = pack_bias(h->f_layer_b[l], b_ih[l], b_hh[l], n4, 1, H, stream))) return rc;
  }
  if (gauss) {
    const int Z = h->dims.output_size;
    if ((rc = pack_wt(h->f_head_wt, h->np_head, head0_w, H, Z, 2, H, stream))) return rc;
    if ((rc = pack_wt(h->f_head_wt, h->np_head, head1_w, H, Z, 3, H, stream))) return rc;
    if ((rc = pack_bias(h->f_head_b, head0_b, nullptr, Z, 2, H, stream))) return rc;
    if ((rc = pack_bias(h->f_head_b, head1_b, nullptr, Z, 3, H, stream))) return rc;
  } else {
    if ((rc = pack_wt(h->f_head_wt, h->np_head, head0_w, H, h->n_head, 0, H, stream))) return rc;
    if ((rc = pack_bias(h->f_head_b, head0_b, nullptr, h->n_head, 0, H, stream))) return rc;
  }
  return DVG_OK;
}

int lstm_fp32_step(dvg_lstm_s* h, int rows, const float* x, int ldx, const float* h_in, const float* c_in,
                   float* h_out, float* c_out, float* y, int ldy, const float* eps, float* z, float* mu,
                   float* logvar, const uint8_t* hold, int rows_per_flag, cudaStream_t stream) {
  const int G = h->dims.input_size, H = h->dims.hidden_size, L = h->dims.n_layers;
  const size_t lsz = (size_t)rows * H;
  const int rt = ceil_div(rows, BM);
  FfmaArgs a{};
  a.rows = rows;
  // embed
  a.a0 = x; a.lda0 = ldx; a.k0 = G; a.a1 = nullptr; a.k1 = 0;
  a.wt = h->f_embed_wt; a.ldw = h->hp; a.bias = h->f_embed_b; a.n = H;
  a.out = h->scratch_e; a.ldo = H;
  h->prof_mark(stream);
  h->prof_mark(stream);
  ffma_gemm_kernel<EPI_BIAS><<<dim3(rt, h->hp / BN), 256, 0, stream>>>(a);
  DVG_LAUNCH_CHECK();
  h->prof_mark(stream);
  const float* layer_in = h->scratch_e;
  for (int l = 0; l < L; ++l) {
    FfmaArgs b{};
    b.rows = rows;
    b.a0 = layer_in; b.lda0 = H; b.k0 = H;
    b.a1 = h_in + l * lsz; b.lda1 = H; b.k1 = H;
    b.wt = h->f_layer_wt[l]; b.ldw = 4 * H; b.bias = h->f_layer_b[l]; b.n = 4 * H;
    b.c_in = c_in + l * lsz; b.h_in = h_in + l * lsz; b.h_out = h_out + l * lsz; b.c_out = c_out + l * lsz;
    b.H = H; b.hold = hold; b.rows_per_flag = rows_per_flag > 0 ? rows_per_flag : 1;
    ffma_gemm_kernel<EPI_LSTM><<<dim3(rt, 4 * H / BN), 256, 0, stream>>>(b);
    DVG_LAUNCH_CHECK();
    h->prof_mark(stream);
    layer_in = h_out + l * lsz;
  }
  FfmaArgs c{};
  c.rows = rows;
  c.a0 = layer_in; c.lda0 = H; c.k0 = H;
  c.wt = h->f_head_wt; c.ldw = h->np_head; c.bias = h->f_head_b; c.n = h->n_head;
  if (h->dims.kind == DVG_GAUSSIAN_LSTM) {
    c.eps = eps; c.z = z; c.mu = mu; c.logvar = logvar; c.Z = h->dims.output_size;
    ffma_gemm_kernel<EPI_GAUSS><<<dim3(rt, h->np_head / BN), 256, 0, stream>>>(c);
  } else {
    c.out = y; c.ldo = ldy;
    ffma_gemm_kernel<EPI_TANH><<<dim3(rt, h->np_head / BN), 256, 0, stream>>>(c);
  }
  DVG_LAUNCH_CHECK();
  h->prof_mark(stream);
  return DVG_OK;
}

}  // namespace dvg
