// GP predictive mean / variance for LARGE inducing sets on the tensor cores (BASELINE configs[4]: M = 128 .. 4096).
// Same math as gp_big.cu (SURVEY 8c eqs. 1-6 with the factors hoisted):
//   V = Linv K_zx,  W = L_q^T K_zx,   var = s - colsum(V^2) + colsum(W^2) + noise,   mean = c + V^T beta
// organised so that the tcgen05 tile matches the reduction the epilogue needs:
//
//     D^T [128 points x 256 factor rows j]  =  K_xz tile [128 points x 64 m]  .  F tile [256 j x 64 m]^T        (F = Linv or L_q^T)
//
//   * A operand = the KERNEL tile k(x_n, z_m), built on the fly by eight SIMT warps (one MUFU ex2 per element, bf16 hi/lo
//     split, written straight into the 128-byte-swizzled K-major shared-memory image): never stored in HBM;
//   * B operand = a 256-row tile of the triangular factor, pre-packed once per weight load as bf16 hi/lo k-block images
//     (only tiles inside the triangle are ever read), streamed L2 -> smem by 1-D TMA bulk copies;
//   * bf16x3 products (hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM): variance within ~3e-6 of fp64 on the test sets;
//   * accumulators: V in TMEM columns 0..255, W in 256..511, lane = point -- so the per-point sums over the factor rows j
//     (|v|^2, |w|^2) are plain per-thread loops over TMEM columns, no cross-lane reduction;
//   * the MEAN does not go through the tensor cores: v . beta with bf16x3 products was 1.7e-4 from fp64 on the test sets
//     (beta weights the cancellation-heavy rows of V); instead the builder warps, which hold every k(x_n, z_m) in fp32
//     anyway, accumulate  k . alpha2  with alpha2 = Linv^T beta pre-computed in fp64 -- the reference's own arithmetic
//     (K_xz K_zz^-1 (m - c)), 4e-7 from fp64 on the same sets -- in the CTAs of the last row tile, whose V passes visit
//     every m block exactly once;
//   * a CTA owns (128 points, 256 factor rows, one latent dim) and walks the 64-wide m blocks: V needs m <= j, W needs
//     m >= j, the four blocks on the diagonal band are passed twice.  Row-block partial sums go to the same scratch /
//     finalize kernel as the FP32 path (fixed summation order: deterministic).
// Bound: tensor pipe / shared-memory bandwidth (1-CTA 128 x 256 SS MMAs); the FP32 FFMA version reaches 24 TFLOP/s.
#include <stdlib.h>

#include "internal.cuh"
#include "ptx.cuh"

namespace dvg {

constexpr int GT_THREADS = 64 + 8 * 32;        // warp 0 TMA producer, warp 1 TMEM + MMA issuer, warps 2-9 builders / epilogue
constexpr int GT_A_IMG = 128 * 128;             // A image part: 128 points x 64 bf16
constexpr int GT_B_IMG = 256 * 128;             // B image part: 256 factor rows x 64 bf16
constexpr int GT_STAGE = 2 * GT_A_IMG + 2 * GT_B_IMG;     // 96 KB
constexpr int GT_STAGES = 2;

struct GpTcArgs {
  int n_rows, n_pad, MT, JT, ldx, want_mean;
  const float* x; const int32_t* row_index;
  const float* z;               // [D][Mz] inducing points (zero padded)
  int Mz;                       // row stride of z / beta (= mp of the handle)
  const float* beta; const float* hyp;
  const float* alpha2;          // [D][Mz]  Linv^T beta = K_ZZ^-1 (m_q - c): the mean is k . alpha2 in fp32 (see the builders)
  const uint8_t* img_v;         // [D][JT][MT][hi | lo][256 x 128 B]  Linv
  const uint8_t* img_w;         // [D][JT][MT][hi | lo][256 x 128 B]  L_q^T
  float* partial;               // [D][JT][n_pad][3]
};

__global__ void __launch_bounds__(GT_THREADS, 1) gp_tc_partial_kernel(const __grid_constant__ GpTcArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = ptx::smem_u32(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * 128, jt = blockIdx.y, d = blockIdx.z;
  uint8_t* tail = smem_raw + (size_t)GT_STAGES * GT_STAGE;
  float* s_x = reinterpret_cast<float*>(tail);                 // [128]
  float* s_dot = s_x + 128;                                    // [2][128] k . alpha2 of the two m halves
  float* s_red = s_dot + 256;                                  // [128][3] second column half's partial sums
  const uint32_t bar0 = ptx::smem_u32(s_red + 384);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto b_full = [&](int s) { return bar0 + 8u * (GT_STAGES + s); };
  auto empty = [&](int s) { return bar0 + 8u * (2 * GT_STAGES + s); };
  const uint32_t acc_full = bar0 + 8u * (3 * GT_STAGES);
  const uint32_t tmem_slot = acc_full + 8;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GT_STAGES; ++s) {
      ptx::mbar_init(a_full(s), 8);
      ptx::mbar_init(b_full(s), 1);
      ptx::mbar_init(empty(s), 1);
    }
    ptx::mbar_init(acc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  if (threadIdx.x < 128) {
    const int n = n0 + threadIdx.x;
    float v = 0.f;
    if (n < p.n_rows) v = __ldg(p.x + (size_t)(p.row_index ? p.row_index[n] : n) * p.ldx + d);
    s_x[threadIdx.x] = v;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // pass list: m blocks 0 .. 4 jt + 3 for V (Linv is lower triangular: m <= j), then 4 jt .. MT - 1 for W (L_q^T upper: m >= j)
  const int kv_end = min(4 * jt + 4, p.MT);
  const int n_pass = kv_end + (p.MT - 4 * jt);
  auto pass_kb = [&](int q) { return q < kv_end ? q : 4 * jt + (q - kv_end); };

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol = ptx::l2_policy_evict_last();
      for (int q = 0; q < n_pass; ++q) {
        const int s = q % GT_STAGES;
        if (q >= GT_STAGES) ptx::mbar_wait(empty(s), ((q / GT_STAGES) - 1) & 1);
        const bool isv = q < kv_end;
        const uint8_t* src = (isv ? p.img_v : p.img_w) + ((size_t)((size_t)d * p.JT + jt) * p.MT + pass_kb(q)) * (2u * GT_B_IMG);
        ptx::mbar_expect_tx(b_full(s), 2u * GT_B_IMG);
        ptx::bulk_g2s_hint(base + (uint32_t)s * GT_STAGE + 2 * GT_A_IMG, src, 2u * GT_B_IMG, b_full(s), pol);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, 256);
      uint32_t accum_v = 0, accum_w = 0;
      for (int q = 0; q < n_pass; ++q) {
        const int s = q % GT_STAGES;
        const uint32_t ph = (uint32_t)(q / GT_STAGES) & 1u;
        ptx::mbar_wait(a_full(s), ph);
        ptx::mbar_wait(b_full(s), ph);
        ptx::tc_fence_after();
        const bool isv = q < kv_end;
        const uint32_t d_tmem = tmem_base + (isv ? 0u : 256u);
        uint32_t& accum = isv ? accum_v : accum_w;
        const uint32_t sa = base + (uint32_t)s * GT_STAGE;
        const uint64_t a_hi = ptx::make_sw128_desc(sa), a_lo = ptx::make_sw128_desc(sa + GT_A_IMG);
        const uint64_t b_hi = ptx::make_sw128_desc(sa + 2 * GT_A_IMG), b_lo = ptx::make_sw128_desc(sa + 2 * GT_A_IMG + GT_B_IMG);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t adv = (uint64_t)(kk * 2);
          ptx::umma_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc, accum);
          accum = 1u;
          ptx::umma_bf16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
          ptx::umma_bf16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
        }
        ptx::umma_commit(empty(s));
      }
      ptx::umma_commit(acc_full);
    }
  } else {
    // ===================== kernel-tile builders (8 warps), then the epilogue =====================
    const int bt = threadIdx.x - 64;                 // 0..255
    const int n = bt & 127, mh = bt >> 7;            // point, half of the 64 m columns (32 each)
    const float xv = s_x[n];
    const float inv_ell = 1.0f / __ldg(p.hyp + d * 4 + 0), sc = __ldg(p.hyp + d * 4 + 1);
    const bool mean_cta = p.want_mean && jt == p.JT - 1;      // its V passes cover m blocks 0 .. MT-1 exactly once
    float dot0 = 0.f, dot1 = 0.f;
    for (int q = 0; q < n_pass; ++q) {
      const int s = q % GT_STAGES;
      if (q >= GT_STAGES) ptx::mbar_wait(empty(s), ((q / GT_STAGES) - 1) & 1);
      const int m0 = pass_kb(q) * 64 + mh * 32;
      uint8_t* img_hi = smem_raw + (size_t)s * GT_STAGE + (size_t)n * 128;
      uint8_t* img_lo = img_hi + GT_A_IMG;
#pragma unroll
      for (int c = 0; c < 4; ++c) {                  // four 16-byte chunks of 8 columns each
        float k[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int m = m0 + c * 8 + e;
          const float zz = m < p.Mz ? __ldg(p.z + (size_t)d * p.Mz + m) : 1e30f;     // padded columns: k = 0
          const float t = (xv - zz) * inv_ell;
          k[e] = sc * ex2_ftz(fmaxf(t * t * (-0.5f * kLog2e), -126.f));
        }
        if (mean_cta && q < kv_end) {
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const int m = m0 + c * 8 + e;
            dot0 = fmaf(k[e], m < p.Mz ? __ldg(p.alpha2 + (size_t)d * p.Mz + m) : 0.f, dot0);
            dot1 = fmaf(k[e + 1], m + 1 < p.Mz ? __ldg(p.alpha2 + (size_t)d * p.Mz + m + 1) : 0.f, dot1);
          }
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split2_bf16(k[2 * e], k[2 * e + 1], hi[e], lo[e]);
        const uint32_t off = (uint32_t)(((mh * 4 + c) ^ (n & 7)) << 4);
        *reinterpret_cast<uint4*>(img_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(img_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      ptx::fence_proxy_async();                      // generic-proxy stores -> visible to the tensor core's reads
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(a_full(s));
    }
    s_dot[mh * 128 + n] = dot0 + dot1;
    // epilogue: thread = point (TMEM lane), this warp's half of the 256 factor rows
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
    const int ew = warp - 2, qd = warp & 3, hc = ew >> 2;          // TMEM lane quarter, column half
    const uint32_t tl = tmem_base + ((uint32_t)(qd * 32) << 16);
    const bool has_w = p.MT - 4 * jt > 0;
    float pv = 0.f, pw = 0.f;
#pragma unroll 1
    for (int c0 = hc * 128; c0 < hc * 128 + 128; c0 += 16) {
      float v[16], w[16];
      ptx::tmem_ld16_wait(tl + (uint32_t)c0, v);
      ptx::tmem_ld16_wait(tl + 256u + (uint32_t)c0, w);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        pv = fmaf(v[i], v[i], pv);
        if (has_w) pw = fmaf(w[i], w[i], pw);
      }
    }
    const int pn = qd * 32 + lane;
    if (hc == 1) { s_red[pn * 3 + 0] = pv; s_red[pn * 3 + 1] = pw; }
    ptx::named_bar_sync(1, 256);
    if (hc == 0 && n0 + pn < p.n_pad) {
      float* r = p.partial + (((size_t)d * p.JT + jt) * p.n_pad + n0 + pn) * 3;
      r[0] = pv + s_red[pn * 3 + 0];
      r[1] = pw + s_red[pn * 3 + 1];
      r[2] = mean_cta ? s_dot[pn] + s_dot[128 + pn] : 0.f;        // mean - c, counted once (last row tile)
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// fp32 factor [D][Mp][Mp] (row-major rows j, columns m; zero padded) -> bf16 hi/lo k-block images
// [D][JT][MT][hi | lo][256 rows x 128 B], K-major, SWIZZLE_128B; rows / columns beyond Mp are zero.
__global__ void gp_tc_pack_kernel(const float* __restrict__ src, int Mp, int MT, int JT, uint8_t* __restrict__ dst) {
  const int kb = blockIdx.x, jt = blockIdx.y, d = blockIdx.z;
  uint8_t* img = dst + ((size_t)((size_t)d * JT + jt) * MT + kb) * (2u * GT_B_IMG);
  for (int u = threadIdx.x; u < 256 * 8; u += blockDim.x) {       // (row, 16-byte chunk)
    const int r = u >> 3, c = u & 7;
    const int j = jt * 256 + r, m0 = kb * 64 + c * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = (j < Mp && m0 + e < Mp) ? src[((size_t)d * Mp + j) * Mp + m0 + e] : 0.f;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split2_bf16(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
    const uint32_t off = sw128_offset((uint32_t)r, (uint32_t)c);
    *reinterpret_cast<uint4*>(img + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(img + GT_B_IMG + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// alpha2[d][m] = sum_j Linv[d][j][m] beta[d][j]  (= K_ZZ^-1 (m_q - c)), fp64 accumulate, once per weight load
__global__ void gp_tc_alpha_kernel(const float* __restrict__ linv, const float* __restrict__ beta, int Mp, float* __restrict__ alpha2) {
  const int d = blockIdx.y, m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mp) return;
  double acc = 0.0;
  for (int j = m; j < Mp; ++j) acc += (double)linv[((size_t)d * Mp + j) * Mp + m] * (double)beta[(size_t)d * Mp + j];
  alpha2[(size_t)d * Mp + m] = (float)acc;
}

bool gp_tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DVG_GP_TC");          // developer switch: 0 = FP32 FFMA tiles (gp_big_partial_kernel)
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

int gp_tc_pack(dvg_gp_s* h, cudaStream_t stream) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess || prop.major != 10) return DVG_OK;
  const int D = h->dims.num_dims, Mp = h->mp;
  const int JT = ceil_div(Mp, 256), MT = JT * 4;
  const size_t bytes = (size_t)D * JT * MT * 2 * GT_B_IMG;
  if (h->tc_img_bytes != bytes) {
    if (h->tc_img_v) h->retired.push_back(h->tc_img_v);
    if (h->tc_img_w) h->retired.push_back(h->tc_img_w);
    h->tc_img_v = h->tc_img_w = nullptr;
    h->tc_img_bytes = 0;
    if (cudaMalloc(&h->tc_img_v, bytes) != cudaSuccess || cudaMalloc(&h->tc_img_w, bytes) != cudaSuccess) {
      cudaGetLastError();                       // not enough memory for the packed copies: stay on the FP32 path
      if (h->tc_img_v) cudaFree(h->tc_img_v);
      h->tc_img_v = h->tc_img_w = nullptr;
      return DVG_OK;
    }
    h->tc_img_bytes = bytes;
  }
  if (!h->tc_alpha2 || h->tc_alpha2_n != (size_t)D * Mp) {
    if (h->tc_alpha2) h->retired.push_back(h->tc_alpha2);
    h->tc_alpha2 = nullptr;
    DVG_CUDA(cudaMalloc(&h->tc_alpha2, sizeof(float) * D * Mp));
    h->tc_alpha2_n = (size_t)D * Mp;
  }
  gp_tc_alpha_kernel<<<dim3(ceil_div(Mp, 128), D), 128, 0, stream>>>(h->linv, h->alpha, Mp, h->tc_alpha2);
  DVG_LAUNCH_CHECK();
  dim3 grid(MT, JT, D);
  gp_tc_pack_kernel<<<grid, 256, 0, stream>>>(h->linv, Mp, MT, JT, h->tc_img_v);
  DVG_LAUNCH_CHECK();
  gp_tc_pack_kernel<<<grid, 256, 0, stream>>>(h->lqt, Mp, MT, JT, h->tc_img_w);
  DVG_LAUNCH_CHECK();
  h->tc_JT = JT; h->tc_MT = MT;
  return DVG_OK;
}

// partial sums of n_rows points into h->partial ([D][JT][n_pad][3], n_pad a multiple of 128); caller runs the finalize kernel
int gp_tc_partial_launch(dvg_gp_s* h, int n_rows, int n_pad, const float* x, int ldx, const int32_t* row_index,
                         int want_mean, cudaStream_t stream) {
  GpTcArgs a{};
  a.want_mean = want_mean;
  a.n_rows = n_rows; a.n_pad = n_pad; a.MT = h->tc_MT; a.JT = h->tc_JT; a.ldx = ldx;
  a.x = x; a.row_index = row_index; a.z = h->z; a.Mz = h->mp; a.beta = h->alpha; a.hyp = h->hyp; a.alpha2 = h->tc_alpha2;
  a.img_v = h->tc_img_v; a.img_w = h->tc_img_w; a.partial = h->partial;
  const size_t smem = (size_t)GT_STAGES * GT_STAGE + sizeof(float) * (128 + 256 + 384) + 128;
  static bool configured = false;
  if (!configured) {
    DVG_CUDA(cudaFuncSetAttribute(gp_tc_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  dim3 grid(n_pad / 128, h->tc_JT, h->dims.num_dims);
  gp_tc_partial_kernel<<<grid, GT_THREADS, smem, stream>>>(a);
  DVG_LAUNCH_CHECK();
  return DVG_OK;
}

}  // namespace dvg
