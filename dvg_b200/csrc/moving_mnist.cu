// Bouncing-digit sequences generated on the device (reference data/moving_mnist.py:38-91,
// MovingMNIST.__getitem__): keeps the SM-MNIST configuration free of CPU data generation and H2D copies.
//
// The reference draws its randomness lazily from np.random (digit index, start position, velocity, and two fresh
// velocity components at every bounce when deterministic=False, which is how utils.load_dataset builds it,
// utils.py:30-43).  Here the k-th np.random.randint(lo, hi) call of a sample is `lo + draws[sample][k] % (hi - lo)`
// of a caller-supplied stream of raw 32-bit integers, consumed in exactly the reference's call order (digit 0
// completely, then digit 1, ...), so the same stream scripted into the reference reproduces the frames bit for bit.
//
//   traj kernel   one warp per sample: stream staged in shared memory, lane 0 walks the n_digits x n_frames steps
//   render kernel one warp per (frame, sample), warp-stride over a grid of 8 CTAs per SM:
//                 out[t][b][0][y][x] = min(1, sum_n digit_n[y - sy_n][x - sx_n]), 128-bit streaming stores; HBM-write
//                 bound (4 B per pixel), the digit bank stays in L1/L2.
#include "internal.cuh"

namespace dvg {
namespace {

constexpr int DIGIT = 32;            // data/moving_mnist.py:15 (and the hard-coded 32 of :57-81)
constexpr int MAX_DIGITS = 8;

__device__ __forceinline__ int draw(const uint32_t* s, int& k, int lo, int hi) {
  return lo + (int)(s[k++] % (uint32_t)(hi - lo));
}

// traj[b][n] = { digit index, (sx, sy) x n_frames }
__global__ void __launch_bounds__(128) mnist_traj_kernel(int B, int T, int W, int n_digits, int deterministic, int n_bank,
                                                         const uint32_t* __restrict__ draws, int draws_per_seq,
                                                         int32_t* __restrict__ traj) {
  extern __shared__ uint32_t s_draws[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= B) return;
  uint32_t* s = s_draws + (size_t)warp * draws_per_seq;
  for (int i = lane; i < draws_per_seq; i += 32) s[i] = draws[(size_t)b * draws_per_seq + i];
  __syncwarp();
  if (lane != 0) return;
  int k = 0;
  const int span = W - DIGIT;
  const int rec = 1 + 2 * T;
  for (int n = 0; n < n_digits; ++n) {
    int32_t* out = traj + ((size_t)b * n_digits + n) * rec;
    out[0] = draw(s, k, 0, n_bank);                     // :47 idx = randint(N)
    int sx = draw(s, k, 0, span);                       // :50
    int sy = draw(s, k, 0, span);                       // :51
    int dx = draw(s, k, -4, 5);                         // :52
    int dy = draw(s, k, -4, 5);                         // :53
    for (int t = 0; t < T; ++t) {
      if (sy < 0) {                                     // :55-61
        sy = 0;
        if (deterministic) dy = -dy;
        else { dy = draw(s, k, 1, 5); dx = draw(s, k, -4, 5); }
      } else if (sy >= span) {                          // :62-68
        sy = span - 1;
        if (deterministic) dy = -dy;
        else { dy = draw(s, k, -4, 0); dx = draw(s, k, -4, 5); }
      }
      if (sx < 0) {                                     // :70-76
        sx = 0;
        if (deterministic) dx = -dx;
        else { dx = draw(s, k, 1, 5); dy = draw(s, k, -4, 5); }
      } else if (sx >= span) {                          // :77-83
        sx = span - 1;
        if (deterministic) dx = -dx;
        else { dx = draw(s, k, -4, 0); dy = draw(s, k, -4, 5); }
      }
      out[1 + 2 * t] = sx;
      out[2 + 2 * t] = sy;
      sy += dy;                                         // :86-87
      sx += dx;
    }
  }
}

// One warp per frame, warp-stride loop over the T*B frames (grid sized to the SM count): no CTA barriers, every lane
// streams 128-bit stores; the n_digits (index, sx, sy) triples are broadcast loads hoisted out of the pixel loop.
template <int ND>
__global__ void __launch_bounds__(256) mnist_render_kernel(int B, int T, int W, int n_digits,
                                                           const float* __restrict__ bank,
                                                           const int32_t* __restrict__ traj, float* __restrict__ frames) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const long long n_frames = (long long)B * T;
  const int quads = W * W / 4;                           // W % 4 == 0 (checked by the host)
  const int rec = 1 + 2 * T;
  for (long long f = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); f < n_frames; f += (long long)gridDim.x * wpb) {
    const int t = (int)(f / B), b = (int)(f - (long long)t * B);
    int pi[ND], px[ND], py[ND];
#pragma unroll
    for (int n = 0; n < ND; ++n) {
      pi[n] = 0; px[n] = 0; py[n] = -2 * DIGIT;          // unused slots never cover a pixel
      if (n < n_digits) {
        const int32_t* tr = traj + ((size_t)b * n_digits + n) * rec;
        pi[n] = __ldg(tr);
        px[n] = __ldg(tr + 1 + 2 * t);
        py[n] = __ldg(tr + 2 + 2 * t);
      }
    }
    float4* out = reinterpret_cast<float4*>(frames + (size_t)f * W * W);
    for (int q = lane; q < quads; q += 32) {
      const int y = (q * 4) / W, x0 = (q * 4) - y * W;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int n = 0; n < ND; ++n) {                     // accumulation order of the reference's `+=` over digits (:85)
        const int dyy = y - py[n], dx0 = x0 - px[n];
        if ((unsigned)dyy >= (unsigned)DIGIT || dx0 <= -4 || dx0 >= DIGIT) continue;   // quad outside this digit
        const float* d = bank + ((size_t)pi[n] * DIGIT + dyy) * DIGIT;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int dxx = dx0 + j;
          if ((unsigned)dxx < (unsigned)DIGIT) v[j] += __ldg(d + dxx);
        }
      }
      float4 o;                                          // x[x > 1] = 1 (:89)
      o.x = v[0] > 1.f ? 1.f : v[0];
      o.y = v[1] > 1.f ? 1.f : v[1];
      o.z = v[2] > 1.f ? 1.f : v[2];
      o.w = v[3] > 1.f ? 1.f : v[3];
      __stcs(out + q, o);                                // written once, read later by another kernel: streaming store
    }
  }
}

}  // namespace

int moving_mnist_launch(int B, int T, int W, int n_digits, int deterministic, const float* bank, int n_bank,
                        const uint32_t* draws, int draws_per_seq, int32_t* traj, float* frames, cudaStream_t stream) {
  if (n_digits > MAX_DIGITS) { set_error("at most %d digits per frame", MAX_DIGITS); return DVG_ERR_ARG; }
  const int warps = 4;
  const size_t smem = sizeof(uint32_t) * (size_t)warps * draws_per_seq;
  if (smem > 48 * 1024) { set_error("draw stream too long for the trajectory kernel (%d words)", draws_per_seq); return DVG_ERR_ARG; }
  mnist_traj_kernel<<<(B + warps - 1) / warps, warps * 32, smem, stream>>>(B, T, W, n_digits, deterministic, n_bank, draws,
                                                                           draws_per_seq, traj);
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long want = ((long long)B * T + 7) / 8;
  const int grid = (int)(want < (long long)sms * 8 ? want : (long long)sms * 8);
  if (n_digits <= 2) mnist_render_kernel<2><<<grid, 256, 0, stream>>>(B, T, W, n_digits, bank, traj, frames);
  else if (n_digits <= 4) mnist_render_kernel<4><<<grid, 256, 0, stream>>>(B, T, W, n_digits, bank, traj, frames);
  else mnist_render_kernel<MAX_DIGITS><<<grid, 256, 0, stream>>>(B, T, W, n_digits, bank, traj, frames);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("moving_mnist launch: %s", cudaGetErrorString(e)); return DVG_ERR_CUDA; }
  return DVG_OK;
}

}  // namespace dvg
