// Bookkeeping kernels for the N-diverse-futures rollout.
//
// dvg_rollout_score: per (rollout s, sequence b) mean squared error of the generated latent sequence against a
// target latent sequence -- the device-side scoring pass of the best-of-N selection (the reference scores on
// the host after a D2H copy of every frame, generate_frames.py:175-178,185-190).  One pass over the [T, S*B, G]
// tensor: HBM bound, algorithmic bytes = T*S*B*G*4 (+ the small target).
#include "internal.cuh"

namespace dvg {

// One warp per row (s*B + b): lanes stride over g, loop over t; 8 rows per CTA.
__global__ void __launch_bounds__(256) rollout_score_kernel(int T, int S, int B, int G, const float* __restrict__ out,
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

int rollout_score_launch(int T, int S, int B, int G, const float* out, const float* target, float* scores,
                         cudaStream_t stream) {
  rollout_score_kernel<<<ceil_div(S * B, 8), 256, 0, stream>>>(T, S, B, G, out, target, scores);
  DVG_LAUNCH_CHECK();
  return DVG_OK;
}

}  // namespace dvg
