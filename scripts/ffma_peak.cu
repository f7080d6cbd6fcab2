// FP32 FMA peak of the device ("fp32 CUDA-core peak not measured by the driver -- the builder measures
// it"): 8 independent FFMA chains per thread, 1024 threads per CTA, 8 CTAs per SM's worth of grid.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ffma_peak scripts/ffma_peak.cu && /tmp/ffma_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024) ffma_kernel(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount, ctas = sms * 8, thr = 1024, iters = 4096;
  float* out;
  cudaMalloc(&out, sizeof(float) * ctas * thr);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  ffma_kernel<<<ctas, thr>>>(out, iters, 0.999f, 0.001f);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    ffma_kernel<<<ctas, thr>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double flops = 2.0 * 8 * 16 * (double)iters * ctas * thr;
  printf("{\"device\": \"%s\", \"sms\": %d, \"ffma_tflops\": %.2f, \"ms\": %.3f, \"err\": \"%s\"}\n", prop.name, sms,
         flops / (best * 1e-3) / 1e12, best, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
