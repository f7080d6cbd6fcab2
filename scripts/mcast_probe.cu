// L2 -> shared-memory operand streaming with and without cluster multicast: the measurement behind DESIGN.md §7
// item 1 (the step kernel's MMA phases run at the ~9.5 TB/s L2 -> SM rate: 64 KB per k-block and CTA, of which the
// 32 KB activation half is the same for the CTAs that work on the N tiles of one row group).
//
// Every CTA streams ITERS stages of 64 KB through a 3-stage ring, like lstm_step_kernel's producer:
//   private half (32 KB)   cp.async.bulk by the CTA itself                      (the weight images)
//   shared half  (32 KB)   CLUSTER == 1: cp.async.bulk by the CTA itself
//                          CLUSTER  > 1: each CTA of the cluster multicasts its 1/CLUSTER slice to all of them
// A stage is "consumed" by one warp reading a few words of it; with multicast a stage may only be refilled once every
// CTA of the cluster has released it (remote arrives on each CTA's empty barrier).  Sources stay L2 resident
// (footprint well below 126 MB), so the result is the L2 -> SM path, not HBM.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I dvg_b200/csrc -o /tmp/mcast_probe \
//        scripts/mcast_probe.cu && /tmp/mcast_probe
// Output: one JSON line per cluster size with the smem fill rate and the L2 read rate.  mbarrier waits trap after ~2 s
// instead of hanging (ptx.cuh), so a protocol bug shows up as a CUDA error.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace dvg;

constexpr int STAGES = 3;
constexpr int HALF = 32 * 1024;
constexpr int STAGE_BYTES = 2 * HALF;
constexpr int THREADS = 128;          // warp 0: producer lane, warp 1: consumer, rest idle

template <int CLUSTER>
__global__ void __launch_bounds__(THREADS, 1) probe_kernel(const uint8_t* __restrict__ shared_src,
                                                           const uint8_t* __restrict__ private_src, int n_shared_chunks,
                                                           int n_private_chunks, int iters, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bars[2 * STAGES];
  const uint32_t full0 = ptx::smem_u32(&bars[0]), empty0 = ptx::smem_u32(&bars[STAGES]);
  const uint32_t rank = CLUSTER > 1 ? ptx::cluster_ctarank() : 0;
  const uint32_t cluster = CLUSTER > 1 ? ptx::cluster_id_x() : blockIdx.x;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full0 + 8 * s, 1);              // producer's arrive.expect_tx
      ptx::mbar_init(empty0 + 8 * s, CLUSTER);       // one release per CTA of the cluster
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (CLUSTER > 1) ptx::cluster_sync_all();          // every CTA's barriers exist before any remote arrive / multicast
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) ptx::mbar_wait(empty0 + 8 * s, ((it / STAGES) - 1) & 1);
      const uint32_t full = full0 + 8 * s;
      const uint32_t dst = ptx::smem_u32(smem + (size_t)s * STAGE_BYTES);
      ptx::mbar_expect_tx(full, STAGE_BYTES);        // the bytes THIS CTA will receive (own copies + peers' slices)
      const uint8_t* priv = private_src + (size_t)((blockIdx.x * 7 + it) % n_private_chunks) * HALF;
      ptx::bulk_g2s(dst + HALF, priv, HALF, full);
      const uint8_t* sh = shared_src + (size_t)((cluster * 5 + it) % n_shared_chunks) * HALF;
      if (CLUSTER == 1) {
        ptx::bulk_g2s(dst, sh, HALF, full);
      } else {
        constexpr int SLICE = HALF / CLUSTER;
        ptx::bulk_g2s_mcast(dst + rank * SLICE, sh + rank * SLICE, SLICE, full, (uint16_t)((1u << CLUSTER) - 1));
      }
    }
  } else if (warp == 1) {
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      ptx::mbar_wait(full0 + 8 * s, (it / STAGES) & 1);
      const unsigned* p = reinterpret_cast<const unsigned*>(smem + (size_t)s * STAGE_BYTES);
      acc += p[lane] + p[HALF / 4 + lane] + p[STAGE_BYTES / 4 - 32 + lane];
      __syncwarp();
      if (lane == 0) {
        if (CLUSTER == 1) ptx::mbar_arrive(empty0 + 8 * s);
        else
          for (int c = 0; c < CLUSTER; ++c) ptx::mbar_arrive_remote(empty0 + 8 * s, c);
      }
    }
    if (acc == 0x12345678u) sink[0] = acc;           // keep the loads
  }
  __syncthreads();
  if (CLUSTER > 1) ptx::cluster_sync_all();          // nobody exits while peers may still arrive on its barriers
}

template <int CLUSTER>
static int run(const uint8_t* shared_src, const uint8_t* private_src, int n_sh, int n_pr, int sms, int iters,
               unsigned* sink) {
  const int grid = sms / CLUSTER * CLUSTER;
  const size_t smem = (size_t)STAGES * STAGE_BYTES;
  cudaFuncSetAttribute(probe_kernel<CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (CLUSTER > 8) cudaFuncSetAttribute(probe_kernel<CLUSTER>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  cudaOccupancyMaxActiveClusters(&max_clusters, probe_kernel<CLUSTER>, &cfg);
  if (max_clusters * CLUSTER < grid) {               // not every GPC fits clusters of this size on all its SMs
    cfg.gridDim = dim3(max_clusters * CLUSTER);
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    cudaError_t le = cudaLaunchKernelEx(&cfg, probe_kernel<CLUSTER>, shared_src, private_src, n_sh, n_pr, iters, sink);
    cudaEventRecord(e1);
    cudaError_t se = cudaEventSynchronize(e1);
    if (le != cudaSuccess || se != cudaSuccess) {
      printf("{\"cluster\": %d, \"error\": \"%s / %s\"}\n", CLUSTER, cudaGetErrorString(le), cudaGetErrorString(se));
      return 1;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;              // first launch warms L2
  }
  const double ctas = cfg.gridDim.x;
  const double fill = ctas * iters * (double)STAGE_BYTES;
  const double l2 = ctas * iters * ((double)HALF + (double)HALF / CLUSTER);
  printf("{\"cluster\": %d, \"ctas\": %d, \"max_active_clusters\": %d, \"iters\": %d, \"ms\": %.4f, "
         "\"smem_fill_TBps\": %.2f, \"l2_read_TBps\": %.2f, \"us_per_stage\": %.3f}\n",
         CLUSTER, (int)ctas, max_clusters, iters, best, fill / (best * 1e-3) / 1e12, l2 / (best * 1e-3) / 1e12,
         best * 1e3 / iters);
  return 0;
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  const int n_sh = 640, n_pr = 256;                  // 20 MB + 8 MB of sources: L2 resident
  uint8_t *sh, *pr;
  unsigned* sink;
  cudaMalloc(&sh, (size_t)n_sh * HALF);
  cudaMalloc(&pr, (size_t)n_pr * HALF);
  cudaMalloc(&sink, 4);
  cudaMemset(sh, 1, (size_t)n_sh * HALF);
  cudaMemset(pr, 2, (size_t)n_pr * HALF);
  int rc = 0;
  rc |= run<1>(sh, pr, n_sh, n_pr, sms, iters, sink);
  rc |= run<2>(sh, pr, n_sh, n_pr, sms, iters, sink);
  rc |= run<4>(sh, pr, n_sh, n_pr, sms, iters, sink);
  rc |= run<8>(sh, pr, n_sh, n_pr, sms, iters, sink);
  return rc;
}
