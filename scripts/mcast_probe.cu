// Operand-ring probes for lstm_step_kernel (results: profiles/r01_mcast_probe.md).
//
// Part 1 -- L2 -> shared-memory streaming with and without cluster multicast.  Every CTA streams ITERS stages of
// 64 KB through a 3-stage ring, like the step kernel's producer:
//   private half (32 KB)   cp.async.bulk by the CTA itself                      (the weight images)
//   shared half  (32 KB)   CLUSTER == 1: cp.async.bulk by the CTA itself
//                          CLUSTER  > 1: each CTA of the cluster multicasts its 1/CLUSTER slice to all of them
// A stage is "consumed" by one warp reading a few words of it; with multicast a stage may only be refilled once every
// CTA of the cluster has released it (remote arrives on each CTA's empty barrier).  Sources stay L2 resident
// (footprint well below 126 MB), so the result is the L2 -> SM path, not HBM.  Measured (round 1): 0.48 us per stage
// = 20 TB/s without multicast, 1.07 / 1.42 / 2.09 us with clusters of 2 / 4 / 8.
//
// Part 2 -- the cta_group::2 pair protocol in isolation (pair_kernel below; not yet run): relay hop vs completing the
// peer's copies on the leader's barrier, with and without emulated tensor work.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I dvg_b200/csrc -o /tmp/mcast_probe \
//        scripts/mcast_probe.cu && /tmp/mcast_probe [iters]
// Output: one JSON line per configuration.  mbarrier waits trap after ~2 s instead of hanging (ptx.cuh), so a protocol
// bug shows up as a CUDA error.
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

// The step kernel's pair protocol in isolation: both CTAs of a cta_group::2 pair fill their own 64 KB stage, ONE
// consumer (the leader's MMA issuer) needs both halves, "computes" for `busy` cycles (0.78 us of tensor work per k-block
// in the real kernel) and releases the stage to both producers (tcgen05.commit multicast in the real kernel, two
// arrives here).
//   DIRECT == 0   as lstm_step_kernel does it today: the peer's relay warp waits for its own full barrier and then
//                 arrives on the leader's `pfull` barrier
//   DIRECT == 1   the peer's bulk copies complete on the LEADER's pfull barrier (remote expect_tx + remote mbarrier
//                 operand), no relay hop -- if the hardware rejects a completion barrier outside the destination CTA
//                 this variant traps / reports a CUDA error, which is the answer too
template <int DIRECT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
    pair_kernel(const uint8_t* __restrict__ src, int n_chunks, int iters, int busy, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bars[3 * STAGES];
  const uint32_t full0 = ptx::smem_u32(&bars[0]), empty0 = ptx::smem_u32(&bars[STAGES]),
                 pfull0 = ptx::smem_u32(&bars[2 * STAGES]);
  const uint32_t rank = ptx::cluster_ctarank();
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full0 + 8 * s, 1);
      ptx::mbar_init(empty0 + 8 * s, 1);             // released by the leader's consumer
      ptx::mbar_init(pfull0 + 8 * s, 1);             // relay arrive, or the peer's remote expect_tx
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  ptx::cluster_sync_all();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    uint32_t leader_pfull0;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(leader_pfull0) : "r"(pfull0), "r"(0));
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) ptx::mbar_wait(empty0 + 8 * s, ((it / STAGES) - 1) & 1);
      const uint32_t dst = ptx::smem_u32(smem + (size_t)s * STAGE_BYTES);
      const uint8_t* g = src + (size_t)((blockIdx.x * 7 + it) % n_chunks) * STAGE_BYTES;
      if (DIRECT && rank == 1) {
        const uint32_t bar = leader_pfull0 + 8 * s;
        asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar),
                     "r"((uint32_t)STAGE_BYTES)
                     : "memory");
        ptx::bulk_g2s(dst, g, HALF, bar);
        ptx::bulk_g2s(dst + HALF, g + HALF, HALF, bar);
      } else {
        ptx::mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
        ptx::bulk_g2s(dst, g, HALF, full0 + 8 * s);
        ptx::bulk_g2s(dst + HALF, g + HALF, HALF, full0 + 8 * s);
      }
    }
  } else if (warp == 1 && lane == 0) {
    if (rank == 0) {
      unsigned acc = 0;
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        ptx::mbar_wait(full0 + 8 * s, (it / STAGES) & 1);
        ptx::mbar_wait(pfull0 + 8 * s, (it / STAGES) & 1);
        acc += reinterpret_cast<const unsigned*>(smem + (size_t)s * STAGE_BYTES)[it & 1023];
        const long long t0 = clock64();
        while (clock64() - t0 < busy) {}
        ptx::mbar_arrive_remote(empty0 + 8 * s, 0);
        ptx::mbar_arrive_remote(empty0 + 8 * s, 1);
      }
      if (acc == 0x12345678u) sink[0] = acc;
    } else if (!DIRECT) {
      for (int it = 0; it < iters; ++it) {           // relay: "my stage landed" -> leader's pfull
        const int s = it % STAGES;
        ptx::mbar_wait(full0 + 8 * s, (it / STAGES) & 1);
        ptx::mbar_arrive_remote(pfull0 + 8 * s, 0);
      }
    }
  }
  __syncthreads();
  ptx::cluster_sync_all();
}

template <int DIRECT>
static int run_pair(const uint8_t* src, int n_chunks, int sms, int iters, int busy, unsigned* sink) {
  const int grid = sms / 2 * 2;
  const size_t smem = (size_t)STAGES * STAGE_BYTES;
  cudaFuncSetAttribute(pair_kernel<DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    pair_kernel<DIRECT><<<grid, THREADS, smem>>>(src, n_chunks, iters, busy, sink);
    cudaEventRecord(e1);
    cudaError_t se = cudaEventSynchronize(e1);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess || se != cudaSuccess) {
      printf("{\"pair\": \"%s\", \"busy_cycles\": %d, \"error\": \"%s / %s\"}\n", DIRECT ? "direct" : "relay", busy,
             cudaGetErrorString(le), cudaGetErrorString(se));
      return 1;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  printf("{\"pair\": \"%s\", \"busy_cycles\": %d, \"ctas\": %d, \"iters\": %d, \"ms\": %.4f, \"us_per_stage\": %.3f, "
         "\"smem_fill_TBps\": %.2f}\n",
         DIRECT ? "direct" : "relay", busy, grid, iters, best, best * 1e3 / iters,
         (double)grid * iters * STAGE_BYTES / (best * 1e-3) / 1e12);
  return 0;
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
  // the pair protocol, without and with 0.78 us (1530 cycles at 1.965 GHz) of emulated tensor work per stage; the
  // direct variant last: if it faults, the context is gone
  const int n_st = n_sh / 2;                          // 64 KB chunks in the `sh` buffer
  rc |= run_pair<0>(sh, n_st, sms, iters, 0, sink);
  rc |= run_pair<0>(sh, n_st, sms, iters, 1530, sink);
  rc |= run_pair<1>(sh, n_st, sms, iters, 0, sink);
  rc |= run_pair<1>(sh, n_st, sms, iters, 1530, sink);
  return rc;
}
