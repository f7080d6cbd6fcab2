// Tensor-pipe / operand-ring probe for lstm_step_kernel (round 2): how fast can a cta_group::2 pair run the step
// kernel's MMA pattern (M = 256, N = 256, bf16 hi/lo split: 3 MMAs per 16-wide k-step) when
//   mode 0  the operands are resident in shared memory (no TMA at all)          -> the tensor pipe's own rate
//   mode 1  every stage is streamed from L2 through the kernel's ring protocol    -> the pipeline's rate
//           (both CTAs fill their half, the peer relays "landed" to the leader, tcgen05.commit frees the stage)
// with k-blocks of 64 elements (128-byte rows, SWIZZLE_128B, 64 KB stages) or 32 elements (64-byte rows,
// SWIZZLE_64B, 32 KB stages), with and without epilogue-like noise (16 warps doing tcgen05.ld + shared-memory
// transposes) and with all pairs or a single pair on the chip.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I dvg_b200/csrc -o scripts/_bin/mma_probe \
//        scripts/mma_probe.cu && scripts/_bin/mma_probe
// Output: one JSON line per configuration (us per 64-element k-block equivalent; nominal 0.78 us at 1.965 GHz).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "ptx.cuh"

using namespace dvg;

constexpr int MAX_STAGES = 8;
constexpr int EW = 16;
constexpr int THREADS = 64 + EW * 32;

struct ProbeArgs {
  int mode, stages, row_bytes, nparts, iters, noise, n_src_stages, direct;
  const uint8_t* src;
  unsigned* sink;
  unsigned long long* cycles;
};

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, int row_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * row_bytes) >> 4) << 32;      // 8-row groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(row_bytes == 128 ? 2 : 4) << 61;  // SWIZZLE_128B / SWIZZLE_64B
  return d;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1) mma_probe_kernel(const ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bars[3 * MAX_STAGES + 2];
  __shared__ uint32_t tmem_slot_s;
  __shared__ volatile int done_flag;
  const uint32_t base = ptx::smem_u32(smem);
  const uint32_t full0 = ptx::smem_u32(&bars[0]), empty0 = ptx::smem_u32(&bars[MAX_STAGES]),
                 pfull0 = ptx::smem_u32(&bars[2 * MAX_STAGES]), fin = ptx::smem_u32(&bars[3 * MAX_STAGES]);
  const uint32_t rank = ptx::cluster_ctarank();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t img = 128u * (uint32_t)p.row_bytes;           // one 128-row operand image
  const uint32_t stage_bytes = 2u * p.nparts * img;            // A (nparts images) + this CTA's half of B
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(full0 + 8 * s, 1);
      ptx::mbar_init(empty0 + 8 * s, 1);
      ptx::mbar_init(pfull0 + 8 * s, 1);
    }
    ptx::mbar_init(fin, 1);
    ptx::fence_barrier_init();
    done_flag = 0;
  }
  if (warp == 1) {
    ptx::tmem_alloc2(ptx::smem_u32(&tmem_slot_s), 512);
    ptx::tmem_relinquish2();
  }
  // define the operand bytes (finite bf16 values) so the tensor pipe does not chew on NaN payloads
  for (uint32_t i = threadIdx.x; i < (uint32_t)p.stages * stage_bytes / 4; i += THREADS)
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 0xff);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_slot_s;
  const int ks = p.row_bytes / 32;                             // 16-element k-steps per stage
  const long long t_start = clock64();
  if (warp == 0 && lane == 0 && p.mode == 1) {
    // producer (both CTAs)
    uint32_t leader_pfull0;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(leader_pfull0) : "r"(pfull0), "r"(0));
    for (int it = 0; it < p.iters; ++it) {
      const int s = it % p.stages;
      if (it >= p.stages) ptx::mbar_wait(empty0 + 8 * s, ((it / p.stages) - 1) & 1);
      const uint32_t dst = base + (uint32_t)s * stage_bytes;
      const uint8_t* g = p.src + (size_t)((blockIdx.x * 37 + it * 3) % p.n_src_stages) * stage_bytes;
      uint32_t bar = full0 + 8 * s;
      if (p.direct && rank == 1) {
        bar = leader_pfull0 + 8 * s;
        asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar),
                     "r"(stage_bytes)
                     : "memory");
      } else {
        ptx::mbar_expect_tx(bar, stage_bytes);
      }
      // as in the step kernel: B half (hi, lo) as separate copies, then the A images in one copy
      const uint32_t a_bytes = p.nparts * img;
      ptx::bulk_g2s(dst + a_bytes, g + a_bytes, img, bar);
      if (p.nparts == 2) ptx::bulk_g2s(dst + a_bytes + img, g + a_bytes + img, img, bar);
      ptx::bulk_g2s(dst, g, a_bytes, bar);
    }
  } else if (warp == 1 && lane == 0) {
    if (rank == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(256, 256);
      for (int it = 0; it < p.iters; ++it) {
        const int s = it % p.stages;
        if (p.mode == 1) {
          ptx::mbar_wait(full0 + 8 * s, (it / p.stages) & 1);
          ptx::mbar_wait(pfull0 + 8 * s, (it / p.stages) & 1);
          ptx::tc_fence_after();
        }
        const uint32_t sa = base + (uint32_t)s * stage_bytes;
        const uint64_t a_hi = make_desc(sa, p.row_bytes), a_lo = make_desc(sa + img, p.row_bytes);
        const uint64_t b_hi = make_desc(sa + p.nparts * img, p.row_bytes);
        const uint64_t b_lo = make_desc(sa + p.nparts * img + img, p.row_bytes);
        const int kb64 = it * p.row_bytes / 128;                // 64-element k-block index
        const uint32_t d_tmem = tmem_base + (uint32_t)(((kb64 / 8) & 1) * 256);
        for (int kk = 0; kk < ks; ++kk) {
          const uint64_t adv = (uint64_t)(kk * 2);
          const uint32_t accum = (kb64 % 8 == 0 && kk == 0 && (it * p.row_bytes) % 128 == 0) ? 0u : 1u;
          ptx::umma2_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc, accum);
          if (p.nparts == 2) {
            ptx::umma2_bf16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
            ptx::umma2_bf16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
          }
        }
        if (p.mode == 1) ptx::umma2_commit_mcast(empty0 + 8 * s, 3);
      }
      ptx::umma2_commit_mcast(fin, 3);
    } else if (p.mode == 1 && !p.direct) {
      for (int it = 0; it < p.iters; ++it) {
        const int s = it % p.stages;
        ptx::mbar_wait(full0 + 8 * s, (it / p.stages) & 1);
        ptx::mbar_arrive_remote(pfull0 + 8 * s, 0);
      }
    }
  } else if (warp >= 2 && p.noise) {
    // epilogue-like noise: TMEM reads of the idle accumulator half + a 32 x 128 B transpose through shared memory
    uint8_t* eb = smem + (size_t)p.stages * stage_bytes + (size_t)(warp - 2) * 2048;   // 32 rows x 64 B
    const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    int n = 0;
    while (!done_flag) {
      float v[16];
      ptx::tmem_ld16_wait(tl + ((n * 16) & 255), v);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<float4*>(eb + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) =
            make_float4(v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(eb + (lane ^ c) * 64 + ((c ^ (((lane ^ c) >> 1) & 3)) << 4));
        acc += t.x * 1.0001f + t.y + t.z + t.w;
      }
      __syncwarp();
      ++n;
    }
    if (acc == 1234.5f) p.sink[1] = 1;
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_wait(fin, 0);
    done_flag = 1;
    if (rank == 0) p.cycles[blockIdx.x / 2] = (unsigned long long)(clock64() - t_start);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc2(tmem_base, 512);
  }
}

static int run(const char* name, int mode, int stages, int row_bytes, int nparts, int noise, int pairs, int iters64,
               const uint8_t* src, size_t src_bytes, unsigned* sink, unsigned long long* cycles, int direct = 0) {
  ProbeArgs a{};
  a.mode = mode; a.stages = stages; a.row_bytes = row_bytes; a.nparts = nparts; a.noise = noise; a.direct = direct;
  a.iters = iters64 * (128 / row_bytes);
  const size_t stage_bytes = (size_t)2 * nparts * 128 * row_bytes;
  a.n_src_stages = (int)(src_bytes / stage_bytes);
  a.src = src; a.sink = sink; a.cycles = cycles;
  const size_t smem = stages * stage_bytes + (noise ? EW * 2048 : 0);
  cudaError_t ae = cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ae != cudaSuccess) { printf("{\"probe\": \"%s\", \"error\": \"set attribute: %s\"}\n", name, cudaGetErrorString(ae)); return 1; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  unsigned long long cyc[74] = {0};
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    mma_probe_kernel<<<pairs * 2, THREADS, smem>>>(a);
    cudaEventRecord(e1);
    cudaError_t se = cudaEventSynchronize(e1);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess || se != cudaSuccess) {
      printf("{\"probe\": \"%s\", \"error\": \"%s / %s\"}\n", name, cudaGetErrorString(le), cudaGetErrorString(se));
      return 1;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) {
      best = ms;
      cudaMemcpy(cyc, cycles, sizeof(unsigned long long) * pairs, cudaMemcpyDeviceToHost);
    }
  }
  unsigned long long cmax = 0, cmin = ~0ull;
  for (int i = 0; i < pairs; ++i) { if (cyc[i] > cmax) cmax = cyc[i]; if (cyc[i] < cmin) cmin = cyc[i]; }
  printf("{\"probe\": \"%s\", \"mode\": %d, \"stages\": %d, \"stage_KB\": %d, \"nparts\": %d, \"noise\": %d, \"pairs\": %d, "
         "\"kblocks64\": %d, \"ms\": %.4f, \"us_per_kblock64\": %.3f, \"cycles_per_kblock64_min\": %.0f, "
         "\"cycles_per_kblock64_max\": %.0f, \"fill_TBps\": %.2f}\n",
         name, mode, stages, (int)(stage_bytes / 1024), nparts, noise, pairs, iters64, best, best * 1e3 / iters64,
         (double)cmin / iters64, (double)cmax / iters64,
         mode == 1 ? (double)pairs * 2 * a.iters * stage_bytes / (best * 1e-3) / 1e12 : 0.0);
  fflush(stdout);
  return 0;
}

int main(int argc, char** argv) {
  const int iters64 = argc > 1 ? atoi(argv[1]) : 2000;
  const int with_direct = argc > 2 ? atoi(argv[2]) : 0;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int pairs = prop.multiProcessorCount / 2;
  const size_t src_bytes = (size_t)24 << 20;              // 24 MB of operand images: L2 resident, like the step kernel's
  uint8_t* src;
  unsigned* sink;
  unsigned long long* cycles;
  cudaMalloc(&src, src_bytes);
  cudaMalloc(&sink, 16);
  cudaMalloc(&cycles, 8 * 128);
  cudaMemset(src, 0x3c, src_bytes);
  int rc = 0;
  // tensor pipe alone
  rc |= run("resident_x3", 0, 3, 128, 2, 0, pairs, iters64, src, src_bytes, sink, cycles);
  rc |= run("resident_x1", 0, 3, 128, 1, 0, pairs, iters64, src, src_bytes, sink, cycles);
  rc |= run("resident_x3_noise", 0, 3, 128, 2, 1, pairs, iters64, src, src_bytes, sink, cycles);
  rc |= run("resident_x3_sw64", 0, 6, 64, 2, 0, pairs, iters64, src, src_bytes, sink, cycles);
  // the ring, all pairs
  rc |= run("stream_3x64K", 1, 3, 128, 2, 0, pairs, iters64, src, src_bytes, sink, cycles);
  rc |= run("stream_6x32K", 1, 6, 64, 2, 0, pairs, iters64, src, src_bytes, sink, cycles);
  rc |= run("stream_3x64K_noise", 1, 3, 128, 2, 1, pairs, iters64, src, src_bytes, sink, cycles);
  rc |= run("stream_6x32K_noise", 1, 6, 64, 2, 1, pairs, iters64, src, src_bytes, sink, cycles);
  rc |= run("stream_5x32K_noise", 1, 5, 64, 2, 1, pairs, iters64, src, src_bytes, sink, cycles);
  rc |= run("stream_x1_6x32K", 1, 6, 128, 1, 0, pairs, iters64, src, src_bytes, sink, cycles);
  rc |= run("stream_2x64K", 1, 2, 128, 2, 0, pairs, iters64, src, src_bytes, sink, cycles);
  // the ring, one pair on the chip (no L2 contention): latency bound part
  rc |= run("stream_3x64K_1pair", 1, 3, 128, 2, 0, 1, iters64, src, src_bytes, sink, cycles);
  rc |= run("stream_6x32K_1pair", 1, 6, 64, 2, 0, 1, iters64, src, src_bytes, sink, cycles);
  rc |= run("resident_x3_1pair", 0, 3, 128, 2, 0, 1, iters64, src, src_bytes, sink, cycles);
  if (with_direct) {
    // the peer's copies complete on the LEADER's barrier (no relay); last: a fault here kills the context
    rc |= run("stream_3x64K_direct", 1, 3, 128, 2, 0, pairs, iters64, src, src_bytes, sink, cycles, 1);
    rc |= run("stream_6x32K_direct", 1, 6, 64, 2, 0, pairs, iters64, src, src_bytes, sink, cycles, 1);
  }
  return rc;
}
