// One LSTM time step (models/lstm.py:65-72) for SMALL batches -- at most 64 rows, the reference's own CPU-runnable
// case (BASELINE configs[0]: batch 16) and the batch-64 case (configs[3]) -- in ONE launch of ONE 16-CTA cluster.
//
// With so few rows a [rows x K] x [K x 4H] GEMM has no M dimension to tile: a 128-row tcgen05 tile would be 7/8
// padding and a single CTA would have to stream all 3.6 MB of (bf16 hi/lo) weights by itself (~26 us).  So the
// product is transposed and split over the gate columns instead (hard part 5):
//
//     D^T [128 gate rows x N batch rows]  =  W_tile [128 x K]  .  X^T [K x N]          N = rows rounded up to 16
//
//   * the WEIGHTS are the UMMA A operand: CTA c of the cluster owns the 128 gate rows {i,f,g,o} x 32 hidden units
//     of tile (c & 7) of layer (c >> 3) and streams only its own 1/16 of the weights (L2 -> smem ring, 32 KB stages);
//   * the ACTIVATIONS are the UMMA B operand (N <= 64 rows, K-major: the same packed k-block images the large-batch
//     kernels use, first N rows), resident in shared memory for the whole step;
//   * layer 1's recurrent half runs on CTAs 8-15 WHILE layer 0 runs on CTAs 0-7; h'_0 then travels to the layer-1
//     CTAs through distributed shared memory (16-byte remote stores into their B-operand images + a remote
//     mbarrier arrive), never through global memory; h'_1 reaches the head (CTA 0, idle by then) the same way;
//   * accumulators live in TMEM with lane = gate row, column = batch row; the four epilogue warps each own one gate
//     (TMEM lane quarter), exchange the pre-activations through shared memory and update 32 units x N cells.
//
// bf16x3 (fp32-grade: hi*hi + hi*lo + lo*hi, fp32 accumulate) and bf16 variants; H = 256, L = 2 (the reference's
// sizes; other shapes take the one-launch-per-GEMM path of lstm_tc.cu).
#include <stdlib.h>

#include "gp_rsample.cuh"
#include "gp_trigger.cuh"
#include "tc_common.cuh"

namespace dvg {

constexpr int SM_CL = 16;                 // cluster size
constexpr int SM_THREADS = 224;           // warp 0 producer, warp 1 TMEM + MMA issuer, warps 2-5 epilogue, warp 6 trigger finaliser
constexpr int SM_TRIG_MAX_S = 2;          // rollouts per launch the fused trigger handles (<= 64 rows: S = 1 in the reference's configs)
constexpr int SM_TRIG_DIMS = 8;           // latent dims per CTA: ceil(D / 16) <= 8, i.e. D <= 128
constexpr int SM_STAGE = 2 * TC_A_IMG;    // one weight k-block: hi + lo image of 128 gate rows
constexpr int SM_MAX_STAGES = 6;
constexpr int SM_H = 256, SM_HK = 4;

struct SmallArgs {
  int rows, N, G, ldx, ldy, n_valid, nparts, rows_per_flag, kbx, x_ksteps, stages, head_rows, debug_stop;
  const float* x; const float* h_in; const float* c_in; float* h_out; float* c_out;
  const uint8_t* hp_in; uint8_t* hp_out;          // packed state images [L][1][4][2][16 KB]
  const uint8_t* w[2]; const uint8_t* wh;         // packed weights: layer 0 (embed folded), layer 1, head
  const float* b[2]; const float* bh;
  float* y; const uint8_t* hold;
  unsigned long long* trace;      // developer timeline (DVG_SMALL_TRACE=1): [16 CTAs][32 slots] of %globaltimer
  // fused GP variance trigger (generate_frames.py:227-232,275,283-289); enabled = 0: plain LSTM step
  struct {
    int enabled, S, D, W, warmup;
    float factor;
    const int32_t* stat_rows;
    const float* z; const float* linv; const float* lqt; const float* hyp;     // M = 40 factors of dvg_gp_prepare
    float* var_rows; float* window; int32_t* count; float* value; float* thr; uint8_t* mask; int* trig_list; int* trig_count;
    const float* rs_eps; const float* alpha; int n_points;     // non-null rs_eps: fired rollouts are resampled in this launch
  } trig;
};

__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster16(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_addr(uint32_t remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
// cluster-scope acquire wait (the arrivals come from other CTAs)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) {
      printf("dvg_b200: cluster mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

#define STRACE(slot)                                                                       \
  do {                                                                                     \
    if (p.trace) {                                                                         \
      unsigned long long _t;                                                               \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t)::"memory");                     \
      p.trace[(size_t)blockIdx.x * 32 + (slot)] = _t;                                      \
    }                                                                                      \
  } while (0)

__global__ void __launch_bounds__(SM_THREADS, 1) lstm_small_kernel(const __grid_constant__ SmallArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = ptx::smem_u32(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t c = ptx::cluster_ctarank();
  const int layer = (int)(c >> 3), tile = (int)(c & 7);
  const int N = p.N;
  const uint32_t img = (uint32_t)N * 128u;                 // one activation image part: N rows x 128 B
  const uint32_t nparts = (uint32_t)p.nparts;
  // shared memory map: [weight ring][8 activation k-block slots x (hi, lo)][gate exchange 4x16x32 f32][h' sub-image]
  //                    [bias 128 f32][head bias 128 f32][barriers]
  const uint32_t act0 = base + (uint32_t)p.stages * SM_STAGE;
  auto act = [&](int slot, int part) { return act0 + (uint32_t)(slot * 2 + part) * img; };
  uint8_t* tail = smem_raw + (size_t)p.stages * SM_STAGE + (size_t)16 * img;
  float* s_gate = reinterpret_cast<float*>(tail);                       // [4 gates][16 rows][32 units]
  uint16_t* s_h = reinterpret_cast<uint16_t*>(tail + 8192);             // [2 parts][N rows][32 units] bf16
  float* s_bias = reinterpret_cast<float*>(tail + 8192 + 128 * (size_t)N);
  float* s_bias_h = s_bias + 128;
  const uint32_t bar0 = ptx::smem_u32(s_bias_h + 128);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (SM_MAX_STAGES + s); };
  const uint32_t bar_act = bar0 + 8u * (2 * SM_MAX_STAGES);       // own activation images landed (TMA)
  const uint32_t bar_x = bar_act + 8;                              // x packed by the epilogue warps (4 warp arrivals)
  const uint32_t bar_in = bar_act + 16;                            // h' of the producing layer arrived through DSMEM (8 CTAs x 4 warps)
  const uint32_t bar_acc = bar_act + 24;                           // LSTM accumulator complete
  const uint32_t bar_acc2 = bar_act + 32;                          // head accumulator complete
  const uint32_t bar_trig = bar_act + 40;                          // (last CTA) partial variances of all 16 CTAs are in global memory
  const uint32_t bar_mask = bar_act + 48;                          // the decision of this step arrived (DSMEM store by the finaliser)
  const uint32_t tmem_slot = bar_act + 56;
  volatile int* s_mask = reinterpret_cast<volatile int*>(s_bias_h + 128 + 64);     // [SM_TRIG_MAX_S] fired flags, after the 256 B of barrier slots

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    ptx::mbar_init(bar_act, 1);
    ptx::mbar_init(bar_x, 4);
    ptx::mbar_init(bar_in, 32);
    ptx::mbar_init(bar_acc, 1);
    ptx::mbar_init(bar_acc2, 1);
    ptx::mbar_init(bar_trig, SM_CL);
    ptx::mbar_init(bar_mask, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  // this tile's gate biases, pre-scaled for lstm_cell_fast (-log2e for i, f, o; -2 log2e for g); head bias for tanh
  if (threadIdx.x < 128) {
    const int g = threadIdx.x >> 5, u = threadIdx.x & 31;
    const float bv = __ldg(p.b[layer] + (tile >> 1) * 256 + g * 64 + (tile & 1) * 32 + u);
    s_bias[threadIdx.x] = bv * (g == 2 ? -2.f * kLog2e : -kLog2e);
    s_bias_h[threadIdx.x] = threadIdx.x < p.head_rows ? __ldg(p.bh + threadIdx.x) * (-2.f * kLog2e) : 0.f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();                 // every CTA's barriers exist before any remote arrive / remote store
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  ptx::griddep_launch_dependents();
  if (threadIdx.x == 0) STRACE(0);
  ptx::griddep_wait();                     // the previous step (same stream) wrote our state
  if (threadIdx.x == 0) STRACE(1);

  const bool is_head = c == 0;
  const int n_kb_lstm = layer == 0 ? SM_HK + p.kbx : 2 * SM_HK;
  if (p.debug_stop == 1) {                 // developer timing switch (results invalid): launch + set-up + teardown only
    __syncthreads();
    ptx::cluster_sync_all();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, 512);
    return;
  }

  if (warp == 0) {
    // ===================== producer: own activation images, then this CTA's weight k-blocks =====================
    if (lane == 0) {
      const uint64_t pol_keep = ptx::l2_policy_evict_last();
      // recurrent operand h_l of the previous step: first N rows of the 4 packed k-block images (hi, lo)
      ptx::mbar_expect_tx(bar_act, SM_HK * nparts * img);
      for (int kb = 0; kb < SM_HK; ++kb)
        for (uint32_t part = 0; part < nparts; ++part)
          ptx::bulk_g2s(act(kb, part), p.hp_in + ((size_t)(layer * SM_HK + kb) * 2 + part) * TC_A_IMG, img, bar_act);
      int s = 0;
      uint32_t phs = 0;
      const uint8_t* wl = p.w[layer];
      const int KB = n_kb_lstm;
      const int kb_in = layer == 0 ? p.kbx : SM_HK;
      const int total = KB + (is_head ? SM_HK : 0);
      for (int i = 0; i < total; ++i) {
        ptx::mbar_wait(empty_bar(s), phs ^ 1u);
        const uint32_t dst = base + (uint32_t)s * SM_STAGE;
        if (i < KB) {
          // weight K order [input | recurrent]; recurrent k-blocks are consumed first
          const int wk = i < SM_HK ? kb_in + i : i - SM_HK;
          // (the tile's 128 gate rows are stored contiguously per (tile, k-block): [hi image | lo image], one bulk copy --
          //  gathering them as eight 4 KB chunks of the large-batch weight images cost ~0.6 us of issue time per k-block)
          ptx::mbar_expect_tx(full_bar(s), nparts * (uint32_t)TC_A_IMG);
          ptx::bulk_g2s_hint(dst, wl + (size_t)(tile * KB + wk) * SM_STAGE, nparts * (uint32_t)TC_A_IMG, full_bar(s), pol_keep);
        } else {
          const int kb = i - KB;
          const uint32_t hb = (uint32_t)p.head_rows * 128u;
          ptx::mbar_expect_tx(full_bar(s), nparts * hb);
          for (uint32_t part = 0; part < nparts; ++part)
            ptx::bulk_g2s_hint(dst + part * TC_A_IMG, p.wh + ((size_t)kb * 2 + part) * hb, hb, full_bar(s), pol_keep);
        }
        if (++s == p.stages) { s = 0; phs ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // bf16x3 as TWO MMAs per k-step: the hi and lo images of a B operand are adjacent in shared memory, so
      //   A_hi x [X_hi ; X_lo]  (one MMA, N' = 2N: columns [0,N) get hi*hi, [N,2N) hi*lo)  +  A_lo x X_hi  (columns [2N,3N))
      // give all three split products; the epilogue adds the three column groups.  At these sizes an MMA costs ~80 cycles
      // whatever its N, so dropping one of three shortens every GEMM phase of the step's dependency chain by a third.
      const uint32_t idesc = ptx::make_idesc_bf16(128, N), idesc2 = ptx::make_idesc_bf16(128, 2 * N);
      int s = 0;
      uint32_t phs = 0;
      auto kblock = [&](uint32_t d_tmem, int slot, int ks, uint32_t& accum) {
        ptx::mbar_wait(full_bar(s), phs);
        ptx::tc_fence_after();
        const uint32_t sa = base + (uint32_t)s * SM_STAGE;
        const uint64_t a_hi = ptx::make_sw128_desc(sa), a_lo = ptx::make_sw128_desc(sa + TC_A_IMG);
        const uint64_t b_hi = ptx::make_sw128_desc(act(slot, 0));
        for (int kk = 0; kk < ks; ++kk) {
          const uint64_t adv = (uint64_t)(kk * 2);
          if (nparts == 2) {
            // (separate accumulators: interleaving two instruction shapes on ONE accumulator serialised them at ~140
            //  cycles each; two independent accumulate chains pipeline)
            ptx::umma_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc2, accum);              // B = [hi ; lo], 2N rows -> cols [0, 2N)
            ptx::umma_bf16(d_tmem + 2 * N, a_lo + adv, b_hi + adv, idesc, accum);       // lo * hi -> cols [2N, 3N)
          } else {
            ptx::umma_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc, accum);
          }
          accum = 1u;
        }
        ptx::umma_commit(empty_bar(s));
        if (++s == p.stages) { s = 0; phs ^= 1u; }
      };
      uint32_t accum = 0;
      ptx::mbar_wait(bar_act, 0);
      STRACE(2);
      for (int kb = 0; kb < SM_HK; ++kb) { kblock(tmem_base, kb, 4, accum); if (kb == 0) STRACE(3); }   // recurrent half (previous step's h_l)
      STRACE(4);
      if (layer == 0) {
        ptx::mbar_wait(bar_x, 0);
        STRACE(5);
        ptx::fence_proxy_async();
        for (int kb = 0; kb < p.kbx; ++kb) {
          const int left = p.x_ksteps - 4 * kb;
          kblock(tmem_base, SM_HK + kb, left < 4 ? left : 4, accum);
        }
      } else {
        mbar_wait_cluster(bar_in, 0);                                             // h'_0 from the eight layer-0 CTAs
        STRACE(5);
        ptx::fence_proxy_async();
        for (int kb = 0; kb < SM_HK; ++kb) kblock(tmem_base, SM_HK + kb, 4, accum);
      }
      ptx::umma_commit(bar_acc);
      STRACE(6);
      if (is_head) {
        mbar_wait_cluster(bar_in, 0);                                             // h'_1 from the eight layer-1 CTAs
        STRACE(7);
        ptx::fence_proxy_async();
        accum = 0;
        for (int kb = 0; kb < SM_HK; ++kb) kblock(tmem_base + 256, kb, 4, accum);
        ptx::umma_commit(bar_acc2);
        STRACE(8);
      }
    }
  } else if (warp == 6) {
    // ===================== trigger finaliser (last CTA): window / threshold / decision, then the mask to every CTA ======
    if (p.trig.enabled && c == SM_CL - 1) {
      mbar_wait_cluster(bar_trig, 0);
      const int S = p.trig.S;
      if (lane == 0) *p.trig.trig_count = 0;
      __syncwarp();
      const int cnt = *reinterpret_cast<volatile int32_t*>(p.trig.count);
      int fired = 0;
      if (lane < S) {
        if (p.trig.W <= 16)
          gp_trig_finalize_rollout16(lane, S, p.trig.D, p.trig.var_rows, p.trig.window, p.trig.W, cnt, p.trig.warmup, p.trig.factor,
                                     p.trig.value, p.trig.thr, p.trig.mask, p.trig.trig_list, p.trig.trig_count);
        else
          gp_trig_finalize_rollout(lane, S, p.trig.D, p.trig.var_rows, p.trig.window, p.trig.W, cnt, p.trig.warmup, p.trig.factor,
                                   p.trig.value, p.trig.thr, p.trig.mask, p.trig.trig_list, p.trig.trig_count);
        fired = p.trig.mask != nullptr ? (int)p.trig.mask[lane] : 0;
      }
      __syncwarp();
      if (lane == 0 && p.trig.warmup && cnt < p.trig.W) p.trig.count[0] = cnt + 1;
      if (!p.trig.warmup) {
        const uint32_t mask_addr = ptx::smem_u32(const_cast<int*>(s_mask));
        if (lane < S)
          for (uint32_t q2 = 0; q2 < SM_CL; ++q2)
            asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(mapa(mask_addr + 4u * lane, q2)), "r"(fired) : "memory");
        __syncwarp();
        if (lane == 0)
          for (uint32_t q2 = 0; q2 < SM_CL; ++q2) mbar_arrive_remote_addr(mapa(bar_mask, q2));
      }
    }
  } else {
    // ===================== epilogue / SIMT warps (2..5) =====================
    const int et = threadIdx.x - 64;                   // 0..127
    const int q = warp & 3;                            // TMEM lane quarter of this warp == gate index it reads
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    if (layer == 0) {
      // x-pack: fp32 [rows, G] -> bf16 hi/lo K-major images (slots 4..4+kbx), zero padded to N rows / 64-column k-blocks
      const int chunks = p.kbx * 8;
      for (int u = et; u < N * chunks; u += 128) {
        const int r = u / chunks, ch = u - r * chunks;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int col = ch * 8 + e;
          v[e] = (r < p.rows && col < p.G) ? __ldg(p.x + (size_t)r * p.ldx + col) : 0.f;
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split2_bf16(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
        const uint32_t off = (uint32_t)r * 128u + ((uint32_t)((ch & 7) ^ (r & 7)) << 4);
        const int slot = SM_HK + (ch >> 3);
        *reinterpret_cast<uint4*>(smem_raw + (act(slot, 0) - base) + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (nparts == 2) *reinterpret_cast<uint4*>(smem_raw + (act(slot, 1) - base) + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_x);
      if (et == 0) STRACE(10);
    }
    // c of the previous step for this thread's cells: unit = et & 31, rows (et >> 5) + 4 j
    const int u = et & 31;
    const int unit = tile * 32 + u;
    const float* c_in = p.c_in + (size_t)layer * p.rows * SM_H;
    const float* h_in = p.h_in + (size_t)layer * p.rows * SM_H;
    float* c_out = p.c_out + (size_t)layer * p.rows * SM_H;
    float* h_out = p.h_out + (size_t)layer * p.rows * SM_H;
    float cp[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int r = (et >> 5) + 4 * j;
      cp[j] = (r < p.rows && r < N) ? __ldg(c_in + (size_t)r * SM_H + unit) : 0.f;
    }
    if (p.trig.enabled) {
      // ---- GP variance at the statistic row of every rollout, dims c, c+16, ... of this CTA, in the idle window before
      //      the accumulator is ready: 128 threads = (dim slot, factor row) pairs; |Linv k|^2 and |L_q^T k|^2 are summed in a
      //      fixed order by one thread per (dim, rollout), so the result does not depend on scheduling.
      constexpr int MP = 40;
      const int S = p.trig.S, D = p.trig.D;
      float* s_k = s_gate;                                   // [dims][S][MP]
      float* s_part = s_gate + SM_TRIG_DIMS * SM_TRIG_MAX_S * MP;      // [dims][S][2 MP]
      int nd = 0;
      for (int d = (int)c; d < D; d += SM_CL) ++nd;
      for (int i = et; i < nd * S * MP; i += 128) {
        const int ds = i / (S * MP), rem = i - ds * (S * MP), sidx = rem / MP, m = rem - sidx * MP;
        const int d = (int)c + ds * SM_CL;
        const float xv = __ldg(p.x + (size_t)p.trig.stat_rows[sidx] * p.ldx + d);
        const float t = (xv - __ldg(p.trig.z + (size_t)d * MP + m)) * (1.0f / __ldg(p.trig.hyp + d * 4 + 0));
        s_k[i] = __ldg(p.trig.hyp + d * 4 + 1) * ex2_ftz(t * t * (-0.5f * kLog2e));
      }
      ptx::named_bar_sync(1, 128);
      for (int i = et; i < nd * 2 * MP; i += 128) {
        const int ds = i / (2 * MP), row = i - ds * (2 * MP);
        const int d = (int)c + ds * SM_CL;
        const float4* mrow = reinterpret_cast<const float4*>((row < MP ? p.trig.linv : p.trig.lqt) + ((size_t)d * MP + (row % MP)) * MP);
        float4 mv[MP / 4];
#pragma unroll
        for (int m4 = 0; m4 < MP / 4; ++m4) mv[m4] = __ldg(mrow + m4);
        for (int sidx = 0; sidx < S; ++sidx) {
          const float* kk = s_k + (ds * S + sidx) * MP;
          float a0 = 0.f, a1 = 0.f;
#pragma unroll
          for (int m4 = 0; m4 < MP / 4; ++m4) {
            a0 = fmaf(mv[m4].x, kk[4 * m4], a0); a1 = fmaf(mv[m4].y, kk[4 * m4 + 1], a1);
            a0 = fmaf(mv[m4].z, kk[4 * m4 + 2], a0); a1 = fmaf(mv[m4].w, kk[4 * m4 + 3], a1);
          }
          const float a = a0 + a1;
          s_part[(ds * S + sidx) * 2 * MP + row] = a * a;
        }
      }
      ptx::named_bar_sync(1, 128);
      if (et < nd * S) {
        const int ds = et / S, sidx = et - ds * S;
        const int d = (int)c + ds * SM_CL;
        const float* pp = s_part + (ds * S + sidx) * 2 * MP;
        float sv = 0.f, sw = 0.f;
        for (int m = 0; m < MP; ++m) { sv += pp[m]; sw += pp[MP + m]; }
        p.trig.var_rows[(size_t)d * S + sidx] = (__ldg(p.trig.hyp + d * 4 + 1) - sv) + sw + __ldg(p.trig.hyp + d * 4 + 3);
      }
      ptx::named_bar_sync(1, 128);               // the stores above precede thread 0's release (cumulative)
      if (et == 0) mbar_arrive_remote_addr(mapa(bar_trig, SM_CL - 1));
      // (nobody waits for the decision here: the LSTM advances every rollout and the few that fired are restored from
      //  the input state at the end of the launch -- waiting cost the decision steps ~9 us on the critical path)
    }
    if (et == 0) STRACE(11);
    ptx::mbar_wait(bar_acc, 0);
    if (et == 0) STRACE(12);
    ptx::tc_fence_after();
#pragma unroll
    for (int rc = 0; rc < 4; ++rc) {
      const int r0 = rc * 16;
      if (r0 >= N) break;
      float v[16];
      ptx::tmem_ld16_wait(tmem_base + tlane + (uint32_t)r0, v);       // gate q, unit = lane, batch rows r0 .. r0+15
      if (nparts == 2) {                                              // + the hi*lo and lo*hi column groups
        float v2[16], v3[16];
        ptx::tmem_ld16_wait(tmem_base + tlane + (uint32_t)(N + r0), v2);
        ptx::tmem_ld16_wait(tmem_base + tlane + (uint32_t)(2 * N + r0), v3);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += v2[j] + v3[j];
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) s_gate[(q * 16 + j) * 32 + lane] = v[j];
      ptx::named_bar_sync(1, 128);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = (et >> 5) + 4 * jj;
        const int r = r0 + j;
        const float cprev = cp[rc * 4 + jj];
        float hn, cn;
        lstm_cell_fast<0>(s_gate[(0 * 16 + j) * 32 + u], s_gate[(1 * 16 + j) * 32 + u], s_gate[(2 * 16 + j) * 32 + u],
                          s_gate[(3 * 16 + j) * 32 + u], s_bias[u], s_bias[32 + u], s_bias[64 + u], s_bias[96 + u], cprev, hn, cn);
        if (r < p.rows) {
          if (p.hold != nullptr && p.hold[r / p.rows_per_flag] != 0) {      // generate_frames.py:289-295: state not advanced
            hn = __ldg(h_in + (size_t)r * SM_H + unit);
            cn = cprev;
          }
          c_out[(size_t)r * SM_H + unit] = cn;
          h_out[(size_t)r * SM_H + unit] = hn;
        } else {
          hn = 0.f;
        }
        __nv_bfloat16 bh, bl;
        split_bf16(hn, bh, bl);
        s_h[r * 32 + u] = __bfloat16_as_ushort(bh);
        s_h[(N + r) * 32 + u] = __bfloat16_as_ushort(bl);
      }
      ptx::named_bar_sync(1, 128);
    }
    if (et == 0) STRACE(13);
    ptx::tc_fence_before();
    // distribute this tile's h' (N rows x 32 units, bf16 hi / lo): 16-byte chunks into the B-operand images of the
    // consumers (layer 0 -> the eight layer-1 CTAs, slots 4..7; layer 1 -> the head CTA, slots 0..3) through
    // distributed shared memory, and into the packed state block in global memory for the next time step
    {
      const int kb = tile >> 1, ch0 = (tile & 1) * 4;
      const int n_cons = layer == 0 ? 8 : 1;
      const int slot = layer == 0 ? SM_HK + kb : kb;
      for (int it = et; it < N * 4 * (int)nparts; it += 128) {
        const int part = it / (N * 4), rem = it - part * (N * 4);
        const int r = rem >> 2, j = rem & 3;
        const uint4 val = *reinterpret_cast<const uint4*>(s_h + ((size_t)(part * N + r) * 32 + j * 8));
        const uint32_t off = (uint32_t)r * 128u + ((uint32_t)((ch0 + j) ^ (r & 7)) << 4);
        const uint32_t local = act(slot, part) + off;
        for (int q2 = 0; q2 < n_cons; ++q2) st_cluster16(mapa(local, layer == 0 ? 8u + q2 : 0u), val);
        // (padding rows up to N are written as zeros: the next step loads the first N rows as its B operand)
        *reinterpret_cast<uint4*>(p.hp_out + ((size_t)(layer * SM_HK + kb) * 2 + part) * TC_A_IMG + off) = val;
      }
      ptx::fence_proxy_async_all();
      __syncwarp();
      if (lane == 0) {
        // one release fence for the warp's remote stores, then relaxed arrives (a release-arrive per consumer waits for
        // the stores to be performed each time: the eight arrives took ~2 us and reached the consumers 0.3 us apart)
        asm volatile("fence.acq_rel.cluster;" ::: "memory");
        for (int q2 = 0; q2 < n_cons; ++q2)
          asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa(bar_in, layer == 0 ? 8u + q2 : 0u)) : "memory");
      }
      if (et == 0) STRACE(14);
    }
    const bool rs_fused = p.trig.enabled && !p.trig.warmup && p.trig.rs_eps != nullptr;
    const bool decide = p.trig.enabled && !p.trig.warmup;
    if (decide) {
      // this step's decision (one flag per rollout, stored into our shared memory by the finaliser).  A rollout that fired
      // keeps its LSTM state (generate_frames.py:289-295): this CTA's slice of its rows -- fp32 h, c and the packed h
      // image chunks of our 32 units -- is copied back from the input block.
      mbar_wait_cluster(bar_mask, 0);
      const int kb = tile >> 1, ch0 = (tile & 1) * 4;
      for (int sidx = 0; sidx < p.trig.S; ++sidx) {
        if (s_mask[sidx] == 0) continue;
        const int rb = sidx * p.rows_per_flag, re = min(rb + p.rows_per_flag, p.rows);
        for (int i = et; i < (re - rb) * 32; i += 128) {
          const int r = rb + (i >> 5);
          const size_t idx = (size_t)r * SM_H + tile * 32 + (i & 31);
          h_out[idx] = __ldg(h_in + idx);
          c_out[idx] = __ldg(c_in + idx);
        }
        for (int i = et; i < (re - rb) * 4 * (int)nparts; i += 128) {
          const int part = i / ((re - rb) * 4), rem = i - part * ((re - rb) * 4);
          const int r = rb + (rem >> 2), j = rem & 3;
          const size_t off = ((size_t)(layer * SM_HK + kb) * 2 + part) * TC_A_IMG + (size_t)r * 128u + ((uint32_t)((ch0 + j) ^ (r & 7)) << 4);
          *reinterpret_cast<uint4*>(p.hp_out + off) = __ldg(reinterpret_cast<const uint4*>(p.hp_in + off));
        }
      }
    }
    if (is_head) {
      // y^T [head rows x N] = tanh(W_o h'_1 + b_o): lane = output column, TMEM column = batch row
      ptx::mbar_wait(bar_acc2, 0);
      if (et == 0) STRACE(15);
      ptx::tc_fence_after();
      const int jcol = q * 32 + lane;
      for (int r0 = 0; r0 < N; r0 += 16) {
        float v[16];
        ptx::tmem_ld16_wait(tmem_base + 256 + tlane + (uint32_t)r0, v);
        if (nparts == 2) {
          float v2[16], v3[16];
          ptx::tmem_ld16_wait(tmem_base + 256 + tlane + (uint32_t)(N + r0), v2);
          ptx::tmem_ld16_wait(tmem_base + 256 + tlane + (uint32_t)(2 * N + r0), v3);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += v2[j] + v3[j];
        }
        if (jcol < p.n_valid) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int r = r0 + j;
            // (rows of a fired rollout receive the GP sample below instead of the LSTM prediction, generate_frames.py:291-292)
            if (r < p.rows && !(rs_fused && s_mask[r / p.rows_per_flag] != 0))
              p.y[(size_t)r * p.ldy + jcol] = tanh_fast_prescaled(v[j], s_bias_h[jcol]);
          }
        }
      }
      ptx::tc_fence_before();
    }
    if (rs_fused) {
      // Fired rollouts (rare): their decoder input is a GP posterior sample of the step's INPUT latent
      // (generate_frames.py:290-292).  (rollout, dim) problems are dealt to the CTAs by dim; the 128 epilogue threads
      // solve one at a time in the weight ring, which is idle by now (every MMA that read it has completed).
      float* smf = reinterpret_cast<float*>(smem_raw);
      for (int sidx = 0; sidx < p.trig.S; ++sidx) {
        if (s_mask[sidx] == 0) continue;
        for (int d = (int)c; d < p.trig.D; d += SM_CL) {
          gp_rsample_body<128>(smf, et, [] { ptx::named_bar_sync(1, 128); }, sidx, d, p.trig.n_points, p.trig.D, 40, p.x, p.ldx,
                               p.trig.rs_eps, p.trig.z, p.trig.linv, p.trig.lqt, p.trig.alpha, p.trig.hyp, p.y, p.ldy);
          ptx::named_bar_sync(1, 128);
        }
      }
    }
  }
  // ---- teardown: nobody leaves while a peer may still store into its shared memory or arrive on its barriers ----
  if (threadIdx.x == 64) STRACE(16);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (threadIdx.x == 64) STRACE(17);
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// Weight images of the small-batch kernel: [tile 0..7][k-block][hi | lo][128 rows x 128 B], the 128 rows being the four
// 32-row chunks (one per gate) of units 32 tile .. 32 tile + 31 taken from the 256-row image of N tile (tile >> 1) of the
// large-batch packing (rows g*64 + 32 (tile & 1) ..; multiples of 8 rows, so the 128-byte swizzle pattern is unchanged).
__global__ void small_repack_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int KB) {
  const int tile = blockIdx.x, wk = blockIdx.y;
  const uint4* s4 = reinterpret_cast<const uint4*>(src + (size_t)((tile >> 1) * KB + wk) * (2u * 256u * 128u));
  uint4* d4 = reinterpret_cast<uint4*>(dst + (size_t)(tile * KB + wk) * SM_STAGE);
  for (int i = threadIdx.x; i < SM_STAGE / 16; i += blockDim.x) {
    const int part = i / (TC_A_IMG / 16), rem = i % (TC_A_IMG / 16);
    const int row = rem >> 3, ch = rem & 7;                 // 128-byte rows of 8 chunks
    const int g = row >> 5, r32 = row & 31;
    d4[i] = s4[(size_t)part * (256 * 8) + (size_t)(g * 64 + (tile & 1) * 32 + r32) * 8 + ch];
  }
}

// ---------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------
int lstm_small_pack(dvg_lstm_s* h, cudaStream_t stream) {
  if (!h->tc_ok || h->dims.hidden_size != SM_H || h->dims.n_layers != 2 || h->dims.kind != DVG_LSTM) return DVG_OK;
  const TcGemmPlan* pl[2] = {&h->tc_layer0f, &h->tc_layer[1]};
  for (int l = 0; l < 2; ++l) {
    const int KB = pl[l]->kb0 + pl[l]->kb1;
    if (!h->small_w[l]) DVG_CUDA(cudaMalloc(&h->small_w[l], (size_t)8 * KB * SM_STAGE));
    small_repack_kernel<<<dim3(8, KB), 256, 0, stream>>>(pl[l]->w, h->small_w[l], KB);
    DVG_LAUNCH_CHECK();
  }
  return DVG_OK;
}
void lstm_small_free(dvg_lstm_s* h) {
  for (int l = 0; l < 2; ++l) {
    if (h->small_w[l]) cudaFree(h->small_w[l]);
    h->small_w[l] = nullptr;
  }
}
static size_t small_smem_bytes(int N, int stages) {
  return (size_t)stages * SM_STAGE + (size_t)16 * N * 128 + 8192 + (size_t)128 * N + 2 * 128 * sizeof(float) + 256 + 64;
}

bool lstm_small_usable(const dvg_lstm_s* h, int rows) {
  static int off = -1;
  if (off < 0) {
    const char* e = getenv("DVG_TC_SMALL");          // developer switch: 0 = one launch per GEMM for small batches
    off = (e && e[0] == '0') ? 1 : 0;
  }
  if (off || !h->tc_ok || h->dims.kind != DVG_LSTM || h->small_w[0] == nullptr) return false;
  if (h->dims.hidden_size != SM_H || h->dims.n_layers != 2) return false;
  if (h->dims.input_size > 128 || h->tc_head.n_tile > 128 || h->sm_count < SM_CL) return false;
  return rows >= 1 && rows <= 64;
}

bool lstm_small_can_fuse_trigger(const dvg_lstm_s* h, const dvg_gp_s* g, int rows, int S) {
  return lstm_small_usable(h, rows) && !g->big && g->mp == 40 && g->dims.num_dims <= SM_CL * SM_TRIG_DIMS && S >= 1 &&
         S <= SM_TRIG_MAX_S && S <= g->var_rows_cap;
}

int lstm_small_launch(dvg_lstm_s* h, int nsplit, int rows, const float* x, int ldx, const float* h_in, const float* c_in,
                      const uint8_t* hp_in, float* h_out, float* c_out, uint8_t* hp_out, float* y, int ldy,
                      const uint8_t* hold, int rows_per_flag, cudaStream_t stream, dvg_gp_s* g, const StepTrigHost* trig) {
  SmallArgs a{};
  if (trig != nullptr) {
    a.trig.enabled = 1; a.trig.S = trig->S; a.trig.D = g->dims.num_dims; a.trig.W = trig->W; a.trig.warmup = trig->warmup;
    a.trig.factor = trig->factor; a.trig.stat_rows = trig->stat_rows;
    a.trig.z = g->z; a.trig.linv = g->linv; a.trig.lqt = g->lqt; a.trig.hyp = g->hyp;
    a.trig.var_rows = g->var_rows; a.trig.window = trig->window; a.trig.count = trig->count; a.trig.value = trig->value;
    a.trig.thr = trig->thr; a.trig.mask = trig->mask; a.trig.trig_list = g->trig_list; a.trig.trig_count = g->trig_count;
    rows_per_flag = rows / trig->S;
    hold = nullptr;
    a.trig.n_points = rows / trig->S;
    a.trig.alpha = g->alpha;
    a.trig.rs_eps = trig->warmup ? nullptr : trig->rs_eps;
  }
  a.rows = rows; a.N = (rows + 15) / 16 * 16; a.G = h->dims.input_size; a.ldx = ldx; a.ldy = ldy;
  a.n_valid = h->dims.output_size; a.nparts = nsplit == 1 ? 1 : 2; a.rows_per_flag = rows_per_flag > 0 ? rows_per_flag : 1;
  a.kbx = ceil_div(a.G, 64); a.x_ksteps = ceil_div(a.G, 16); a.head_rows = h->tc_head.n_tile;
  a.x = x; a.h_in = h_in; a.c_in = c_in; a.h_out = h_out; a.c_out = c_out; a.hp_in = hp_in; a.hp_out = hp_out;
  a.w[0] = h->small_w[0]; a.w[1] = h->small_w[1]; a.wh = h->tc_head.w;
  a.b[0] = h->tc_layer0f.bias; a.b[1] = h->tc_layer[1].bias; a.bh = h->tc_head.bias;
  a.y = y; a.hold = hold;
  { const char* e = getenv("DVG_SMALL_STOP"); a.debug_stop = e ? atoi(e) : 0; }
  int stages = SM_MAX_STAGES;
  while (stages > 2 && small_smem_bytes(a.N, stages) > 227 * 1024) --stages;
  a.stages = stages;
  if (a.trig.rs_eps != nullptr)
    DVG_REQUIRE(sizeof(float) * gp_rsample_smem_floats(a.trig.n_points, 40) <= (size_t)stages * SM_STAGE,
                "small-batch step: the in-launch resample of %d points does not fit the weight ring", a.trig.n_points);
  const size_t smem = small_smem_bytes(a.N, stages);
  DVG_REQUIRE(smem <= 227 * 1024, "small-batch LSTM step: %d rows need %zu B of shared memory", rows, smem);
  static bool configured = false;
  if (!configured) {
    DVG_CUDA(cudaFuncSetAttribute(lstm_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DVG_CUDA(cudaFuncSetAttribute(lstm_small_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(SM_CL);
  cfg.blockDim = dim3(SM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = SM_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 2;
  static unsigned long long* tbuf = nullptr;
  static int n_launch = 0;
  const bool tr = getenv("DVG_SMALL_TRACE") != nullptr;
  if (tr) {
    if (!tbuf) cudaMalloc(&tbuf, 16 * 32 * 8);
    cudaMemsetAsync(tbuf, 0, 16 * 32 * 8, stream);
    a.trace = tbuf;
  }
  h->prof_mark(stream);
  DVG_CUDA(cudaLaunchKernelEx(&cfg, lstm_small_kernel, (const SmallArgs)a));
  h->prof_mark(stream);
  if (tr) {
    cudaStreamSynchronize(stream);
    if (n_launch++ == 6) {
      unsigned long long hb[16 * 32];
      cudaMemcpy(hb, tbuf, sizeof(hb), cudaMemcpyDeviceToHost);
      unsigned long long t0 = ~0ull;
      for (int b = 0; b < 16; ++b) if (hb[b * 32] && hb[b * 32] < t0) t0 = hb[b * 32];
      fprintf(stderr, "SMALL TRACE (ns): cta: 0 start 1 depwait | mma: 2 act 3 kb0 4 rec-done 5 in-ready 6 issued 7 head-in 8 head-issued | epi: 10 xpack 11 pre-acc 12 acc 13 cells 14 sent 15 head-acc 16 done 17 exit\n");
      for (int b = 0; b < 16; ++b) {
        fprintf(stderr, "cta %2d:", b);
        for (int i = 0; i < 18; ++i) fprintf(stderr, " %lld", hb[b * 32 + i] ? (long long)(hb[b * 32 + i] - t0) : -1ll);
        fprintf(stderr, "\n");
      }
    }
  }
  return DVG_OK;
}

}  // namespace dvg
