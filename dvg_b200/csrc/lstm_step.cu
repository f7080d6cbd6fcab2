// One LSTM time step (models/lstm.py:65-72, gaussian_lstm :166-175) in ONE persistent launch, tensor-core variants.
//
// Grid: one CTA per SM, clusters of 2 (cta_group::2 pairs).  A pair owns work items of 256 rows x n_tile columns:
//   LSTM_l(rg, nt)  layer l, row group rg (256 rows), N tile nt = [i|f|g|o] x 64 hidden units.
//                   K order: recurrent k-blocks (packed h_l of the previous step) first -- they depend on nothing
//                   produced in this launch -- then the input k-blocks (layer 0: x with the embed Linear folded in;
//                   layer l > 0: the packed h'_{l-1} written by this launch).
//   HEAD(rg)        tanh(h'_{L-1} W_o^T + b_o)  or  mu / logvar / z.
// Dependencies are per (layer, rg, nt) counters in global memory: the epilogue of LSTM_l(rg, nt) publishes one packed
// k-block image of h'_l, and a consumer waits for exactly the k-block it is about to load (target 2 = both CTAs of the
// producing pair).  Items are processed in a global order in which every dependency precedes its consumer and all pairs
// are co-resident, so the schedule cannot deadlock.
//
// Warp roles (608 threads with the default 16 epilogue warps):
//   warp 0       TMA producer (one lane): dependency polls, cp.async.bulk of the k-block stages (the weight half of a
//                stage is requested before the poll, L2 eviction hints: weights evict_last, activations evict_first)
//   warp 1       leader CTA: tcgen05.mma issuer;  peer CTA: relays "my stage landed" to the leader
//   warps 2-17   x-pack prologue of the pair's own layer-0 items, GP trigger partial sums (4 threads per
//                (rollout, dim) task, in the idle window before the first accumulator is ready), tile epilogues;
//                after the CTA's last tile: GP resample problems of the fired rollouts from a dynamic queue
//   warp 19      weight-stream producer (one lane): cp.async.bulk of every stage's weight half, L2 evict_last
//   warp 18      auxiliary: in the last ceil(S/32) CTAs it finalises the GP trigger (window / threshold / decision)
//                once the partial variances have been delivered, and publishes the mask
//
// Design decisions that came out of the ncu / %globaltimer traces (profiles/r01_lstm_step.md):
//   * the x operand is packed by the consuming pair itself into a private scratch slab (no all-CTA pre-pass, no
//     cross-CTA flag on the way to the first MMA);
//   * k-block granular dependencies instead of "all N tiles of the row group";
//   * the LSTM epilogue is a compact loop over 4 hidden units (a fully unrolled version was 36 KB of straight-line
//     code and ~70 % of its issue slots were instruction-fetch stalls); results are parked in already consumed TMEM
//     columns (tcgen05.st) so no register array needs dynamic indexing;
//   * nobody waits for the trigger mask: the LSTM always advances and, in the rare step where rollouts fired, their
//     state rows are restored from the input block at the end of the launch (generate_frames.py:289-295: a triggered
//     rollout does not advance its LSTM) and their GP samples (generate_frames.py:291-292) are computed by the CTAs
//     that run out of tiles first;
//   * the trigger is finalised by a dedicated warp instead of stalling one CTA's epilogue warps;
//   * consecutive launches overlap their prologue / teardown through programmatic dependent launch.
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "gp_rsample.cuh"
#include "gp_trigger.cuh"
#include "tc_common.cuh"

namespace dvg {

enum { PH_LSTM = 1, PH_TANH = 2, PH_GAUSS = 3 };
constexpr int STEP_MAX_PHASES = MAX_LAYERS + 1;
#ifndef DVG_STEP_EW
#define DVG_STEP_EW 16
#endif
constexpr int STEP_EW = DVG_STEP_EW;                 // epilogue warps: 8 or 16 (4 per SM partition hide the MUFU latency)
static_assert(STEP_EW == 8 || STEP_EW == 16, "epilogue warps: 8 or 16");
constexpr int STEP_NSUB = STEP_EW / 4;               // warps sharing a TMEM lane quarter split the 64 units of a tile
constexpr int STEP_UPW = 64 / STEP_NSUB;             // hidden units per warp and tile: 32 or 16
constexpr int STEP_RB = STEP_UPW * 4;                // bytes per row of a warp's transpose buffer: 128 or 64
constexpr int STEP_CPR = STEP_UPW / 4;               // 16-byte chunks per row: 8 or 4
constexpr int STEP_EBUF_BYTES = STEP_EW * 32 * STEP_RB;   // 32 KB either way
constexpr int STEP_THREADS = 64 + STEP_EW * 32 + 64;
constexpr int AUX_WARP = 2 + STEP_EW;
constexpr int WPROD_WARP = 3 + STEP_EW;     // weight-stream producer
constexpr int STEP_MAX_STAGES = 6;
constexpr int STEP_XMAX = 8;            // layer-0 items per pair (one x-ready mbarrier each)
#ifndef DVG_STEP_NPOLY
#define DVG_STEP_NPOLY 0      // measured on kth_s100: 3 -> 52.1 us, 2 -> 50.9, 1 -> 50.7, 0 -> 49.7 us per step
#endif
constexpr int STEP_NPOLY = DVG_STEP_NPOLY;   // sigmoid exponentials on the FMA pipe (rest: MUFU), see lstm_cell_fast
#ifndef DVG_STEP_TRIG_EARLY
#define DVG_STEP_TRIG_EARLY 1
#endif
constexpr bool STEP_TRIG_EARLY = DVG_STEP_TRIG_EARLY != 0;   // trigger partial sums before the first tile epilogue
#ifndef DVG_STEP_POLL_BATCH
#define DVG_STEP_POLL_BATCH 0     // measured: sampling 4 flags per poll + grouped acquires is SLOWER (kth_s100 45.1 vs 43.9 us per step)
#endif
#ifndef DVG_STEP_WFENCE
#define DVG_STEP_WFENCE 1
#endif
#ifndef DVG_STEP_WARM
#define DVG_STEP_WARM 0      // measured: no gain (kth_s100 +0.3 us, trigger steps of bair_s32 +1.2 us)
#endif
constexpr int STEP_BAR_BYTES = 384;     // mbarriers + tmem slot + misc words

struct StepPhase {
  int type, n_tile, n_tiles, kb_in, kb_rec, item_begin, in_ksteps;
  const uint8_t* a_in;    // layer 0: packed-x scratch [n_tiles][RT][kb_in]; else h' images of the layer below [RT][kb_in]
  const uint8_t* a_rec;   // packed h of the previous step [RT][kb_rec]
  const uint8_t* w; const float* bias;
  const int* wait_flags;  // [groups][kb_in] counters of the producing phase (nullptr: layer 0, local x-ready barrier)
  int* done_flags;        // [groups][n_tiles]
  const int* prev_done;   // chained launches: done_flags of the same layer in the PREVIOUS launch (its packed h' is our h)
  const float* c_in; const float* h_in; float* h_out; float* c_out; uint8_t* hp_out;   // LSTM
  float* y; int ldy; int n_valid;                                                       // TANH
  const float* eps; float* z; float* mu; float* logvar; int Z;                          // GAUSS
};
struct StepTrig {
  int enabled, S, D, mp, W, warmup;
  float factor;
  const int32_t* stat_rows;
  const float* z; const float* linv; const float* lqt; const float* hyp;
  float* var_rows; unsigned int* ticket; float* window; int32_t* count;
  float* value; float* thr; uint8_t* mask; int* trig_list; int* trig_count;
  const float* rs_eps; const float* alpha; float* rs_out; int rs_ldo; int n_points;   // in-kernel rsample of fired rollouts
};
struct StepArgs {
  int rows, row_tiles, groups, nsplit, stages, n_phases, total_items, H, L, G, ldx, kbx, rows_per_flag, restore;
  uint32_t stage_bytes;
  const float* x; uint8_t* xp;
  const uint8_t* hold;            // mask known BEFORE the launch (plain dvg_lstm_step); nullptr in trigger-fused steps
  const int* sched; int sched_len;   // optional host-built item order: pair p runs sched[p], sched[p + pairs], ...
  int* flag_words; int n_flag_words;     // dependency counters, then [mask_ready][done_ctr]; exit counter follows
  int* mask_ready; int* done_ctr; int* rs_next; int* rs_done;
  int chained, in_chain, chain_idx, rot, self_reset, prev_trig, prev_restore;
  int* retired;                   // number of launches of the chain that have completely exited
  const int* prev_mask_ready; const int* prev_fin; const int* prev_trig_count;
  int* fin_ctr;
  int* dim_ctr; int* fin_claim;   // GP trigger work is claimed in start order (see trigger_partials)
  int* reset_set;
  float* rs_buf;                  // [rows, G] side buffer of the in-kernel resample
  unsigned long long* trace;
  StepTrig trig;
  StepPhase ph[STEP_MAX_PHASES];
};

// Spin with relaxed loads (an acquire load drags a CCTL.IVALL -- an L1 invalidate -- through the SM on every
// iteration); one acquire load after the condition holds orders the subsequent reads.
__device__ __forceinline__ void poll_ge(const int* flag, int target, int item) {
  // fast path: the flag is usually at target already (the tiles of a row group publish close together, and the
  // producer lane gets here one k-block at a time): ONE acquire load instead of relaxed load + acquire load
  if (ptx::ld_acquire_gpu(flag) >= target) return;
  if (ptx::ld_relaxed_gpu(flag) < target) {
    const long long t0 = clock64();
    while (ptx::ld_relaxed_gpu(flag) < target) {
      if (clock64() - t0 > 4000000000LL) {
        printf("dvg_b200: dependency wait timed out (block %d item %d)\n", (int)blockIdx.x, item);
        __trap();
      }
    }
  }
  (void)ptx::ld_acquire_gpu(flag);
}

// Head tile epilogue  y = tanh(acc + b)  of one warp (see the call site).  Deliberately NOT inlined: the step kernel
// calls it once per head tile, at the very end of the step's dependency chain, where its instructions were cold in
// the instruction caches (a 128 x 96 tile took 2.3 us); every epilogue warp therefore also runs it once as a dry run
// (n_rows = 0: nothing is stored) in the idle window before its first accumulator is ready, which only works if both
// calls execute the same code.
__device__ __noinline__ void head_tanh_tile(uint32_t tacc, int n_tile, int sub, int lane, const float* sb, uint8_t* hb,
                                            float* yout, int ldy, int n_valid, int n_rows, int row_w0) {
  const int ncw = n_tile / STEP_NSUB;          // columns of this warp (n_tile is a multiple of 32)
  const int c_begin = sub * ncw;
  const bool vec2 = (ldy & 1) == 0 && (n_valid & 1) == 0 && (reinterpret_cast<uintptr_t>(yout) & 7) == 0;
  const int lr = lane >> 2, cc = (lane & 3) * 2;
  uint32_t cur[8], nxt[8];
  ptx::tmem_ld8(tacc + c_begin, cur);
  ptx::tmem_ld_wait8(cur);
#pragma unroll 1
  for (int c8 = 0; c8 < ncw; c8 += 8) {
    const int cn = c8 + 8 < ncw ? c8 + 8 : c8;   // last trip: harmless re-read
    ptx::tmem_ld8(tacc + c_begin + cn, nxt);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)                   // XU bound: every other exponential goes to the FMA pipe
      v[i] = (i & 1) ? tanh_fast_prescaled_poly(__uint_as_float(cur[i]), sb[c_begin + c8 + i])
                     : tanh_fast_prescaled(__uint_as_float(cur[i]), sb[c_begin + c8 + i]);
    *reinterpret_cast<float4*>(hb + lane * 32) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(hb + lane * 32 + 16) = make_float4(v[4], v[5], v[6], v[7]);
    __syncwarp();
    const int col = c_begin + c8 + cc;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = i * 8 + lr;
      const float2 t = *reinterpret_cast<const float2*>(hb + rr * 32 + cc * 4);
      const int grow = row_w0 + rr;
      if (grow < n_rows && col < n_valid) {
        float* dst = yout + (size_t)grow * ldy + col;
        if (vec2) __stcs(reinterpret_cast<float2*>(dst), t);
        else { dst[0] = t.x; if (col + 1 < n_valid) dst[1] = t.y; }
      }
    }
    __syncwarp();
    ptx::tmem_ld_wait8(nxt);
#pragma unroll
    for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
  }
}

// Dependency wait of a consumer item: k-block `j` of flags[0..n) must have reached `target`.
// Traced cost of the round-1 form (per k-block: relaxed poll, ld.acquire, fence.proxy.async, each a dependent L2 round
// trip or worse -- the proxy fence alone ~0.45 us) was ~1.2 us per k-block on the producer lane although the four
// producing tiles of a row group publish within ~1 us of each other (profiles/r02_step_trace.md).  Now:
//   * DVG_STEP_WFENCE: the generic->async proxy fence sits on the WRITER side (every epilogue thread fences its own
//     stores of the packed h' image before the CTA barrier that precedes the release) -- it is on the causality path
//     from the writes to our TMA reads either way, but there it is paid once per tile, in parallel, not per k-block
//     on the consumer's serial path;
//   * DVG_STEP_POLL_BATCH: the flags of up to four k-blocks are sampled per poll iteration (independent relaxed
//     loads, one round trip), the ones found at target are acquired together (independent ld.acquire, one round
//     trip) and remembered in `ready`, so their k-blocks skip the wait entirely.
__device__ __forceinline__ void poll_deps(const int* flags, int n, int j, int target, uint32_t& ready, int item) {
  if ((ready >> j) & 1u) return;
#if !DVG_STEP_POLL_BATCH
  poll_ge(flags + j, target, item);
#else
  const long long t0 = clock64();
  for (;;) {
    int v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = j + q < n ? ptx::ld_relaxed_gpu(flags + j + q) : 0;
    if (v[0] >= target) {
      int a[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) a[q] = v[q] >= target ? ptx::ld_acquire_gpu(flags + j + q) : 0;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (a[q] >= target) ready |= 1u << (j + q);
      break;
    }
    if (clock64() - t0 > 4000000000LL) {
      printf("dvg_b200: dependency wait timed out (block %d item %d k-block %d)\n", (int)blockIdx.x, item, j);
      __trap();
    }
  }
#endif
#if !DVG_STEP_WFENCE
  ptx::fence_proxy_async_all();       // generic-proxy writes of the producing pairs -> visible to our TMA
#endif
}

// CH: the launch belongs to a chain (dvg_lstm_chain_begin/_end).  Two instantiations, so that the stream-ordered kernel
// is compiled without a trace of the chain code: the kernel sits at its register cap and the tile epilogue's code
// generation reacts to every addition (a uniform `if (in_chain)` around the c' stores alone cost the stream-ordered
// step 1 us).
template <bool CH>
__global__ void __launch_bounds__(STEP_THREADS, 1) lstm_step_kernel(const __grid_constant__ StepArgs p) {
  const bool in_chain = CH;
  const bool chained = CH && p.chained != 0;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = ptx::smem_u32(smem_raw);
  if ((base & 1023u) != 0) {
    if (threadIdx.x == 0) printf("dvg_b200: dynamic shared memory base is not 1024-byte aligned\n");
    __trap();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int CM = 2;
  const uint32_t rank = ptx::cluster_ctarank();
  const uint32_t stage_bytes = p.stage_bytes;
  const uint32_t nparts = p.nsplit == 1 ? 1u : 2u;
  const uint32_t a_bytes = nparts * (uint32_t)TC_A_IMG;
  const uint32_t bar_base = base + (uint32_t)p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STEP_MAX_STAGES + s); };
  auto pfull_bar = [&](int s) { return bar_base + 8u * (2 * STEP_MAX_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (3 * STEP_MAX_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (3 * STEP_MAX_STAGES + 2 + a); };
  auto xready_bar = [&](int j) { return bar_base + 8u * (3 * STEP_MAX_STAGES + 4 + j); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * STEP_MAX_STAGES + 4 + STEP_XMAX);
  uint8_t* tail = smem_raw + (size_t)p.stages * stage_bytes;
  int* s_misc = reinterpret_cast<int*>(tail + 8 * (3 * STEP_MAX_STAGES + 4 + STEP_XMAX) + 16);     // 4 words
  float* s_bias = reinterpret_cast<float*>(tail + STEP_BAR_BYTES);                // [2][256] floats
  uint8_t* s_ebuf = tail + STEP_BAR_BYTES + 2 * 256 * sizeof(float);             // [STEP_EW][32 rows x STEP_RB] (32 KB)

  const int ncl = (int)ptx::cluster_count_x();
  const int cid = ((int)ptx::cluster_id_x() + p.rot) % ncl;
  // k-th item of this pair (-1: none)
  auto item_at = [&](int k) -> int {
    const int pos = cid + k * ncl;
    if (p.sched != nullptr) return pos < p.sched_len ? __ldg(p.sched + pos) : -1;
    return pos < p.total_items ? pos : -1;
  };
  auto phase_of = [&](int item) {
    int k = 0;
    while (k + 1 < p.n_phases && item >= p.ph[k + 1].item_begin) ++k;
    return k;
  };
  // x-pack (epilogue warps): a layer-0 item gets its x rows (this CTA's row tile) as bf16 hi/lo operand images in
  // the item's private scratch slab.  fp32 [rows, G] row-major in, zero padded to the k-steps the MMA reads;
  // (row, 8-column chunk) units are spread so that a warp reads contiguous memory, and all loads of a thread are
  // in flight before the first use (the latents come from HBM).  Ends with the generic->async proxy fence and a
  // barrier of the epilogue warps; the caller then arrives on the item's x-ready mbarrier.
  constexpr int XB = STEP_EW == 8 ? 8 : 4;   // (row, chunk) units per thread and batch
  auto x_load = [&](int item, int u0, float (&v)[XB][8]) {
    const StepPhase& f = p.ph[0];
    const int etid = threadIdx.x - 64;
    const int nchunks = f.in_ksteps * 2;
    const int units = TC_ROWS * nchunks;
    const bool vec2 = (p.ldx & 1) == 0 && (p.G & 1) == 0 && (reinterpret_cast<uintptr_t>(p.x) & 7) == 0;
    const int rg = item / f.n_tiles;
    const int rt = rg * CM + (int)rank;
#pragma unroll
    for (int i = 0; i < XB; ++i) {
      const int u = u0 + etid + i * (STEP_EW * 32);
      const int r = u / nchunks, chunk = u - r * nchunks;
      const int row = rt * TC_ROWS + r;
      const float* src = p.x + (size_t)row * p.ldx + chunk * 8;
      const bool ok = u < units && row < p.rows;
      if (vec2) {
#pragma unroll
        for (int e = 0; e < 8; e += 2) {
          float2 t = make_float2(0.f, 0.f);
          if (ok && chunk * 8 + e < p.G) t = __ldg(reinterpret_cast<const float2*>(src + e));
          v[i][e] = t.x; v[i][e + 1] = t.y;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[i][e] = (ok && chunk * 8 + e < p.G) ? __ldg(src + e) : 0.f;
      }
    }
  };
  auto x_store = [&](int item, int u0, const float (&v)[XB][8]) {
    const StepPhase& f = p.ph[0];
    const int etid = threadIdx.x - 64;
    const int nchunks = f.in_ksteps * 2;
    const int units = TC_ROWS * nchunks;
    const int rg = item / f.n_tiles, nt = item - rg * f.n_tiles;
    const int rt = rg * CM + (int)rank;
    if (rt >= p.row_tiles) return;
    uint8_t* slab = p.xp + (size_t)(nt * p.row_tiles + rt) * f.kb_in * (2u * TC_A_IMG);
#pragma unroll
    for (int i = 0; i < XB; ++i) {
      const int u = u0 + etid + i * (STEP_EW * 32);
      if (u < units) {
        const int r = u / nchunks, chunk = u - r * nchunks;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split2_bf16(v[i][2 * e], v[i][2 * e + 1], hi[e], lo[e]);
        uint8_t* img = slab + (size_t)(chunk >> 3) * (2u * TC_A_IMG);
        const uint32_t o = sw128_offset((uint32_t)r, (uint32_t)(chunk & 7));
        *reinterpret_cast<uint4*>(img + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (nparts == 2) *reinterpret_cast<uint4*>(img + TC_A_IMG + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
  };
  const int x_units = TC_ROWS * p.ph[0].in_ksteps * 2;
  // pack batches [u_begin, units) of an item, then the generic->async proxy fence and a barrier of the epilogue warps
  auto pack_x = [&](int item, int u_begin) {
    for (int u0 = u_begin; u0 < x_units; u0 += XB * STEP_EW * 32) {
      float v[XB][8];
      x_load(item, u0, v);
      x_store(item, u0, v);
    }
    ptx::fence_proxy_async_all();
    ptx::named_bar_sync(1, STEP_EW * 32);
  };
  if (threadIdx.x == 0) {
    TRACE(0);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(full_bar(s), 2);        // activation producer + weight producer, each with its own expect_tx
      ptx::mbar_init(empty_bar(s), 1);
      ptx::mbar_init(pfull_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), 2 * STEP_EW);
    }
    for (int j = 0; j < STEP_XMAX; ++j) ptx::mbar_init(xready_bar(j), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc2(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish2();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // PDL: everything above (barrier init, TMEM allocation, cluster sync) may overlap the tail of the previous kernel
  // in the stream (normally the previous time step, whose last pairs finish ~10 us after the first); nothing below
  // may run before that grid has completed: it produced our state, and it resets the dependency counters on exit.
  ptx::griddep_launch_dependents();
  if (!chained) {
    ptx::griddep_wait();
  } else {
    if (threadIdx.x == 0) {
      // At most TWO launches of a chain may be in flight: the counter sets rotate mod 3 and the x slabs / fired lists
      // alternate.  "My SM is free" does not prove that the launch before the previous one is gone (its stragglers can
      // outlive the first CTAs of the previous launch: measured, launches 3 / 5 of a chain), so it is checked.
      // (Polling from an idle lane at kernel entry, to hide the round trip behind the set-up, measured 0.8 us slower.)
      poll_ge(p.retired, p.chain_idx - 1, -7);
      if (p.prev_trig) {
        poll_ge(p.prev_mask_ready, 1, -4);
        if (p.prev_restore && *reinterpret_cast<const volatile int*>(p.prev_trig_count) > 0)
          poll_ge(p.prev_fin, (int)gridDim.x, -5);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) TRACE(1);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs of the pair) =====================
    if (lane == 0) {
      int s = 0, pm = 0, xj = 0;
#ifdef DVG_TRACE
      int dep_item = -1;
#endif
      uint32_t phs = 0;
      // L2 eviction priorities: the packed h images of the previous step and the packed x slab are read exactly once
      // per consumer and are dead afterwards (the state blocks ping-pong) -> evict first; the weights are re-read by
      // every row group of every step -> evict last.  (Measured neutral on kth_s100, where both state blocks and the
      // weights already stay in the 126 MB L2; it matters when the caller's conv nets stream through L2 in between.)
      const uint64_t pol_stream = ptx::l2_policy_evict_first();
      for (int k = 0;; ++k) {
        const int item = item_at(k);
        if (item < 0) break;
        const int pi = phase_of(item);
        const StepPhase& f = p.ph[pi];
        if (pm < 3) TRACE(2 + pm * 8 + 0);
        const int j = item - f.item_begin;
        const int rg = j / f.n_tiles, nt = j % f.n_tiles;
        int rt = rg * CM + (int)rank;
        if (rt >= p.row_tiles) rt = p.row_tiles - 1;      // padding CTA of an odd last group: loads stay in bounds
        // (phase fields are copied to registers once per item: the phase index is dynamic, and register-indexed
        //  constant-bank loads inside the k-block loops were ~100 cycles each on the issue path)
        const int kb_rec = f.kb_rec, kb_in = f.kb_in;
        const int KB = kb_rec + kb_in;
        const int* wait_flags = f.wait_flags;
        const bool layer0 = wait_flags == nullptr;
        const int* prev_done = chained ? f.prev_done : nullptr;
        uint32_t ready_rec = 0;
        const uint8_t* a_rec = f.a_rec;
        const uint8_t* a_in = f.a_in;
        if (layer0) a_in += (size_t)nt * p.row_tiles * kb_in * (2u * TC_A_IMG);
        uint32_t ready = 0;                                // input k-blocks whose dependency has been acknowledged
#ifdef DVG_TRACE
        if (dep_item < 0 && !layer0) dep_item = pm;        // first dependent item of this CTA: poll phases are traced
#endif
        for (int i = 0; i < KB; ++i) {
          const bool rec = i < kb_rec;
          const int kb = rec ? i : i - kb_rec;
          // (the weight half of the stage is requested by the weight producer warp as soon as the stage is free --
          //  it never depends on this launch -- so half of the stage's bytes are already in flight while we wait for
          //  the activation k-block: dependency counter / x-pack)
          ptx::mbar_wait(empty_bar(s), phs ^ 1u);
          const uint32_t sa = base + (uint32_t)s * stage_bytes;
          if (!rec) {
            if (!layer0) {
#ifdef DVG_TRACE
              if (pm == dep_item && kb < 4) {
                TRACE(112 + kb * 4 + 0);
                const int* fl = wait_flags + rg * kb_in + kb;
                while (ptx::ld_relaxed_gpu(fl) < 2) {}
                TRACE(112 + kb * 4 + 1);
                const int acq = ptx::ld_acquire_gpu(fl);
                if (acq >= 2) TRACE(112 + kb * 4 + 2);
#if !DVG_STEP_WFENCE
                ptx::fence_proxy_async_all();
#endif
                TRACE(112 + kb * 4 + 3);
              } else
#endif
              poll_deps(wait_flags + rg * kb_in, kb_in, kb, 2, ready, item);
            } else if (kb == 0) {
              ptx::mbar_wait(xready_bar(xj), 0);   // our own epilogue warps packed this item's x rows
              ptx::fence_proxy_async_all();
              ++xj;
            }
            if (pm < 3 && kb == 0) TRACE(2 + pm * 8 + 1);
          } else if (prev_done != nullptr) {
            poll_deps(prev_done + rg * kb_rec, kb_rec, kb, 2, ready_rec, item);
          }
          const uint8_t* asrc = rec ? a_rec + (size_t)(rt * kb_rec + kb) * (2u * TC_A_IMG)
                                    : a_in + (size_t)(rt * kb_in + kb) * (2u * TC_A_IMG);
          ptx::mbar_expect_tx(full_bar(s), a_bytes);
          if (rec || layer0) ptx::bulk_g2s_hint(sa, asrc, a_bytes, full_bar(s), pol_stream);
          else ptx::bulk_g2s(sa, asrc, a_bytes, full_bar(s));        // h' of the layer below: re-read by every N tile
          if (pm < 3 && i < 8) TRACE(88 + pm * 8 + i);
          if (++s == p.stages) { s = 0; phs ^= 1u; }
        }
        if (pm < 3) TRACE(2 + pm * 8 + 7);
        ++pm;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int s = 0, mit = 0;
      uint32_t phs = 0;
      for (int k = 0;; ++k, ++mit) {
        const int item = item_at(k);
        if (item < 0) break;
        const StepPhase& f = p.ph[phase_of(item)];
        const int kb_rec = f.kb_rec, in_ksteps = f.in_ksteps;
        const int KB = kb_rec + f.kb_in;
        if (rank == 0) {
          // ===================== MMA issuer (leader CTA) =====================
          const uint32_t idesc = ptx::make_idesc_bf16(2 * TC_ROWS, f.n_tile);
          const uint32_t b_half = (uint32_t)f.n_tile * 64u;
          const int acc = mit & 1;
          const uint32_t aph = (uint32_t)(mit >> 1) & 1u;
          ptx::mbar_wait(tempty_bar(acc), aph ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_STRIDE);
          uint32_t accum = 0;
          for (int i = 0; i < KB; ++i) {
            int ks = TC_KBLK / 16;
            if (i >= kb_rec) {
              const int left = in_ksteps - (i - kb_rec) * (TC_KBLK / 16);
              ks = left < ks ? left : ks;
            }
            ptx::mbar_wait(full_bar(s), phs);
            ptx::mbar_wait(pfull_bar(s), phs);
            if (mit < 3 && i == 0) TRACE(2 + mit * 8 + 2);
            if (mit < 3 && i < 8) TRACE(64 + mit * 8 + i);
            ptx::tc_fence_after();
            const uint32_t sa = base + (uint32_t)s * stage_bytes;
            const uint64_t a_hi = ptx::make_sw128_desc(sa);
            const uint64_t a_lo = ptx::make_sw128_desc(sa + TC_A_IMG);
            const uint64_t b_hi = ptx::make_sw128_desc(sa + a_bytes);
            const uint64_t b_lo = ptx::make_sw128_desc(sa + a_bytes + b_half);
            for (int kk = 0; kk < ks; ++kk) {
              const uint64_t adv = (uint64_t)(kk * 2);
              ptx::umma2_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc, accum);
              accum = 1u;
              if (nparts == 2) {
                ptx::umma2_bf16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                ptx::umma2_bf16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
              }
            }
            ptx::umma2_commit_mcast(empty_bar(s), 3);
            if (++s == p.stages) { s = 0; phs ^= 1u; }
          }
          ptx::umma2_commit_mcast(tfull_bar(acc), 3);
          if (mit < 3) TRACE(2 + mit * 8 + 3);
        } else {
          // ===================== relay (peer CTA): "my stage landed" -> leader's pfull =====================
          for (int i = 0; i < KB; ++i) {
            ptx::mbar_wait(full_bar(s), phs);
            ptx::mbar_arrive_remote(pfull_bar(s), 0);
            if (++s == p.stages) { s = 0; phs ^= 1u; }
          }
        }
      }
    }
  } else if (warp == WPROD_WARP) {
    // ===================== weight-stream producer (both CTAs of the pair) =====================
    // Weights never depend on anything produced in this launch: their half of every stage is requested the moment the
    // stage is free, by a lane of its own, so the dependency polls / x-ready waits of the activation producer (each a
    // chain of L2 round trips) neither delay them nor are delayed by their issue cost.
    if (lane == 0) {
      int s = 0;
      uint32_t phs = 0;
      const uint64_t pol_keep = ptx::l2_policy_evict_last();   // re-read by every row group of every step
      for (int k = 0;; ++k) {
        const int item = item_at(k);
        if (item < 0) break;
        const StepPhase& f = p.ph[phase_of(item)];
        const int j = item - f.item_begin;
        const int nt = j % f.n_tiles;
        const int kb_rec = f.kb_rec, kb_in = f.kb_in;
        const int KB = kb_rec + kb_in;
        const uint32_t b_half = (uint32_t)f.n_tile * 64u, b_part = (uint32_t)f.n_tile * 128u;
        const uint8_t* wbase = f.w + (size_t)nt * KB * (2u * b_part) + rank * b_half;
        for (int i = 0; i < KB; ++i) {
          const int wk = i < kb_rec ? kb_in + i : i - kb_rec;     // weight K order: [input | recurrent]
          ptx::mbar_wait(empty_bar(s), phs ^ 1u);
          const uint32_t sb = base + (uint32_t)s * stage_bytes + a_bytes;
          const uint8_t* bsrc = wbase + (size_t)wk * (2u * b_part);
          ptx::mbar_expect_tx(full_bar(s), nparts * b_half);
          ptx::bulk_g2s_hint(sb, bsrc, b_half, full_bar(s), pol_keep);
          if (nparts == 2) ptx::bulk_g2s_hint(sb + b_half, bsrc + b_part, b_half, full_bar(s), pol_keep);
          if (++s == p.stages) { s = 0; phs ^= 1u; }
        }
      }
    }
  } else if (warp == AUX_WARP) {
    // ===================== GP trigger finalisers (generate_frames.py:283-289) =====================
    // The auxiliary warps of the last NF CTAs take one rollout per lane (window slide, threshold, decision) once
    // the D trigger CTAs have delivered their variances; the last of them publishes mask_ready.
    if (p.trig.enabled) {
      const StepTrig& g = p.trig;
      const int nblk = (g.S + 31) / 32;
      const int NF = nblk < (int)gridDim.x ? nblk : (int)gridDim.x;
      // (chained launches: the finaliser roles go to the first NF CTAs to get here -- the CTAs start as the previous
      //  launch's CTAs exit, spread over ~20 us, and the mask should not wait for the stragglers)
      int fi = (int)gridDim.x - 1 - (int)blockIdx.x;
      if (chained) {
        if (lane == 0) fi = atomicAdd(p.fin_claim, 1);
        fi = __shfl_sync(0xffffffffu, fi, 0);
      }
      if (fi < NF) {
        if (lane == 0) {
          const long long t0 = clock64();
          while (ptx::ld_relaxed_gpu(reinterpret_cast<const int*>(g.ticket)) < g.D) {
            __nanosleep(100);
            if (clock64() - t0 > 4000000000LL) {
              printf("dvg_b200: trigger ticket wait timed out\n");
              __trap();
            }
          }
          (void)ptx::ld_acquire_gpu(reinterpret_cast<const int*>(g.ticket));
        }
        __syncwarp();
        __threadfence();
        const int cnt = *reinterpret_cast<volatile int32_t*>(g.count);
        for (int sidx = fi * 32 + lane; sidx < g.S; sidx += NF * 32) {
          if (g.W <= 16)
            gp_trig_finalize_rollout16(sidx, g.S, g.D, g.var_rows, g.window, g.W, cnt, g.warmup, g.factor, g.value,
                                       g.thr, g.mask, g.trig_list, g.trig_count);
          else
            gp_trig_finalize_rollout(sidx, g.S, g.D, g.var_rows, g.window, g.W, cnt, g.warmup, g.factor, g.value,
                                     g.thr, g.mask, g.trig_list, g.trig_count);
        }
        __threadfence();
        __syncwarp();
        if (lane == 0 && atomicAdd(g.ticket + 1, 1u) == (unsigned)NF - 1u) {
          g.ticket[0] = 0;
          g.ticket[1] = 0;
          if (g.warmup && cnt < g.W) g.count[0] = cnt + 1;
          __threadfence();
          atomicExch(p.mask_ready, 1);
          TRACE(31);
        }
      }
    }
    // Decision steps: the auxiliary lane of EVERY CTA fetches the fired count as soon as the mask is published (long
    // before the CTA's last tile), so the end-of-kernel restore check reads shared memory instead of paying three
    // dependent global round trips (fence, mask poll, count load ~1.3 us) on the tail of every decision step.
    if (p.restore && lane == 0) {
      const long long t0 = clock64();
      while (ptx::ld_relaxed_gpu(p.mask_ready) < 1) {
        __nanosleep(200);
        if (clock64() - t0 > 4000000000LL) {
          printf("dvg_b200: mask wait timed out (block %d)\n", (int)blockIdx.x);
          __trap();
        }
      }
      (void)ptx::ld_acquire_gpu(p.mask_ready);
      s_misc[0] = *reinterpret_cast<volatile int*>(p.trig.trig_count);
    }
  } else {
    // ===================== epilogue / SIMT worker warps (2 .. 2+STEP_EW-1) =====================
    const int ew = warp - 2;
    const int q = warp & 3;
    const int half = ew >> 2;
    const uint32_t r_in_tile = (uint32_t)(q * 32 + lane);
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const int etid = ew * 32 + lane;
#if DVG_STEP_WARM
    // instruction-cache warm-up of the head epilogue (see head_tanh_tile): reads this warp's (unwritten) TMEM lanes and
    // whatever the bias staging holds, stores nothing
    if (p.ph[p.n_phases - 1].type == PH_TANH)
      head_tanh_tile(tmem_base + tlane, p.ph[p.n_phases - 1].n_tile, ew >> 2, lane, s_bias, s_ebuf + ew * (32 * STEP_RB),
                     p.ph[p.n_phases - 1].y, p.ph[p.n_phases - 1].ldy, p.ph[p.n_phases - 1].n_valid, 0, 0);
#endif
    if (p.trig.enabled && chained && etid == 0) s_misc[3] = atomicAdd(p.dim_ctr, 1);   // see trigger_partials
    // x-pack of every layer-0 item of this pair, in item order
    {
      int xj = 0;
      for (int k = 0; xj < STEP_XMAX; ++k) {
        const int item = item_at(k);
        if (item < 0) break;
        if (item >= p.groups * p.ph[0].n_tiles) continue;     // not a layer-0 item
        pack_x(item, 0);
        if (etid == 0) ptx::mbar_arrive(xready_bar(xj));
        ++xj;
      }
      if (etid == 0) TRACE(27);
    }
    // ---- GP variance trigger (generate_frames.py:227-232,275): the last D CTAs evaluate one latent dim for every
    //      rollout: FOUR threads per (rollout, dim) task -- |Linv k|^2 and |L_q^T k|^2, each split into two row halves
    //      -- with the factors staged in the idle transpose buffers.  Nobody in this launch waits for the result
    //      except the finalisers and the end-of-kernel restore / resample; running it in the idle window before the
    //      first accumulator (STEP_TRIG_EARLY) makes the mask known ~15 us earlier, so a fired step's resample
    //      problems start as soon as the first pairs finish.
    //      (small grids -- fewer CTAs than latent dims -- take several dims per CTA)
    auto trigger_partials = [&]() {
      if (!p.trig.enabled) return;
      const StepTrig& g = p.trig;
      // Stream-ordered launches: the last D CTAs take one dim each (with the default item order those pairs have the
      // lightest tile load; measured 43.8 vs 49.3 us per step against first-come claiming).  Chained launches: dims are
      // CLAIMED (atomic counter, up to `quota` per CTA) in the order the CTAs get here -- the CTAs start as the previous
      // launch's CTAs exit, spread over ~20 us, and the next launch waits for this launch's mask: it must not wait for
      // the dims of the last CTAs to start (measured 41.4 -> 38.1 us per step).
      const int trig_ctas = g.D < (int)gridDim.x ? g.D : (int)gridDim.x;
      const int quota = (g.D + (int)gridDim.x - 1) / (int)gridDim.x;
      const int MP = g.mp;
      constexpr int TPQ = STEP_EW * 8;              // threads per quarter
      float* s_linv = reinterpret_cast<float*>(s_ebuf);
      float* s_lqt = s_linv + MP * MP;
      float* s_z = s_lqt + MP * MP;
      float* s_part = s_z + MP;                     // [3][TPQ]
      const int qt = etid / TPQ, li = etid % TPQ;
      for (int q = 0; q < quota; ++q) {
        if (chained && q > 0 && etid == 0) s_misc[3] = atomicAdd(p.dim_ctr, 1);   // (first claim: before the x-pack)
        ptx::named_bar_sync(1, STEP_EW * 32);       // every warp is done with its transpose buffer / the last dim
        const int d = chained ? s_misc[3] : (int)gridDim.x - 1 - (int)blockIdx.x + q * trig_ctas;
        if (d >= g.D) break;
        {
          const float4* g1 = reinterpret_cast<const float4*>(g.linv + (size_t)d * MP * MP);
          const float4* g2 = reinterpret_cast<const float4*>(g.lqt + (size_t)d * MP * MP);
          for (int e = etid; e < MP * MP / 4; e += STEP_EW * 32) {
            reinterpret_cast<float4*>(s_linv)[e] = __ldg(g1 + e);
            reinterpret_cast<float4*>(s_lqt)[e] = __ldg(g2 + e);
          }
          for (int e = etid; e < MP; e += STEP_EW * 32) s_z[e] = g.z[(size_t)d * MP + e];
        }
        const float ell = g.hyp[d * 4 + 0], sc = g.hyp[d * 4 + 1], noise = g.hyp[d * 4 + 3];
        float xv[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {               // latents come from HBM: issue the loads of up to 4 rounds first
          const int i = b * TPQ + li;
          xv[b] = i < g.S ? __ldg(p.x + (size_t)g.stat_rows[i] * p.ldx + d) : 0.f;
        }
        ptx::named_bar_sync(1, STEP_EW * 32);
        for (int base_s = 0; base_s < g.S; base_s += TPQ) {
          const int i = base_s + li;
          const int b = base_s / TPQ;
          const float xi = b < 4 ? (b == 0 ? xv[0] : b == 1 ? xv[1] : b == 2 ? xv[2] : xv[3])
                                 : (i < g.S ? __ldg(p.x + (size_t)g.stat_rows[i] * p.ldx + d) : 0.f);
          float part = 0.f;
          if (i < g.S) {
            if (MP == 40) {
              float kk[40];
              gp_trig_kvec<40>(xi, sc, 1.0f / ell, s_z, kk);
              part = qt == 0 ? gp_trig_rows<40, 0, 20>(s_linv, kk, 0, 20, 0.f)
                   : qt == 1 ? gp_trig_rows<40, 0, 40>(s_linv, kk, 20, 40, 0.f)
                   : qt == 2 ? gp_trig_rows<40, 0, 40>(s_lqt, kk, 0, 20, 0.f)
                             : gp_trig_rows<40, 20, 40>(s_lqt, kk, 20, 40, 0.f);
            } else if ((qt & 1) == 0) {
              part = gp_trig_partial<0>(xi, sc, 1.0f / ell, MP, qt == 0 ? s_linv : s_lqt, s_z, qt != 0);
            }
          }
          if (qt > 0) s_part[(qt - 1) * TPQ + li] = part;
          ptx::named_bar_sync(1, STEP_EW * 32);
          if (qt == 0 && i < g.S)
            g.var_rows[(size_t)d * g.S + i] = (sc - (part + s_part[li])) + (s_part[TPQ + li] + s_part[2 * TPQ + li]) + noise;
          ptx::named_bar_sync(1, STEP_EW * 32);
        }
        if (d == 0 && etid == 0) *g.trig_count = 0;   // ordered before the finalisers by the ticket below
        ptx::named_bar_sync(1, STEP_EW * 32);      // also: the transpose buffers go back to the tile epilogues
        if (etid == 0) {
          __threadfence();       // ONE thread fences at gpu scope: cumulative over the stores the barrier ordered before it
          atomicAdd(g.ticket, 1u);
          TRACE(26);
        }
      }
    };
    if (STEP_TRIG_EARLY) trigger_partials();

    int mit = 0;
    for (int k = 0;; ++k, ++mit) {
      const int item = item_at(k);
      if (item < 0) break;
      const StepPhase& f = p.ph[phase_of(item)];
      const int j = item - f.item_begin;
      const int rg = j / f.n_tiles, nt = j % f.n_tiles;
      const int rt = rg * CM + (int)rank;
      const int acc = mit & 1;
      const uint32_t aph = (uint32_t)(mit >> 1) & 1u;
      const int tm = mit;
      const int row = rt * TC_ROWS + (int)r_in_tile;
      const bool valid = row < p.rows;
      float* sb = s_bias + acc * 256;
      if (etid < f.n_tile) {
        float bv = __ldg(f.bias + (size_t)nt * f.n_tile + etid);
        if (f.type == PH_LSTM) bv *= (etid >> 6) == 2 ? -2.f * kLog2e : -kLog2e;
        else if (f.type == PH_TANH) bv *= -2.f * kLog2e;
        sb[etid] = bv;
      }
      // All fp32 state I/O goes through a warp-private 32-row transpose buffer (STEP_RB bytes per row, XOR-swizzled
      // 16-byte chunks) so that every global access covers whole row segments; the head tiles use 4 KB buffers of
      // the first 8 epilogue warps.
      const int sub = ew >> 2;                             // which STEP_UPW units of the tile this warp owns
      const int half = sub;                                // gaussian head tiles: column half (warps with sub < 2 only)
      uint8_t* eb = s_ebuf + ew * (32 * STEP_RB);
      auto swz = [](int row) { return STEP_RB == 128 ? (row & 7) : ((row >> 1) & 3); };
      const int er = lane / STEP_CPR, ec = lane % STEP_CPR;   // coalesced mapping: (32 / CPR) rows x CPR chunks per instruction
      constexpr int RPI = 32 / STEP_CPR;                   // rows per instruction
      const int row_w0 = rt * TC_ROWS + q * 32;            // first row of this warp
      const int ucol = nt * 64 + sub * STEP_UPW;           // first hidden unit of this warp within the layer
      if (f.type == PH_LSTM) {
        if (chained) {
          if (lane == 0) poll_ge(f.prev_done + rg * f.n_tiles + nt, 2, item);
          __syncwarp();
        }
        float4 cin[STEP_CPR];
#pragma unroll
        for (int i = 0; i < STEP_CPR; ++i) {
          const int rr = i * RPI + er;
          cin[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_w0 + rr < p.rows)
            cin[i] = __ldcs(reinterpret_cast<const float4*>(f.c_in + (size_t)(row_w0 + rr) * p.H + ucol) + ec);
        }
#pragma unroll
        for (int i = 0; i < STEP_CPR; ++i) {
          const int rr = i * RPI + er;
          *reinterpret_cast<float4*>(eb + rr * STEP_RB + ((ec ^ swz(rr)) << 4)) = cin[i];
        }
        __syncwarp();
      }
      ptx::named_bar_sync(1, STEP_EW * 32);
      ptx::mbar_wait(tfull_bar(acc), aph);
      if (etid == 0 && tm < 3) TRACE(2 + tm * 8 + 4);
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + tlane + (uint32_t)(acc * ACC_STRIDE);
      if (f.type == PH_LSTM) {
        const bool held = valid && p.hold != nullptr && p.hold[row / p.rows_per_flag] != 0;
        const size_t idx0 = (size_t)row * p.H + ucol;
        const int cb0 = sub * STEP_UPW;
        // tile columns: [i: 64 units][f: 64][g: 64][o: 64]; this warp handles units cb0 .. cb0+STEP_UPW-1 of its 32
        // rows, 4 units per trip (the loop body must fit the ~6 KB L0 instruction cache: with a larger body the
        // four SM partitions were instruction-fetch bound at ~0.5 IPC).  The next group's accumulators are fetched
        // from TMEM while the current group is computed.  h' (fp32) is parked in the consumed i columns, its bf16
        // hi/lo words in the consumed f columns.
        uint32_t rc[16];
        ptx::tmem_ld4x4(tacc + cb0, tacc + 64 + cb0, tacc + 128 + cb0, tacc + 192 + cb0, rc);
        ptx::tmem_ld_wait16(rc);
#pragma unroll 1
        for (int u = 0; u < STEP_UPW / 4; ++u) {
          const int cb = cb0 + u * 4;
          uint32_t rn[16];
          const int cbn = u < STEP_UPW / 4 - 1 ? cb + 4 : cb;      // last trip: harmless re-read
          ptx::tmem_ld4x4(tacc + cbn, tacc + 64 + cbn, tacc + 128 + cbn, tacc + 192 + cbn, rn);
          float4* cslot = reinterpret_cast<float4*>(eb + lane * STEP_RB + ((u ^ swz(lane)) << 4));
          const float4 c4 = *cslot;
          const float cp[4] = {c4.x, c4.y, c4.z, c4.w};
          float hn[4], cn[4];
          if (held) {
            const float4 h4 = __ldg(reinterpret_cast<const float4*>(f.h_in + idx0 + u * 4));
            hn[0] = h4.x; hn[1] = h4.y; hn[2] = h4.z; hn[3] = h4.w;
#pragma unroll
            for (int i = 0; i < 4; ++i) cn[i] = cp[i];
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              lstm_cell_fast<STEP_NPOLY>(__uint_as_float(rc[i]), __uint_as_float(rc[4 + i]), __uint_as_float(rc[8 + i]),
                             __uint_as_float(rc[12 + i]), sb[cb + i], sb[64 + cb + i], sb[128 + cb + i],
                             sb[192 + cb + i], cp[i], hn[i], cn[i]);
          }
          *cslot = make_float4(cn[0], cn[1], cn[2], cn[3]);
          uint32_t hw[4], sw[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) hw[i] = __float_as_uint(hn[i]);
          split2_bf16(hn[0], hn[1], sw[0], sw[2]);
          split2_bf16(hn[2], hn[3], sw[1], sw[3]);
          ptx::tmem_ld_wait16(rn);                 // rn landed; also orders the stores below after our loads of cb
          ptx::tmem_st4(tacc + cb, hw);
          ptx::tmem_st4(tacc + 64 + cb, sw);       // {hi(0,1), hi(2,3), lo(0,1), lo(2,3)}
#pragma unroll
          for (int i = 0; i < 16; ++i) rc[i] = rn[i];
        }
        ptx::tmem_st_wait();
        if (etid == 0 && tm == 0) TRACE(32);
        {
          // packed h' image of this CTA's 128 rows x 64 units (one k-block of the next GEMM's A operand)
          uint8_t* img = f.hp_out + (size_t)(rt * (p.H / 64) + nt) * (2u * TC_A_IMG);
          uint32_t w[STEP_UPW];
          ptx::tmem_ldw_wait(tacc + 64 + cb0, w);
          if (valid) {
#pragma unroll
            for (int pr = 0; pr < STEP_UPW / 16; ++pr) {
              // units pr*16 .. pr*16+15 of this warp -> 16-byte chunks (chunk0, chunk0 + 1): one aligned 32-byte
              // sector per image, chunk order swapped when bit 0 of (row & 7) is set.  Chunk c (8 units) is made of
              // trips 2c, 2c+1: hi words w[8c+0], w[8c+1], w[8c+4], w[8c+5]; lo words w[8c+2], w[8c+3], w[8c+6], w[8c+7].
              const uint32_t chunk0 = (uint32_t)(sub * (STEP_UPW / 8) + 2 * pr);
              const uint32_t o0 = sw128_offset(r_in_tile, chunk0), o1 = sw128_offset(r_in_tile, chunk0 + 1);
              const bool swap = o1 < o0;
              const uint32_t ob = swap ? o1 : o0;
              const int b0 = 16 * pr, b1 = 16 * pr + 8;
              const uint32_t h0[4] = {w[b0], w[b0 + 1], w[b0 + 4], w[b0 + 5]};
              const uint32_t l0[4] = {w[b0 + 2], w[b0 + 3], w[b0 + 6], w[b0 + 7]};
              const uint32_t h1[4] = {w[b1], w[b1 + 1], w[b1 + 4], w[b1 + 5]};
              const uint32_t l1[4] = {w[b1 + 2], w[b1 + 3], w[b1 + 6], w[b1 + 7]};
              uint32_t th[8], tl[8];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                th[i] = swap ? h1[i] : h0[i]; th[4 + i] = swap ? h0[i] : h1[i];
                tl[i] = swap ? l1[i] : l0[i]; tl[4 + i] = swap ? l0[i] : l1[i];
              }
              st256u(img + ob, th);
              if (nparts == 2) st256u(img + TC_A_IMG + ob, tl);
            }
          }
        }
        // c' rows of this warp -> global (whole row segments).  Stream-ordered launches store them after the publish
        // below (nobody in the launch reads them, and the publish is on the critical path of the launch's own
        // consumers); inside a chain the NEXT launch's tile (rg, nt) reads them as soon as it has seen this tile's
        // counter, so there they go first (+0.45 us per tile; a second "rows stored" counter bumped later -- behind
        // the next CTA barrier, or on the next publish's fence -- measured slower: the next launch does wait for it).
        auto store_c = [&]() {
#pragma unroll
          for (int i = 0; i < STEP_CPR; ++i) {
            const int rr = i * RPI + er;
            const float4 t = *reinterpret_cast<const float4*>(eb + rr * STEP_RB + ((ec ^ swz(rr)) << 4));
            if (row_w0 + rr < p.rows)
              reinterpret_cast<float4*>(f.c_out + (size_t)(row_w0 + rr) * p.H + ucol)[ec] = t;
          }
        };
        if (in_chain) {
          __syncwarp();
          store_c();
        }
        // publish the k-block: consumers (next layer / head) only read the packed image
        // (CTA barrier, then ONE thread fences at gpu scope and bumps the counter: the release is cumulative over
        // the stores the barrier ordered before it -- no per-thread membar)
        if (etid == 0 && tm == 0) TRACE(33);
#if DVG_STEP_WFENCE
        ptx::fence_proxy_async_global();          // our stores of the image -> visible to the consumers' TMA loads
#endif
        ptx::named_bar_sync(1, STEP_EW * 32);
        if (etid == 0) {
          if (tm == 0) TRACE(34);
          __threadfence();
          atomicAdd(f.done_flags + rg * f.n_tiles + nt, 1);
          if (tm == 0) TRACE(28);
        }
        uint32_t hv[STEP_UPW];
        ptx::tmem_ldw_wait(tacc + cb0, hv);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) ptx::mbar_arrive(tempty_bar(acc));
          else ptx::mbar_arrive_remote(tempty_bar(acc), 0);
        }
        if (!in_chain) store_c();
        __syncwarp();
#pragma unroll
        for (int c8 = 0; c8 < STEP_CPR; ++c8)
          *reinterpret_cast<uint4*>(eb + lane * STEP_RB + ((c8 ^ swz(lane)) << 4)) =
              make_uint4(hv[c8 * 4], hv[c8 * 4 + 1], hv[c8 * 4 + 2], hv[c8 * 4 + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < STEP_CPR; ++i) {      // h' tile -> global
          const int rr = i * RPI + er;
          const float4 t = *reinterpret_cast<const float4*>(eb + rr * STEP_RB + ((ec ^ swz(rr)) << 4));
          if (row_w0 + rr < p.rows)     // fp32 h' is only there for the caller's `hidden` views: streaming store
            __stcs(reinterpret_cast<float4*>(f.h_out + (size_t)(row_w0 + rr) * p.H + ucol) + ec, t);
        }
        __syncwarp();
        if (etid == 0 && tm == 0) TRACE(29);
      } else if (f.type == PH_TANH) {
        // y = tanh(acc + b): the STEP_NSUB warps of a TMEM lane quarter split the tile's columns; 8 columns per trip go
        // through a warp-private 32 x 32 B transpose buffer so that the [rows, G] output is written in row segments
        // (4 lanes per row, 8 rows per store instruction).  The next trip's accumulators are fetched from TMEM while
        // the current trip is computed.  (This code runs once per head tile, at the very end of the step's dependency
        // chain: all epilogue warps take part and the loop body is small -- it is cold in the instruction caches.)
        head_tanh_tile(tacc, f.n_tile, sub, lane, sb, s_ebuf + ew * (32 * STEP_RB), f.y, f.ldy, f.n_valid, p.rows, row_w0);
        if (etid == 0 && tm == 2) TRACE(38);
      } else if (f.type == PH_GAUSS && sub < 2) {
        const int nchunks = f.n_tile / 16;
#pragma unroll 1
        for (int jc = half; jc < nchunks; jc += 2) {
          float v[16];
          ptx::tmem_ld16_wait(tacc + jc * 16, v);
          if (valid) {
            const int col0 = nt * f.n_tile + jc * 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += sb[jc * 16 + i];
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
              const int zi = (col0 + i) >> 1;
              if (zi < f.Z) {
                const size_t idx = (size_t)row * f.Z + zi;
                f.mu[idx] = v[i];
                f.logvar[idx] = v[i + 1];
                f.z[idx] = fmaf(f.eps[idx], expf(0.5f * v[i + 1]), v[i]);
              }
            }
          }
        }
      }
      if (f.type != PH_LSTM) {
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) ptx::mbar_arrive(tempty_bar(acc));
          else ptx::mbar_arrive_remote(tempty_bar(acc), 0);
        }
      }
      if (etid == 0 && tm < 3) {
        TRACE(2 + tm * 8 + 5);
#ifdef DVG_TRACE
        if (p.trace) p.trace[(size_t)blockIdx.x * TRACE_SLOTS + 2 + tm * 8 + 6] = 1000000ull + item;
#endif
      }
      if (!STEP_TRIG_EARLY && k == 0) trigger_partials();
    }
  }

  // ---- teardown ----------------------------------------------------------------------------------------------
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TRACE(30);
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc2(tmem_base, TMEM_COLS);
  }
  // Decision step of the fused trigger: rollouts that fired keep their LSTM state (generate_frames.py:289-295).
  // Rare, so it is handled after the fact: once every CTA has finished its items, the state rows of the fired
  // rollouts are copied back from the input block (fp32 h, c and the packed h images), spread over all CTAs.
  int n_fired = 0;
  if (p.restore) {
    n_fired = s_misc[0];
    if (threadIdx.x == 0) TRACE(39);
    if (n_fired > 0) {
      if (threadIdx.x == 0) {
        __threadfence();               // cumulative over this CTA's stores (ordered by the barrier above)
        atomicAdd(p.done_ctr, 1);
      }
      // Phase A -- their decoder input becomes a GP posterior sample of the encoder latent instead of the LSTM
      // prediction (generate_frames.py:291-292).  The (fired rollout, latent dim) problems only need the step's INPUT,
      // so they start as soon as the mask is known, taken from a dynamic queue by whichever CTAs have finished their
      // tiles (most pairs are done ~15 us before the last heads): all 16 epilogue warps on one problem in the now idle operand
      // stages, results into a side buffer because the head tiles of other pairs may still be writing y.
      // (Two problems at a time per CTA -- the two halves of the 16 epilogue warps -- measured slower.)
      const bool rs = p.trig.rs_eps != nullptr;
      const int n_tasks = rs ? n_fired * p.trig.D : 0;
      if (rs) {
        const StepTrig& g = p.trig;
        float* smf = reinterpret_cast<float*>(smem_raw);
        for (;;) {
          if (threadIdx.x == 0) s_misc[1] = atomicAdd(p.rs_next, 1);
          __syncthreads();
          const int task = s_misc[1];
          if (task >= n_tasks) break;
          if (warp >= 2 && warp < 2 + STEP_EW) {    // all epilogue warps on one problem
            const int sr = g.trig_list[task / g.D], d = task % g.D;
            gp_rsample_body<STEP_EW * 32>(smf, (int)threadIdx.x - 64, [] { ptx::named_bar_sync(2, STEP_EW * 32); }, sr, d,
                                          g.n_points, g.D, g.mp, p.x, p.ldx, g.rs_eps, g.z, g.linv, g.lqt, g.alpha, g.hyp,
                                          p.rs_buf, p.G);
          }
          __syncthreads();                          // shared memory and s_misc[1] are reused by the next problem
          if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(p.rs_done, 1);
          }
        }
      }
      if (threadIdx.x == 0) TRACE(40);
      // Phase B -- once every CTA has finished its tiles (and every problem is solved): copy the fired rollouts' state
      // rows (fp32 h, c and the packed h images) back from the input block, and their samples into y.
      if (threadIdx.x == 0) {
        poll_ge(p.done_ctr, (int)gridDim.x, -2);
        if (rs) poll_ge(p.rs_done, n_tasks, -3);
        __threadfence();
        TRACE(41);
      }
      __syncthreads();
      const int chunks = 3 * p.L + (rs ? 1 : 0);
      const int hk = p.H / 64;
      for (int wi = blockIdx.x; wi < n_fired * chunks; wi += gridDim.x) {
        const int s = p.trig.trig_list[wi / chunks];
        const int c = wi % chunks;
        const int r0 = s * p.rows_per_flag;
        const int r1 = r0 + p.rows_per_flag < p.rows ? r0 + p.rows_per_flag : p.rows;
        if (c == 3 * p.L) {                          // samples -> y rows of the rollout
          const int n = (r1 - r0) * p.G;
          for (int i = threadIdx.x; i < n; i += STEP_THREADS) {
            const int r = r0 + i / p.G, gcol = i % p.G;
            p.trig.rs_out[(size_t)r * p.trig.rs_ldo + gcol] = __ldcg(p.rs_buf + (size_t)r * p.G + gcol);
          }
          continue;
        }
        const int l = c / 3, kind = c % 3;
        const StepPhase& f = p.ph[l];
        if (kind < 2) {
          const float4* src = reinterpret_cast<const float4*>((kind == 0 ? f.h_in : f.c_in) + (size_t)r0 * p.H);
          float4* dst = reinterpret_cast<float4*>((kind == 0 ? f.h_out : f.c_out) + (size_t)r0 * p.H);
          const int n4 = (r1 - r0) * p.H / 4;
          for (int i = threadIdx.x; i < n4; i += STEP_THREADS) dst[i] = __ldcg(src + i);
        } else {
          const int per_row = hk * 2 * 8;       // 16-byte units per row: k-blocks x (hi, lo) x 8 chunks
          for (int i = threadIdx.x; i < (r1 - r0) * per_row; i += STEP_THREADS) {
            const int r = r0 + i / per_row, e = i % per_row;
            const int kb = e >> 4, part = (e >> 3) & 1, qd = e & 7;
            const size_t off = ((size_t)((r / TC_ROWS) * hk + kb) * 2 + part) * TC_A_IMG + (size_t)(r % TC_ROWS) * 128 + qd * 16;
            *reinterpret_cast<uint4*>(f.hp_out + off) = __ldcg(reinterpret_cast<const uint4*>(f.a_rec + off));
          }
        }
      }
    }
  }
  // Self-resetting dependency counters: the last CTA to get here (every CTA has finished reading them) zeroes
  // them for the next launch -- no cudaMemset node per step.
  __syncthreads();
  if (threadIdx.x == 0) {
    TRACE(42);
    int* exit_ctr = p.flag_words + p.n_flag_words;
    __threadfence();
    if (n_fired > 0) atomicAdd(p.fin_ctr, 1);       // a chained successor waits for the restored state rows
    if (atomicAdd(exit_ctr, 1) == (int)gridDim.x - 1) {
      if (p.self_reset)
        for (int i = 0; i < p.n_flag_words; ++i) p.flag_words[i] = 0;
      if (p.reset_set != nullptr)
        for (int i = 0; i < p.n_flag_words; ++i) p.reset_set[i] = 0;
      __threadfence();
      *exit_ctr = 0;
      if (in_chain) {
        poll_ge(p.retired, p.chain_idx, -8);        // retire in order (the previous launch's last CTA may be behind us)
        __threadfence();
        atomicExch(p.retired, p.chain_idx + 1);     // launches up to chain_idx are gone, their counters are reset
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------
static bool use_fused() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DVG_TC_FUSED");     // developer switch: 0 = one launch per GEMM
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

static bool use_chain() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DVG_STEP_CHAIN");   // developer switch: 0 = every launch waits for the previous grid
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// dvg_lstm_chain_begin / _end: see include/dvg_b200.h
#ifdef DVG_TRACE
static unsigned long long* g_chain_tbuf = nullptr;     // [16 launches][256 CTAs][TRACE_SLOTS]
static int g_chain_traced = 0;
static void chain_trace_dump(cudaStream_t stream) {
  if (!g_chain_tbuf || g_chain_traced < 2) return;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &cs);
  if (cs != cudaStreamCaptureStatusNone) return;
  cudaStreamSynchronize(stream);
  const char* which = getenv("DVG_TC_TRACE_LAUNCH");
  const int want = which ? atoi(which) : 6;
  const size_t per = 256 * (size_t)TRACE_SLOTS;
  std::vector<unsigned long long> hbuf(16 * per);
  cudaMemcpy(hbuf.data(), g_chain_tbuf, 16 * per * 8, cudaMemcpyDeviceToHost);
  unsigned long long t0 = ~0ull;
  for (int b = 0; b < 256; ++b) { const unsigned long long v = hbuf[(size_t)(want % 16) * per + (size_t)b * TRACE_SLOTS]; if (v && v < t0) t0 = v; }
  for (int li = want; li < want + 3 && li < g_chain_traced; ++li) {
    fprintf(stderr, "STEP TRACE chain launch %d\n", li);
    for (int b = 0; b < 256; ++b) {
      const unsigned long long* r = &hbuf[(size_t)(li % 16) * per + (size_t)b * TRACE_SLOTS];
      if (!r[0]) continue;
      fprintf(stderr, "cta %3d:", b);
      for (int i = 0; i < 128; ++i) {
        const unsigned long long v = r[i];
        if (v >= 1000000ull && v < 2000000ull) fprintf(stderr, " #%lld", (long long)(v - 1000000ull));
        else fprintf(stderr, " %lld", v ? (long long)v - (long long)t0 : -1ll);
      }
      fprintf(stderr, "\n");
    }
  }
  g_chain_traced = 0;
}
#endif

int lstm_step_chain(dvg_lstm_s* h, bool begin, cudaStream_t stream) {
#ifdef DVG_TRACE
  if (!begin) chain_trace_dump(stream);
#endif
  if (h->fused_flags != nullptr && h->chain_dirty) {
    DVG_CUDA(cudaMemsetAsync(h->fused_flags, 0, sizeof(int) * (3 * h->flag_set_words + 8), stream));
    h->chain_dirty = false;
  }
  h->chain_on = begin;
  h->chain_ok = false;
  h->chain_idx = 0;
  return DVG_OK;
}

size_t lstm_step_flag_words(const dvg_lstm_s* h, int rows) {
  const int groups = ceil_div(ceil_div(rows, TC_ROWS), 2), hk = h->dims.hidden_size / 64;
  return (size_t)h->dims.n_layers * align_up((size_t)groups * hk, 32) + 8;
}
size_t lstm_step_xp_bytes(const dvg_lstm_s* h, int rows) {
  return (size_t)(h->dims.hidden_size / 64) * ceil_div(rows, TC_ROWS) * ceil_div(h->dims.input_size, 64) * 2 * TC_A_IMG;
}

// ---------------------------------------------------------------------------------------------------
// Item schedule.  Position i of the order runs on pair i % pairs as its (i / pairs)-th item (-1 = none); every
// item's dependencies sit at smaller positions of the sequence the lists were built from, so the kernel cannot
// deadlock.  With 80 + 80 + 20 items on 74 pairs the layer-major order leaves a three-deep tail (late layer-0
// tiles -> their layer-1 consumers as third items -> heads).  The lists are therefore built by a small list
// scheduler on a cost model calibrated from the timestamp traces (profiles/): layer-0 tiles go round-robin to all
// but d pairs, the remaining items are placed one by one on the pair that completes them earliest, and d is
// chosen by simulating the makespan.
// ---------------------------------------------------------------------------------------------------
namespace {
struct SchedCost {
  double t_first = 3.9, kt = 1.0, kt_head = 0.4, te_lstm = 6.3, t_pub = 5.0, te_head = 2.3, t_dep = 1.6;
};
struct PairSim {
  double mma_free = 0.0, epi_free = 0.0, drain[2] = {0.0, 0.0};
  int n = 0;
  std::vector<int> items;
};
}  // namespace

static double sched_place(PairSim& ps, const SchedCost& c, int layer, int L, int kb_rec, double in_kb, double dep_ready,
                          bool commit, double* publish) {
  const bool head = layer == L;
  double start = ps.n == 0 ? c.t_first : ps.mma_free;
  if (ps.n >= 2 && ps.drain[ps.n & 1] > start) start = ps.drain[ps.n & 1];   // TMEM buffer still being drained
  const double kt = head ? c.kt_head : c.kt;
  const double rec_done = start + kb_rec * kt;
  double in_start = rec_done;
  if (dep_ready + c.t_dep > in_start) in_start = dep_ready + c.t_dep;
  const double acc = in_start + in_kb * kt;
  const double epi_start = acc > ps.epi_free ? acc : ps.epi_free;
  const double te = head ? c.te_head : c.te_lstm;
  const double done = epi_start + te;
  if (commit) {
    ps.mma_free = acc;
    ps.drain[ps.n & 1] = epi_start + (head ? te : c.t_pub);
    ps.epi_free = done;
    ++ps.n;
    if (publish) *publish = epi_start + (head ? te : c.t_pub);
  }
  return done;
}

// returns the makespan; fills lists (per pair) when out != nullptr
static double sched_build(int P, int L, int groups, int hk, double x_kb, int d, const SchedCost& c,
                          std::vector<std::vector<int>>* out) {
  std::vector<PairSim> ps(P);
  const int per_layer = groups * hk;
  std::vector<double> pub((size_t)L * per_layer, 0.0);
  // layer 0: round-robin over pairs d .. P-1
  for (int i = 0; i < per_layer; ++i) {
    PairSim& q = ps[d + i % (P - d)];
    sched_place(q, c, 0, L, hk, x_kb, 0.0, true, &pub[i]);
    q.items.push_back(i);
  }
  auto place_ect = [&](int item, int layer, int kb_rec, double in_kb, double dep) {
    int best = 0;
    double bt = 1e30;
    for (int q = 0; q < P; ++q) {
      const double t = sched_place(ps[q], c, layer, L, kb_rec, in_kb, dep, false, nullptr);
      if (t < bt - 1e-9) { bt = t; best = q; }
    }
    double pb = 0.0;
    sched_place(ps[best], c, layer, L, kb_rec, in_kb, dep, true, &pb);
    ps[best].items.push_back(item);
    return pb;
  };
  for (int l = 1; l <= L; ++l) {
    // consumers of layer l-1, in the order their dependencies become ready
    const int n_items = l < L ? per_layer : groups;
    std::vector<std::pair<double, int>> order;
    for (int j = 0; j < n_items; ++j) {
      const int rg = l < L ? j / hk : j;
      double dep = 0.0;
      for (int nt = 0; nt < hk; ++nt) dep = std::max(dep, pub[(size_t)(l - 1) * per_layer + rg * hk + nt]);
      order.push_back({dep, j});
    }
    std::stable_sort(order.begin(), order.end());
    for (auto& o : order) {
      const int item = l * per_layer + o.second;
      const double pb = place_ect(item, l, l < L ? hk : 0, hk, o.first);
      if (l < L) pub[(size_t)l * per_layer + o.second] = pb;
    }
  }
  double mk = 0.0;
  for (auto& q : ps) mk = std::max(mk, q.epi_free);
  if (out) {
    out->clear();
    for (auto& q : ps) out->push_back(q.items);
  }
  return mk;
}

// Two-layer pattern for "a few more tiles per layer than pairs" (n = P + e, 0 < e <= P/4; kth_s100: 80 tiles, 74 pairs).
// The layer-major order makes e pairs run (L0, L0, L1) and e pairs (L0, L1, L1), and the heads of the last row groups
// wait for third-slot L1 tiles that themselves follow two full tiles: a three-deep tail ~12 us behind the median
// pair.  Here no pair gets (L0, L1, L1):
//   pairs [0, e)        (L1, L1)      layer-1 tiles of the first row groups; their recurrent k-blocks run at once, the
//                                      input k-blocks as soon as the (first-slot) layer-0 tiles publish
//   pairs [e, 3e)       (L0, L0, L1)  the second L0 is one of the LAST 2e layer-0 tiles, the L1 one of the last 2e
//                                      layer-1 tiles (whose inputs are exactly those late L0 tiles)
//   pairs [3e, P)       (L0, L1)      + one head each for the first `groups` of them
// Every layer-0 tile is first or second in its list behind another layer-0 tile, so none can block: no deadlock.
static bool sched_pattern_two_layer(int P, int groups, int hk, std::vector<std::vector<int>>& lists) {
  const int n = groups * hk, e = n - P;
  if (e <= 0 || e > P / 4 || P - 3 * e < 1) return false;
  lists.assign(P, {});
  const int a = 2 * e;
  auto L0 = [&](int i) { return i; };
  auto L1 = [&](int j) { return n + j; };
  auto HEAD = [&](int g) { return 2 * n + g; };
  for (int q = 0; q < e; ++q) { lists[q].push_back(L1(q)); lists[q].push_back(L1(e + q)); }
  for (int i = 0; i < P - e; ++i) lists[e + i].push_back(L0(i));                 // first-slot L0 tiles
  for (int i = 0; i < a; ++i) lists[e + i].push_back(L0(P - e + i));             // the last 2e L0 tiles, second slot
  for (int i = 0; i < a; ++i) lists[e + i].push_back(L1(n - a + i));             // their consumers, third slot
  for (int i = 0; i < P - 3 * e; ++i) lists[3 * e + i].push_back(L1(2 * e + i)); // one L1 per single-L0 pair
  for (int g = 0; g < groups; ++g) lists[3 * e + g % (P - 3 * e)].push_back(HEAD(g));
  return true;
}

int lstm_step_build_schedule(dvg_lstm_s* h, int rows) {
  if (h->sched_dev) { h->retired.push_back(h->sched_dev); h->sched_dev = nullptr; }   // captured graphs may hold it
  h->sched_len = h->sched_rows = h->sched_pairs = 0;
  // DVG_STEP_SCHED: 0 = layer-major identity order, 1 = list scheduler on the cost model (experimental: its cost model
  // dates from round 1 and it measures +1.3 us on kth_s100), 2 / unset = the two-layer pattern above where it applies
  // (a few more tiles per layer than pairs; identity order otherwise).  Measured on kth_s100, round 2: 43.4 vs 43.8 us
  // per plain step, 43.5 vs 44.2 us per decision step.
  const char* e = getenv("DVG_STEP_SCHED");
  const int mode = e ? atoi(e) : 2;       // round 2: the two-layer pattern is the default where it applies (kth_s100 -0.5 us)
  if (mode == 0) return DVG_OK;
  const int L = h->dims.n_layers, hk = h->dims.hidden_size / 64;
  const int RT = ceil_div(rows, TC_ROWS), groups = ceil_div(RT, 2);
  const int total = L * groups * hk + groups;
  int P = h->sm_count / 2;
  if (P > total) P = total;
  if (P < 2 || L != 2 || !h->tc_ok) return DVG_OK;
  std::vector<std::vector<int>> lists;
  double best = 0.0;
  int best_d = -1;
  if (mode == 2) {
    if (!sched_pattern_two_layer(P, groups, hk, lists)) return DVG_OK;
  } else {
    SchedCost c;
    const double x_kb = ceil_div(h->dims.input_size, 16) / 4.0;
    best = 1e30;
    for (int d = 0; d <= P / 3; ++d) {
      const double mk = sched_build(P, L, groups, hk, x_kb, d, c, nullptr);
      if (mk < best - 1e-9) { best = mk; best_d = d; }
    }
    if (const char* fd = getenv("DVG_STEP_SCHED_D")) best_d = std::max(0, std::min(P / 2, atoi(fd)));   // developer override
    best = sched_build(P, L, groups, hk, x_kb, best_d, c, &lists);
  }
  size_t depth = 0;
  int l0_max = 0;
  for (auto& li : lists) {
    depth = std::max(depth, li.size());
    int n0 = 0;
    for (int it : li) n0 += it < groups * hk ? 1 : 0;
    l0_max = std::max(l0_max, n0);
  }
  if (l0_max > STEP_XMAX) return DVG_OK;
  std::vector<int> flat(depth * P, -1);
  size_t placed = 0;
  for (int q = 0; q < P; ++q)
    for (size_t k = 0; k < lists[q].size(); ++k) { flat[k * P + q] = lists[q][k]; ++placed; }
  if ((int)placed != total) return DVG_OK;          // defensive: a malformed schedule would hang the kernel
  DVG_CUDA(cudaMalloc(&h->sched_dev, sizeof(int) * flat.size()));
  DVG_CUDA(cudaMemcpy(h->sched_dev, flat.data(), sizeof(int) * flat.size(), cudaMemcpyHostToDevice));
  h->sched_len = (int)flat.size(); h->sched_rows = rows; h->sched_pairs = P;
  if (getenv("DVG_STEP_SCHED_VERBOSE"))
    fprintf(stderr, "dvg_b200: step schedule mode=%d rows=%d pairs=%d d=%d depth=%zu model makespan %.1f us\n", mode, rows, P,
            best_d, depth, best);
  return DVG_OK;
}

bool lstm_step_usable(const dvg_lstm_s* h, int rows) {
  if (!h->tc_ok || !use_fused()) return false;
  const int RT = ceil_div(rows, TC_ROWS), groups = ceil_div(RT, 2), hk = h->dims.hidden_size / 64;
  const int pairs = h->sm_count / 2;
  if (RT < 1 || pairs < 1 || hk > 16) return false;    // poll_deps samples at most 16 k-block flags per item
  if (lstm_small_usable(h, rows)) return false;        // <= 64 rows at H = 256: the 16-CTA cluster kernel is faster
  return ceil_div(groups * hk, pairs) <= STEP_XMAX;
}

int lstm_step_launch(dvg_lstm_s* h, dvg_gp_s* g, int nsplit, int rows, const float* x, int ldx, const float* h_in,
                     const float* c_in, const uint8_t* hp_in, float* h_out, float* c_out, uint8_t* hp_out, float* y,
                     int ldy, const float* eps, float* z, float* mu, float* logvar, const uint8_t* hold,
                     int rows_per_flag, cudaStream_t stream, const StepTrigHost* trig) {
  const int G = h->dims.input_size, H = h->dims.hidden_size, L = h->dims.n_layers;
  const int hk = H / 64, RT = ceil_div(rows, TC_ROWS), kbx = ceil_div(G, 64);
  const int groups = ceil_div(RT, 2);
  const size_t lsz = (size_t)rows * H;
  const size_t lpk = (size_t)RT * hk * 2 * TC_A_IMG;
  StepArgs a{};
  a.rows = rows; a.row_tiles = RT; a.groups = groups; a.nsplit = nsplit; a.H = H; a.L = L; a.G = G; a.ldx = ldx;
  a.kbx = kbx; a.x = x;
  a.hold = hold; a.rows_per_flag = rows_per_flag > 0 ? rows_per_flag : 1;
  const int fstride = (int)align_up((size_t)groups * hk, 32);
  int pairs = h->sm_count / 2;
  {
    const int total_items = L * groups * hk + groups;
    if (pairs > total_items) pairs = total_items;
  }
  // Chain bookkeeping (see StepArgs::chained).  Inside dvg_lstm_chain_begin/_end the launches rotate through three
  // counter sets and two x slabs / fired lists; a launch is CHAINED to its predecessor (skips the grid dependency
  // wait) when that was a step launch of the same shape on the same stream whose output state is our input state.
  // Needs the grid to cover the machine: only then is "my SM is free" proof that the launch before the previous one
  // has drained completely.
  const bool in_chain = h->chain_on && use_chain() && pairs * 2 == h->sm_count && hold == nullptr;
  if (h->chain_on && !in_chain && h->chain_dirty) {       // not chainable: drop back to the self-resetting protocol
    DVG_CUDA(cudaMemsetAsync(h->fused_flags, 0, sizeof(int) * (3 * h->flag_set_words + 8), stream));
    h->chain_dirty = false; h->chain_ok = false; h->chain_idx = 0;
  }
  const int ci = in_chain ? h->chain_idx : 0;
  const int set = ci % 3, pset = (ci + 2) % 3, par = ci & 1;
  const bool chained = in_chain && h->chain_ok && h->chain_out == (const void*)h_in && h->chain_rows == rows &&
                       h->chain_nsplit == nsplit && h->chain_stream == stream && h->chain_gp == (const void*)g;
  if (getenv("DVG_STEP_CHAIN_VERBOSE"))
    fprintf(stderr, "dvg_b200: step launch rows=%d in_chain=%d idx=%d chained=%d (ok=%d out==in %d rows %d nsplit %d stream %d gp %d) trig=%d\n",
            rows, (int)in_chain, ci, (int)chained, (int)h->chain_ok, (int)(h->chain_out == (const void*)h_in),
            (int)(h->chain_rows == rows), (int)(h->chain_nsplit == nsplit), (int)(h->chain_stream == stream),
            (int)(h->chain_gp == (const void*)g), (int)(trig != nullptr));
  int* flags = h->fused_flags + (size_t)set * h->flag_set_words;
  int* pflags = h->fused_flags + (size_t)pset * h->flag_set_words;
  a.xp = h->tc_xp + (size_t)par * h->xp_stride;
  a.flag_words = flags;
  a.n_flag_words = L * fstride + 7;
  a.mask_ready = flags + (size_t)L * fstride;
  a.done_ctr = a.mask_ready + 1;
  a.rs_next = a.mask_ready + 2;
  a.rs_done = a.mask_ready + 3;
  a.fin_ctr = a.mask_ready + 4;
  a.dim_ctr = a.mask_ready + 5;
  a.fin_claim = a.mask_ready + 6;
  a.rs_buf = h->rs_buf;
  a.chained = chained ? 1 : 0;
  a.in_chain = in_chain ? 1 : 0;
  a.chain_idx = ci;
  a.retired = h->fused_flags + 3 * h->flag_set_words;
  a.self_reset = in_chain ? 0 : 1;
  a.reset_set = in_chain ? pflags : nullptr;
  a.rot = in_chain ? (ci * 13) % pairs : 0;
  a.prev_trig = chained ? h->chain_prev_trig : 0;
  a.prev_restore = chained ? h->chain_prev_restore : 0;
  a.prev_mask_ready = pflags + (size_t)L * fstride;
  a.prev_fin = a.prev_mask_ready + 4;
  a.prev_trig_count = g != nullptr ? g->trig_count + (par ^ 1) : nullptr;
  if (trig != nullptr) {
    StepTrig& t = a.trig;
    t.enabled = 1; t.S = trig->S; t.D = g->dims.num_dims; t.mp = g->mp; t.W = trig->W; t.warmup = trig->warmup;
    t.factor = trig->factor; t.stat_rows = trig->stat_rows;
    t.z = g->z; t.linv = g->linv; t.lqt = g->lqt; t.hyp = g->hyp;
    t.var_rows = g->var_rows; t.ticket = g->ticket; t.window = trig->window; t.count = trig->count;
    t.value = trig->value; t.thr = trig->thr; t.mask = trig->mask;
    t.trig_list = g->trig_list + (size_t)par * g->var_rows_cap;
    t.trig_count = g->trig_count + par;
    t.rs_eps = trig->warmup ? nullptr : trig->rs_eps; t.alpha = g->alpha; t.rs_out = y; t.rs_ldo = ldy;
    t.n_points = rows / trig->S;
    a.restore = trig->warmup ? 0 : 1;
    a.rows_per_flag = rows / trig->S;
    a.hold = nullptr;
  }
  int np = 0, item = 0;
  for (int l = 0; l < L; ++l) {
    StepPhase& f = a.ph[np];
    const TcGemmPlan& pl = l == 0 ? h->tc_layer0f : h->tc_layer[l];
    f.type = PH_LSTM; f.n_tile = 256; f.n_tiles = hk; f.item_begin = item;
    f.kb_in = l == 0 ? kbx : hk; f.kb_rec = hk;
    f.in_ksteps = l == 0 ? ceil_div(G, 16) : hk * 4;
    f.a_in = l == 0 ? a.xp : hp_out + (l - 1) * lpk;
    f.a_rec = hp_in + l * lpk;
    f.w = pl.w; f.bias = pl.bias;
    f.wait_flags = l == 0 ? nullptr : flags + (size_t)(l - 1) * fstride;
    f.done_flags = flags + (size_t)l * fstride;
    f.prev_done = pflags + (size_t)l * fstride;
    f.c_in = c_in + l * lsz; f.h_in = h_in + l * lsz; f.h_out = h_out + l * lsz; f.c_out = c_out + l * lsz;
    f.hp_out = hp_out + l * lpk;
    item += groups * hk; ++np;
  }
  {  // head
    StepPhase& f = a.ph[np];
    const bool gauss = h->dims.kind == DVG_GAUSSIAN_LSTM;
    f.type = gauss ? PH_GAUSS : PH_TANH; f.n_tile = h->tc_head.n_tile; f.n_tiles = 1; f.item_begin = item;
    f.kb_in = hk; f.kb_rec = 0; f.in_ksteps = hk * 4;
    f.a_in = hp_out + (L - 1) * lpk; f.a_rec = nullptr;
    f.w = h->tc_head.w; f.bias = h->tc_head.bias;
    f.wait_flags = flags + (size_t)(L - 1) * fstride;
    f.done_flags = nullptr;
    f.y = y; f.ldy = ldy; f.n_valid = h->dims.output_size;
    f.eps = eps; f.z = z; f.mu = mu; f.logvar = logvar; f.Z = h->dims.output_size;
    item += groups; ++np;
  }
  a.n_phases = np; a.total_items = item;
  const uint32_t nparts = nsplit == 1 ? 1 : 2;
  const size_t stage_bytes = nparts * ((size_t)TC_A_IMG + (size_t)256 * 64);
  const size_t tail = STEP_BAR_BYTES + 2 * 256 * sizeof(float) + (size_t)STEP_EBUF_BYTES;
  int stages = (int)((227 * 1024 - tail) / stage_bytes);
  if (stages > STEP_MAX_STAGES) stages = STEP_MAX_STAGES;
  a.stages = stages; a.stage_bytes = (uint32_t)stage_bytes;
  const size_t smem = stages * stage_bytes + tail;
  static bool configured = false;
  if (!configured) {
    DVG_CUDA(cudaFuncSetAttribute(lstm_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DVG_CUDA(cudaFuncSetAttribute(lstm_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  if (h->sched_dev != nullptr && h->sched_rows == rows && h->sched_pairs == pairs) {
    a.sched = h->sched_dev; a.sched_len = h->sched_len;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(pairs * 2);
  cfg.blockDim = dim3(STEP_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  static int use_pdl = -1;
  if (use_pdl < 0) {
    const char* e = getenv("DVG_STEP_PDL");       // developer switch: 0 = plain stream order
    use_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  cfg.attrs = attr; cfg.numAttrs = use_pdl ? 2 : 1;
#ifdef DVG_TRACE
  static unsigned long long* tbuf = nullptr;
  bool tr = getenv("DVG_TC_TRACE") != nullptr;
  if (tr && in_chain) {          // chained launches: one slab per launch, nothing between the launches, dumped at chain end
    const size_t per = 256 * (size_t)TRACE_SLOTS;
    if (!g_chain_tbuf) cudaMalloc(&g_chain_tbuf, 16 * per * 8);
    if (ci == 0) cudaMemsetAsync(g_chain_tbuf, 0, 16 * per * 8, stream);
    if (ci < 16) { a.trace = g_chain_tbuf + (size_t)ci * per; g_chain_traced = ci + 1; }
    tr = false;
  }
  if (tr) {
    if (!tbuf) cudaMalloc(&tbuf, 256 * TRACE_SLOTS * 8);
    cudaMemsetAsync(tbuf, 0, 256 * TRACE_SLOTS * 8, stream);
    a.trace = tbuf;
  }
#endif
  h->prof_mark(stream);
  if (in_chain) DVG_CUDA(cudaLaunchKernelEx(&cfg, lstm_step_kernel<true>, (const StepArgs)a));
  else DVG_CUDA(cudaLaunchKernelEx(&cfg, lstm_step_kernel<false>, (const StepArgs)a));
  h->prof_mark(stream);
  if (in_chain) {
    h->chain_idx = ci + 1; h->chain_ok = true; h->chain_dirty = true;
    h->chain_out = (const void*)h_out; h->chain_rows = rows; h->chain_nsplit = nsplit; h->chain_stream = stream;
    h->chain_gp = (const void*)g;
    h->chain_prev_trig = trig != nullptr ? 1 : 0; h->chain_prev_restore = a.restore;
  }
#ifdef DVG_TRACE
  if (tr) {
    static int n_dump = 0;
    const char* which = getenv("DVG_TC_TRACE_LAUNCH");
    const int want = which ? atoi(which) : 4;
    cudaStreamSynchronize(stream);
    if (n_dump++ == want) {
      std::vector<unsigned long long> hbuf(256 * TRACE_SLOTS);
      cudaMemcpy(hbuf.data(), tbuf, 256 * TRACE_SLOTS * 8, cudaMemcpyDeviceToHost);
      unsigned long long t0 = ~0ull;
      for (int b = 0; b < (int)cfg.gridDim.x; ++b) if (hbuf[b * TRACE_SLOTS] && hbuf[b * TRACE_SLOTS] < t0) t0 = hbuf[b * TRACE_SLOTS];
      fprintf(stderr, "STEP TRACE grid=%d items=%d stages=%d restore=%d: start setup | per item: pstart depok stage0 mma_issued acc_ready epi_done item requested | trigdone xpack published stored end finalizer\n",
              (int)cfg.gridDim.x, a.total_items, stages, a.restore);
      for (int b = 0; b < (int)cfg.gridDim.x; ++b) {
        fprintf(stderr, "cta %3d:", b);
        for (int i = 0; i < 128; ++i) {
          unsigned long long v = hbuf[b * TRACE_SLOTS + i];
          if (i == 2 || i == 10 || i == 18 || i == 26) fprintf(stderr, " |");
          if (v >= 1000000ull && v < 2000000ull) fprintf(stderr, " #%lld", (long long)(v - 1000000ull));
          else fprintf(stderr, " %lld", v ? (long long)(v - t0) : -1ll);
        }
        fprintf(stderr, "\n");
      }
    }
  }
#endif
  return DVG_OK;
}

bool lstm_tc_can_fuse_trigger(const dvg_lstm_s* h, const dvg_gp_s* g, int rows) {
  const size_t need = sizeof(float) * ((size_t)2 * g->mp * g->mp + g->mp + 3 * STEP_EW * 8);
  const int pairs = h->sm_count / 2;
  return lstm_step_usable(h, rows) && !g->big && need <= (size_t)STEP_EBUF_BYTES && g->dims.num_dims <= pairs * 2;
}
// the in-kernel rsample needs its scratch to fit the stage buffers (3 x 64 KB)
bool lstm_tc_can_fuse_rsample(const dvg_gp_s* g, int n_points) {
  return sizeof(float) * gp_rsample_smem_floats(n_points, g->mp) <= (size_t)3 * 65536 && n_points <= 128;
}

// trigger + LSTM step in one launch; caller guarantees lstm_tc_can_fuse_trigger().
int lstm_tc_rollout_step(dvg_lstm_s* h, dvg_gp_s* g, int nsplit, int rows, const float* x, int ldx, const float* h_in,
                         const float* c_in, const uint8_t* hp_in, float* h_out, float* c_out, uint8_t* hp_out,
                         float* y, int ldy, int S, const int32_t* stat_rows, float* window, int W, int32_t* count,
                         int warmup, float factor, float* value, float* thr, uint8_t* mask, const float* rs_eps,
                         cudaStream_t stream) {
  StepTrigHost t{};
  t.rs_eps = rs_eps;
  t.S = S; t.W = W; t.warmup = warmup; t.factor = factor; t.stat_rows = stat_rows; t.window = window; t.count = count;
  t.value = value; t.thr = thr; t.mask = mask;
  return lstm_step_launch(h, g, nsplit, rows, x, ldx, h_in, c_in, hp_in, h_out, c_out, hp_out, y, ldy, nullptr, nullptr,
                          nullptr, nullptr, nullptr, rows / S, stream, &t);
}

}  // namespace dvg
