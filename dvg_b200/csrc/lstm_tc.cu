// Tensor-core variants (DVG_BF16X3 / DVG_BF16) of the LSTM hot path for sm_100a.
//
// One persistent, warp-specialised GEMM kernel (tc_gemm_kernel) with four fused epilogues:
//   EPI_PACK   embed (models/lstm.py:66): + bias, result written as bf16 hi/lo operand images
//   EPI_LSTM   LSTMCell (models/lstm.py:69): gates -> sigmoid/tanh -> c', h' (fp32) + packed h'
//   EPI_TANH   output Linear + Tanh (models/lstm.py:72)
//   EPI_GAUSS  mu/logvar heads + reparameterize (models/lstm.py:161-164,172-174)
//
// Data layout.  Every GEMM operand lives in HBM as "k-block images": [rows x 64] bf16 tiles stored
// exactly in the tcgen05 shared-memory layout (K-major, 128-byte rows, SWIZZLE_128B).  An fp32 value
// a is carried as a = hi + lo (two bf16 images, adjacent in memory).  Because the images are already
// swizzled, operands move HBM/L2 -> smem with plain 1-D TMA bulk copies (cp.async.bulk, no tensor
// maps) and the epilogue of one GEMM writes the A operand of the next one directly.
//
// DVG_BF16X3: D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi (3 tcgen05.mma per k-step, fp32 accumulate in
// TMEM) -- products carry ~16 mantissa bits, i.e. fp32-grade for the <=1e-4 parity bar.
// DVG_BF16:   D += A_hi*B_hi only.
//
// Kernel structure (320 threads, 1 CTA / SM, persistent over (row tile, N tile) pairs):
//   warp 0     TMA producer: one lane issues the bulk copies of each k-block stage
//   warp 1     TMEM allocator + MMA issuer: one lane issues tcgen05.mma, commits to mbarriers
//   warps 2-9  epilogue: tcgen05.ld the accumulator (lane = row; two warps per TMEM lane quarter split the
//              columns), fused pointwise math with the bias staged in smem and c prefetched, global stores
// Clusters of up to 4 CTAs (consecutive row tiles, same N tile) share each weight k-block via TMA multicast.
// Two TMEM accumulator buffers (2 x 256 columns) let the epilogue of tile i overlap the MMAs of tile i+1.
#include <stdlib.h>

#include "tc_common.cuh"

namespace dvg {

enum { EPI_PACK = 0, EPI_LSTM = 1, EPI_TANH = 2, EPI_GAUSS = 3 };

struct TcArgs {
  int rows, row_tiles;
  const uint8_t* a0; int kb0;
  const uint8_t* a1; int kb1;
  const uint8_t* w;
  const float* bias;
  int n_tile, n_tiles, nsplit, stages, cm;
  // EPI_PACK
  uint8_t* out_packed; int out_kb_total;
  // EPI_LSTM
  const float* c_in; const float* h_in; float* h_out; float* c_out; uint8_t* hp_out; int H;
  const uint8_t* hold; int rows_per_flag;
  // EPI_TANH
  float* y; int ldy; int n_valid;
  // EPI_GAUSS
  const float* eps; float* z; float* mu; float* logvar; int Z;
  unsigned long long* trace;  // DVG_TRACE builds only: per-CTA timestamps
};

// Cluster of CM CTAs along the row-tile axis: the CM CTAs of a cluster work on CM consecutive row tiles
// and the SAME N tile, so the weight k-block images are identical for all of them -- each CTA fetches
// 1/CM of every weight image and TMA-multicasts it into all CM shared memories (L2 -> SM traffic for the
// dominant operand drops by CM).  Activation images are private.  A smem stage is recycled only after the
// MMAs of ALL CM CTAs that read it have retired (multicast tcgen05.commit onto every CTA's empty barrier).
template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(const TcArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B images need 1024-byte alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.kb0 + p.kb1;
  const int CM = p.cm;
  const uint32_t rank = CM > 1 ? ptx::cluster_ctarank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << CM) - 1u);
  const uint32_t b_part = (uint32_t)p.n_tile * 128u;
  const uint32_t stage_bytes = 2u * TC_A_IMG + 2u * b_part;
  const uint32_t bar_base = base + (uint32_t)p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * p.stages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 4);
  float* s_bias = reinterpret_cast<float*>(smem_raw + (bar_base - raw) + 128);  // [2][256] floats

  if (threadIdx.x == 0) {
    TRACE(0);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), (uint32_t)CM);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), EPI_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CM > 1) ptx::cluster_sync_all();   // remote CTAs must see initialised barriers before any multicast
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (threadIdx.x == 0) TRACE(1);

  // cluster-tile schedule: ct -> (row-tile group, N tile); this CTA's row tile = group*CM + rank
  const int groups = (p.row_tiles + CM - 1) / CM;
  const int num_ct = groups * p.n_tiles;
  const int cid = CM > 1 ? (int)ptx::cluster_id_x() : (int)blockIdx.x;
  const int ncl = CM > 1 ? (int)ptx::cluster_count_x() : (int)gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t a_copy = p.nsplit == 1 ? (uint32_t)TC_A_IMG : 2u * TC_A_IMG;
      const uint32_t b_copy = p.nsplit == 1 ? b_part : 2u * b_part;
      const uint32_t slice = b_part / (uint32_t)CM;
      int s = 0;
      uint32_t ph = 0;
      for (int ct = cid; ct < num_ct; ct += ncl) {
        const int nt = ct % p.n_tiles;
        int rt = (ct / p.n_tiles) * CM + (int)rank;
        if (rt >= p.row_tiles) rt = p.row_tiles - 1;   // padding CTA of the last group: loads stay in bounds
        for (int kb = 0; kb < KB; ++kb) {
          ptx::mbar_wait(empty_bar(s), ph ^ 1u);
          const uint8_t* asrc = kb < p.kb0 ? p.a0 + (size_t)(rt * p.kb0 + kb) * (2u * TC_A_IMG)
                                           : p.a1 + (size_t)(rt * p.kb1 + (kb - p.kb0)) * (2u * TC_A_IMG);
          const uint8_t* bsrc = p.w + (size_t)(nt * KB + kb) * (2u * b_part);
          const uint32_t sa = base + (uint32_t)s * stage_bytes;
          const uint32_t sb = sa + 2u * TC_A_IMG;
          ptx::mbar_expect_tx(full_bar(s), a_copy + b_copy);
          ptx::bulk_g2s(sa, asrc, a_copy, full_bar(s));
          if (CM == 1) {
            ptx::bulk_g2s(sb, bsrc, b_copy, full_bar(s));
          } else {
            const uint32_t off = rank * slice;
            ptx::bulk_g2s_mcast(sb + off, bsrc + off, slice, full_bar(s), cmask);
            if (p.nsplit != 1) ptx::bulk_g2s_mcast(sb + b_part + off, bsrc + b_part + off, slice, full_bar(s), cmask);
          }
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(TC_ROWS, p.n_tile);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int ct = cid; ct < num_ct; ct += ncl, ++it) {
        const int acc = it & 1;
        const uint32_t aph = (uint32_t)(it >> 1) & 1u;
        ptx::mbar_wait(tempty_bar(acc), aph ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_STRIDE);
        for (int kb = 0; kb < KB; ++kb) {
          ptx::mbar_wait(full_bar(s), ph);
          if (it < 2) TRACE(2 + it * 12 + kb);     // stage kb of tile `it` landed
          ptx::tc_fence_after();
          const uint32_t sa = base + (uint32_t)s * stage_bytes;
          const uint64_t a_hi = ptx::make_sw128_desc(sa);
          const uint64_t a_lo = ptx::make_sw128_desc(sa + TC_A_IMG);
          const uint64_t b_hi = ptx::make_sw128_desc(sa + 2u * TC_A_IMG);
          const uint64_t b_lo = ptx::make_sw128_desc(sa + 2u * TC_A_IMG + b_part);
#pragma unroll
          for (int k = 0; k < TC_KBLK / 16; ++k) {
            const uint64_t adv = (uint64_t)(k * 2);  // 16 bf16 = 32 bytes = 2 x 16-byte units
            ptx::umma_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0 ? 1u : 0u);
            if (p.nsplit != 1) {
              ptx::umma_bf16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
              ptx::umma_bf16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
            }
          }
          // free the smem stage (in every CTA of the cluster: our multicast slices live there too)
          if (CM == 1) ptx::umma_commit(empty_bar(s));
          else ptx::umma_commit_mcast(empty_bar(s), cmask);
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
        ptx::umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
        if (it < 2) TRACE(2 + it * 12 + 8);        // all MMAs of tile `it` issued
      }
    }
  } else {
    // ===================== epilogue warps (2 .. 2+EPI_WARPS-1) =====================
    const int ew = warp - 2;
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int half = ew >> 2;                      // two warps share a lane quarter and split the columns
    const uint32_t r_in_tile = (uint32_t)(q * 32 + lane);
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const int etid = ew * 32 + lane;               // 0 .. 255
    int it = 0;
    for (int ct = cid; ct < num_ct; ct += ncl, ++it) {
      const int nt = ct % p.n_tiles;
      const int rt = (ct / p.n_tiles) * CM + (int)rank;
      const int acc = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      const int row = rt * TC_ROWS + (int)r_in_tile;
      const bool valid = row < p.rows;
      float* sb = s_bias + acc * 256;
      // stage this tile's bias and prefetch the first c chunk while the MMAs are still running
      if (etid < p.n_tile) {
        float bv = __ldg(p.bias + (size_t)nt * p.n_tile + etid);
        if (EPI == EPI_LSTM) bv *= (etid >> 6) == 2 ? -2.f * kLog2e : -kLog2e;
        sb[etid] = bv;
      }
      float cp[16];
      size_t idx0 = 0;
      bool held = false;
      if (EPI == EPI_LSTM) {
        idx0 = (size_t)row * p.H + nt * 64 + half * 32;
        held = valid && p.hold != nullptr && p.hold[row / p.rows_per_flag] != 0;
        if (valid) load16(p.c_in + idx0, cp);
      }
      ptx::named_bar_sync(1, EPI_WARPS * 32);
      ptx::mbar_wait(tfull_bar(acc), aph);
      if (etid == 0 && it < 2) TRACE(2 + it * 12 + 9);   // accumulator ready
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + tlane + (uint32_t)(acc * ACC_STRIDE);

      if (EPI == EPI_LSTM) {
        // tile columns: [i: 64 units][f: 64][g: 64][o: 64] of hidden units nt*64 .. nt*64+63;
        // this warp handles units half*32 .. half*32+31 in two chunks of 16
        uint8_t* img = p.hp_out + (size_t)(rt * (p.H / 64) + nt) * (2u * TC_A_IMG);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int cb = half * 32 + jj * 16;       // first unit of the chunk within the tile
          float cnext[16];
          if (jj == 0 && valid) load16(p.c_in + idx0 + 16, cnext);
          uint32_t r[64];
          ptx::tmem_ld16x4_wait(tacc + cb, tacc + 64 + cb, tacc + 128 + cb, tacc + 192 + cb, r);
          if (valid) {
            const size_t idx = idx0 + jj * 16;
            float hn[16], cn[16];
            if (held) {
              load16(p.h_in + idx, hn);
#pragma unroll
              for (int i = 0; i < 16; ++i) cn[i] = cp[i];
            } else {
              lstm_pointwise16(r, sb, cb, cp, hn, cn);
            }
            store16(p.c_out + idx, cn);
            store16(p.h_out + idx, hn);
            store_split16(img, r_in_tile, (uint32_t)(cb >> 3), hn);
            if (jj == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) cp[i] = cnext[i];
            }
          }
        }
      } else {
        const int nchunks = p.n_tile / 16;
#pragma unroll 1
        for (int j = half; j < nchunks; j += 2) {
          float v[16];
          ptx::tmem_ld16_wait(tacc + j * 16, v);
          if (valid) {
            const int col0 = nt * p.n_tile + j * 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += sb[j * 16 + i];
            if (EPI == EPI_PACK) {
              uint8_t* img = p.out_packed + (size_t)(rt * p.out_kb_total + (col0 >> 6)) * (2u * TC_A_IMG);
              store_split16(img, r_in_tile, (uint32_t)((col0 & 63) >> 3), v);
            } else if (EPI == EPI_TANH) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (col0 + i < p.n_valid) p.y[(size_t)row * p.ldy + col0 + i] = tanh_f(v[i]);
            } else {  // EPI_GAUSS: columns (2z, 2z+1) = (mu_z, logvar_z)
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                const int zi = (col0 + i) >> 1;
                if (zi < p.Z) {
                  const size_t idx = (size_t)row * p.Z + zi;
                  p.mu[idx] = v[i];
                  p.logvar[idx] = v[i + 1];
                  p.z[idx] = fmaf(p.eps[idx], expf(0.5f * v[i + 1]), v[i]);
                }
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
      if (etid == 0 && it < 2) TRACE(2 + it * 12 + 10);  // epilogue of warp 2 done
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TRACE(30);
  if (CM > 1) ptx::cluster_sync_all();   // no CTA may exit while peers can still multicast into it
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ---------------------------------------------------------------------------------------------------
// cta_group::2 version: a CTA PAIR owns a 256-row x n_tile tile.  Each CTA stages its own 128 activation
// rows and HALF of the weight image (n_tile/2 columns); the leader's tcgen05.mma.cta_group::2 drives both
// tensor cores, each accumulating its 128 rows x n_tile columns in its own TMEM.  Per CTA this halves the
// weight bytes read from shared memory per MMA (96 -> 64 B/clk) and the stage size (96 -> 64 KB, so three
// stages fit): the 1-CTA kernel above is shared-memory-bandwidth bound (operand reads + TMA fill > 128 B/clk).
//   full[s]   (per CTA)  own TMA bytes landed
//   pfull[s]  (leader)   peer's stage landed  -- relayed by the peer's otherwise idle MMA warp
//   empty[s]  (per CTA)  multicast tcgen05.commit of the leader: stage free in both CTAs
//   tfull[a]  (per CTA)  multicast commit: accumulator a complete in both TMEMs
//   tempty[a] (leader)   all epilogue warps of BOTH CTAs drained accumulator a
// ---------------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm2_kernel(const TcArgs p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = p.kb0 + p.kb1;
  constexpr int CM = 2;
  const uint32_t rank = ptx::cluster_ctarank();
  const uint32_t b_half = (uint32_t)p.n_tile * 64u;          // bytes of this CTA's half of one weight image part
  const uint32_t b_part = (uint32_t)p.n_tile * 128u;         // bytes of a full weight image part in HBM
  const uint32_t stage_bytes = 2u * TC_A_IMG + 2u * b_half;
  const uint32_t bar_base = base + (uint32_t)p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto pfull_bar = [&](int s) { return bar_base + 8u * (2 * p.stages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (3 * p.stages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (3 * p.stages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * p.stages + 4);
  float* s_bias = reinterpret_cast<float*>(smem_raw + (bar_base - raw) + 128);  // [2][256] floats

  if (threadIdx.x == 0) {
    TRACE(0);
    for (int s = 0; s < p.stages; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
      ptx::mbar_init(pfull_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);
      ptx::mbar_init(tempty_bar(a), 2 * EPI_WARPS);
    }
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
  if (threadIdx.x == 0) TRACE(1);

  const int groups = (p.row_tiles + CM - 1) / CM;
  const int num_ct = groups * p.n_tiles;
  const int cid = (int)ptx::cluster_id_x();
  const int ncl = (int)ptx::cluster_count_x();

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      const uint32_t a_copy = p.nsplit == 1 ? (uint32_t)TC_A_IMG : 2u * TC_A_IMG;
      const uint32_t nparts = p.nsplit == 1 ? 1u : 2u;
      int s = 0;
      uint32_t ph = 0;
      for (int ct = cid; ct < num_ct; ct += ncl) {
        const int nt = ct % p.n_tiles;
        int rt = (ct / p.n_tiles) * CM + (int)rank;
        if (rt >= p.row_tiles) rt = p.row_tiles - 1;
        for (int kb = 0; kb < KB; ++kb) {
          ptx::mbar_wait(empty_bar(s), ph ^ 1u);
          const uint8_t* asrc = kb < p.kb0 ? p.a0 + (size_t)(rt * p.kb0 + kb) * (2u * TC_A_IMG)
                                           : p.a1 + (size_t)(rt * p.kb1 + (kb - p.kb0)) * (2u * TC_A_IMG);
          const uint8_t* bsrc = p.w + (size_t)(nt * KB + kb) * (2u * b_part) + rank * b_half;
          const uint32_t sa = base + (uint32_t)s * stage_bytes;
          const uint32_t sb = sa + 2u * TC_A_IMG;
          ptx::mbar_expect_tx(full_bar(s), a_copy + nparts * b_half);
          ptx::bulk_g2s(sa, asrc, a_copy, full_bar(s));
          ptx::bulk_g2s(sb, bsrc, b_half, full_bar(s));
          if (nparts == 2) ptx::bulk_g2s(sb + b_half, bsrc + b_part, b_half, full_bar(s));
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      if (rank == 0) {
        // ===================== MMA issuer (leader CTA) =====================
        const uint32_t idesc = ptx::make_idesc_bf16(2 * TC_ROWS, p.n_tile);
        int s = 0;
        uint32_t ph = 0;
        int it = 0;
        for (int ct = cid; ct < num_ct; ct += ncl, ++it) {
          const int acc = it & 1;
          const uint32_t aph = (uint32_t)(it >> 1) & 1u;
          ptx::mbar_wait(tempty_bar(acc), aph ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * ACC_STRIDE);
          for (int kb = 0; kb < KB; ++kb) {
            ptx::mbar_wait(full_bar(s), ph);
            ptx::mbar_wait(pfull_bar(s), ph);
            if (it < 2) TRACE(2 + it * 12 + kb);
            ptx::tc_fence_after();
            const uint32_t sa = base + (uint32_t)s * stage_bytes;
            const uint64_t a_hi = ptx::make_sw128_desc(sa);
            const uint64_t a_lo = ptx::make_sw128_desc(sa + TC_A_IMG);
            const uint64_t b_hi = ptx::make_sw128_desc(sa + 2u * TC_A_IMG);
            const uint64_t b_lo = ptx::make_sw128_desc(sa + 2u * TC_A_IMG + b_half);
#pragma unroll
            for (int k = 0; k < TC_KBLK / 16; ++k) {
              const uint64_t adv = (uint64_t)(k * 2);
              ptx::umma2_bf16(d_tmem, a_hi + adv, b_hi + adv, idesc, (kb | k) != 0 ? 1u : 0u);
              if (p.nsplit != 1) {
                ptx::umma2_bf16(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                ptx::umma2_bf16(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
              }
            }
            ptx::umma2_commit_mcast(empty_bar(s), 3);
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
          ptx::umma2_commit_mcast(tfull_bar(acc), 3);
          if (it < 2) TRACE(2 + it * 12 + 8);
        }
      } else {
        // ===================== relay (peer CTA): "my stage landed" -> leader's pfull =====================
        int s = 0;
        uint32_t ph = 0;
        for (int ct = cid; ct < num_ct; ct += ncl) {
          for (int kb = 0; kb < KB; ++kb) {
            ptx::mbar_wait(full_bar(s), ph);
            ptx::mbar_arrive_remote(pfull_bar(s), 0);
            if (++s == p.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps (2 .. 2+EPI_WARPS-1) =====================
    const int ew = warp - 2;
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int half = ew >> 2;                      // two warps share a lane quarter and split the columns
    const uint32_t r_in_tile = (uint32_t)(q * 32 + lane);
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const int etid = ew * 32 + lane;               // 0 .. 255
    int it = 0;
    for (int ct = cid; ct < num_ct; ct += ncl, ++it) {
      const int nt = ct % p.n_tiles;
      const int rt = (ct / p.n_tiles) * CM + (int)rank;
      const int acc = it & 1;
      const uint32_t aph = (uint32_t)(it >> 1) & 1u;
      const int row = rt * TC_ROWS + (int)r_in_tile;
      const bool valid = row < p.rows;
      float* sb = s_bias + acc * 256;
      // stage this tile's bias and prefetch the first c chunk while the MMAs are still running
      if (etid < p.n_tile) {
        float bv = __ldg(p.bias + (size_t)nt * p.n_tile + etid);
        if (EPI == EPI_LSTM) bv *= (etid >> 6) == 2 ? -2.f * kLog2e : -kLog2e;
        sb[etid] = bv;
      }
      float cp[16];
      size_t idx0 = 0;
      bool held = false;
      if (EPI == EPI_LSTM) {
        idx0 = (size_t)row * p.H + nt * 64 + half * 32;
        held = valid && p.hold != nullptr && p.hold[row / p.rows_per_flag] != 0;
        if (valid) load16(p.c_in + idx0, cp);
      }
      ptx::named_bar_sync(1, EPI_WARPS * 32);
      ptx::mbar_wait(tfull_bar(acc), aph);
      if (etid == 0 && it < 2) TRACE(2 + it * 12 + 9);   // accumulator ready
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + tlane + (uint32_t)(acc * ACC_STRIDE);

      if (EPI == EPI_LSTM) {
        // tile columns: [i: 64 units][f: 64][g: 64][o: 64] of hidden units nt*64 .. nt*64+63;
        // this warp handles units half*32 .. half*32+31 in two chunks of 16
        uint8_t* img = p.hp_out + (size_t)(rt * (p.H / 64) + nt) * (2u * TC_A_IMG);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int cb = half * 32 + jj * 16;       // first unit of the chunk within the tile
          float cnext[16];
          if (jj == 0 && valid) load16(p.c_in + idx0 + 16, cnext);
          uint32_t r[64];
          ptx::tmem_ld16x4_wait(tacc + cb, tacc + 64 + cb, tacc + 128 + cb, tacc + 192 + cb, r);
          if (valid) {
            const size_t idx = idx0 + jj * 16;
            float hn[16], cn[16];
            if (held) {
              load16(p.h_in + idx, hn);
#pragma unroll
              for (int i = 0; i < 16; ++i) cn[i] = cp[i];
            } else {
              lstm_pointwise16(r, sb, cb, cp, hn, cn);
            }
            store16(p.c_out + idx, cn);
            store16(p.h_out + idx, hn);
            store_split16(img, r_in_tile, (uint32_t)(cb >> 3), hn);
            if (jj == 0) {
#pragma unroll
              for (int i = 0; i < 16; ++i) cp[i] = cnext[i];
            }
          }
        }
      } else {
        const int nchunks = p.n_tile / 16;
#pragma unroll 1
        for (int j = half; j < nchunks; j += 2) {
          float v[16];
          ptx::tmem_ld16_wait(tacc + j * 16, v);
          if (valid) {
            const int col0 = nt * p.n_tile + j * 16;
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += sb[j * 16 + i];
            if (EPI == EPI_PACK) {
              uint8_t* img = p.out_packed + (size_t)(rt * p.out_kb_total + (col0 >> 6)) * (2u * TC_A_IMG);
              store_split16(img, r_in_tile, (uint32_t)((col0 & 63) >> 3), v);
            } else if (EPI == EPI_TANH) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (col0 + i < p.n_valid) p.y[(size_t)row * p.ldy + col0 + i] = tanh_f(v[i]);
            } else {  // EPI_GAUSS: columns (2z, 2z+1) = (mu_z, logvar_z)
#pragma unroll
              for (int i = 0; i < 16; i += 2) {
                const int zi = (col0 + i) >> 1;
                if (zi < p.Z) {
                  const size_t idx = (size_t)row * p.Z + zi;
                  p.mu[idx] = v[i];
                  p.logvar[idx] = v[i + 1];
                  p.z[idx] = fmaf(p.eps[idx], expf(0.5f * v[i + 1]), v[i]);
                }
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) ptx::mbar_arrive(tempty_bar(acc));
        else ptx::mbar_arrive_remote(tempty_bar(acc), 0);
      }
      if (etid == 0 && it < 2) TRACE(2 + it * 12 + 10);  // epilogue of warp 2 done
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TRACE(30);
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc2(tmem_base, TMEM_COLS);
  }
}


// W_x = W_ih0 W_e  ([4H, G]),  b_x = W_ih0 b_e + b_ih0 + b_hh0  -- the embed Linear folded into layer 0 (fp64 accumulate).
__global__ void fold_embed_kernel(const float* __restrict__ w_ih0, const float* __restrict__ embed_w,
                                  const float* __restrict__ embed_b, const float* __restrict__ b_ih0,
                                  const float* __restrict__ b_hh0, int H, int G, float* __restrict__ wx,
                                  float* __restrict__ bx) {
  const int r = blockIdx.x;   // gate row 0..4H-1
  for (int g = threadIdx.x; g <= G; g += blockDim.x) {
    double acc = 0.0;
    if (g < G) {
      for (int k = 0; k < H; ++k) acc += (double)w_ih0[(size_t)r * H + k] * (double)embed_w[(size_t)k * G + g];
      wx[(size_t)r * G + g] = (float)acc;
    } else {
      for (int k = 0; k < H; ++k) acc += (double)w_ih0[(size_t)r * H + k] * (double)embed_b[k];
      bx[r] = (float)(acc + (double)b_ih0[r] + (double)b_hh0[r]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Packing kernels
// ---------------------------------------------------------------------------------------------------

// fp32 rows [rows][K] (ld) -> A images [RT][kbs][2][16 KB].  grid (RT, kbs), 256 threads.
__global__ void __launch_bounds__(256) tc_pack_rows_kernel(uint8_t* dst, const float* src, int ld, int rows, int K,
                                                           int kbs) {
  const int rt = blockIdx.x, kb = blockIdx.y;
  uint8_t* img = dst + (size_t)(rt * kbs + kb) * (2u * TC_A_IMG);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qd = threadIdx.x + i * 256;
    const int r = qd >> 3, chunk = qd & 7;
    const int row = rt * TC_ROWS + r;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kb * 64 + chunk * 8 + e;
      v[e] = (row < rows && k < K) ? __ldg(src + (size_t)row * ld + k) : 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v[2 * e], h0, l0);
      split_bf16(v[2 * e + 1], h1, l1);
      hi[e] = pack2_bf16(h0, h1);
      lo[e] = pack2_bf16(l0, l1);
    }
    const uint32_t o = sw128_offset((uint32_t)r, (uint32_t)chunk);
    *reinterpret_cast<uint4*>(img + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(img + TC_A_IMG + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// Weight images [n_tiles][KB][2][n_tile*128 B] from up to two fp32 matrices concatenated along K.
//   mode 0: column c = nt*n_tile + n  <- row c of w0/w1           (c < n_rows, else zero)
//   mode 1: gates (n_tile == 256): n = g*64 + u <- row g*H + nt*64 + u
//   mode 2: gaussian heads: column c <- row c>>1 of (c even ? w0 : w1), no K concatenation
__global__ void __launch_bounds__(256) tc_pack_weights_kernel(uint8_t* dst, const float* w0, int K0, int kb0,
                                                              const float* w1, int K1, int kb1, int n_tile,
                                                              int n_rows, int mode, int H) {
  const int nt = blockIdx.x, kb = blockIdx.y;
  const int KB = kb0 + kb1;
  const uint32_t b_part = (uint32_t)n_tile * 128u;
  uint8_t* img = dst + (size_t)(nt * KB + kb) * (2u * b_part);
  for (int qd = threadIdx.x; qd < n_tile * 8; qd += blockDim.x) {
    const int n = qd >> 3, chunk = qd & 7;
    const int c = nt * n_tile + n;
    int row;
    const float* w;
    int K, kloc;
    if (mode == 2) {
      row = c >> 1;
      w = (c & 1) ? w1 : w0;
      K = K0;
      kloc = kb * 64;
    } else {
      row = mode == 1 ? (n >> 6) * H + nt * 64 + (n & 63) : c;
      if (kb < kb0) { w = w0; K = K0; kloc = kb * 64; }
      else { w = w1; K = K1; kloc = (kb - kb0) * 64; }
    }
    const bool row_ok = row < n_rows;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kloc + chunk * 8 + e;
      v[e] = (row_ok && k < K) ? w[(size_t)row * K + k] : 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v[2 * e], h0, l0);
      split_bf16(v[2 * e + 1], h1, l1);
      hi[e] = pack2_bf16(h0, h1);
      lo[e] = pack2_bf16(l0, l1);
    }
    const uint32_t o = sw128_offset((uint32_t)n, (uint32_t)chunk);
    *reinterpret_cast<uint4*>(img + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(img + b_part + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__global__ void tc_pack_bias_kernel(float* dst, const float* b0, const float* b1, int n_total, int n_tile, int n_rows,
                                    int mode, int H) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_total) return;
  const int nt = c / n_tile, n = c % n_tile;
  float v = 0.f;
  if (mode == 2) {
    const int row = c >> 1;
    if (row < n_rows) v = (c & 1) ? b1[row] : b0[row];
  } else {
    const int row = mode == 1 ? (n >> 6) * H + nt * 64 + (n & 63) : c;
    if (row < n_rows) v = b0[row] + (b1 ? b1[row] : 0.f);
  }
  dst[c] = v;
}

// ---------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------
static int plan_alloc(TcGemmPlan& pl, int n_tile, int n_tiles, int kb0, int kb1) {
  pl.n_tile = n_tile; pl.n_tiles = n_tiles; pl.kb0 = kb0; pl.kb1 = kb1;
  if (!pl.w) {
    DVG_CUDA(cudaMalloc(&pl.w, (size_t)n_tiles * (kb0 + kb1) * 2 * n_tile * 128));
    DVG_CUDA(cudaMalloc(&pl.bias, sizeof(float) * n_tiles * n_tile));
  }
  return DVG_OK;
}

static int plan_pack(TcGemmPlan& pl, const float* w0, int K0, const float* w1, int K1, const float* b0,
                     const float* b1, int n_rows, int mode, int H, cudaStream_t s) {
  dim3 grid(pl.n_tiles, pl.kb0 + pl.kb1);
  tc_pack_weights_kernel<<<grid, 256, 0, s>>>(pl.w, w0, K0, pl.kb0, w1, K1, pl.kb1, pl.n_tile, n_rows, mode, H);
  DVG_LAUNCH_CHECK();
  const int n_total = pl.n_tiles * pl.n_tile;
  tc_pack_bias_kernel<<<ceil_div(n_total, 128), 128, 0, s>>>(pl.bias, b0, b1, n_total, pl.n_tile, n_rows, mode, H);
  DVG_LAUNCH_CHECK();
  return DVG_OK;
}

int lstm_tc_pack(dvg_lstm_s* h, const float* embed_w, const float* embed_b, const float* const* w_ih,
                 const float* const* w_hh, const float* const* b_ih, const float* const* b_hh,
                 const float* head0_w, const float* head0_b, const float* head1_w, const float* head1_b,
                 cudaStream_t stream) {
  const int G = h->dims.input_size, H = h->dims.hidden_size, L = h->dims.n_layers;
  const bool gauss = h->dims.kind == DVG_GAUSSIAN_LSTM;
  const int hk = H / 64;
  int rc;
  // embed: N = H split into tiles of 256 / 128 / 64 columns
  const int en = (H % 256 == 0) ? 256 : (H % 128 == 0 ? 128 : 64);
  if ((rc = plan_alloc(h->tc_embed, en, H / en, ceil_div(G, 64), 0))) return rc;
  if ((rc = plan_pack(h->tc_embed, embed_w, G, nullptr, 0, embed_b, nullptr, H, 0, H, stream))) return rc;
  for (int l = 0; l < L; ++l) {
    if ((rc = plan_alloc(h->tc_layer[l], 256, hk, hk, hk))) return rc;
    if ((rc = plan_pack(h->tc_layer[l], w_ih[l], H, w_hh[l], H, b_ih[l], b_hh[l], 4 * H, 1, H, stream))) return rc;
  }
  // layer 0 with the embed folded in (fused step kernel): K = [x (G padded to 64) | h_0]
  {
    if (!h->fold_wx) {
      DVG_CUDA(cudaMalloc(&h->fold_wx, sizeof(float) * 4 * H * G));
      DVG_CUDA(cudaMalloc(&h->fold_bx, sizeof(float) * 4 * H));
    }
    fold_embed_kernel<<<4 * H, 128, 0, stream>>>(w_ih[0], embed_w, embed_b, b_ih[0], b_hh[0], H, G, h->fold_wx, h->fold_bx);
    DVG_LAUNCH_CHECK();
    if ((rc = plan_alloc(h->tc_layer0f, 256, hk, ceil_div(G, 64), hk))) return rc;
    if ((rc = plan_pack(h->tc_layer0f, h->fold_wx, G, w_hh[0], H, h->fold_bx, nullptr, 4 * H, 1, H, stream))) return rc;
  }
  const int n_head = gauss ? 2 * h->dims.output_size : h->dims.output_size;
  const int hn = (int)align_up(n_head, 32);
  DVG_REQUIRE(hn <= 256, "tensor-core head supports at most 256 output columns (got %d)", n_head);
  if ((rc = plan_alloc(h->tc_head, hn, 1, hk, 0))) return rc;
  if (gauss) {
    if ((rc = plan_pack(h->tc_head, head0_w, H, head1_w, H, head0_b, head1_b, h->dims.output_size, 2, H, stream)))
      return rc;
  } else {
    if ((rc = plan_pack(h->tc_head, head0_w, H, nullptr, 0, head0_b, nullptr, n_head, 0, H, stream))) return rc;
  }
  return lstm_small_pack(h, stream);       // weight images of the small-batch cluster kernel (lstm_small.cu)
}

void lstm_tc_free(dvg_lstm_s* h) {
  auto fr = [](TcGemmPlan& p) {
    if (p.w) cudaFree(p.w);
    if (p.bias) cudaFree(p.bias);
    p.w = nullptr; p.bias = nullptr;
  };
  lstm_small_free(h);
  fr(h->tc_embed);
  fr(h->tc_head);
  fr(h->tc_layer0f);
  if (h->fold_wx) cudaFree(h->fold_wx);
  if (h->fold_bx) cudaFree(h->fold_bx);
  if (h->fused_flags) cudaFree(h->fused_flags);
  if (h->sched_dev) cudaFree(h->sched_dev);
  h->fold_wx = h->fold_bx = nullptr;
  h->fused_flags = nullptr;
  h->sched_dev = nullptr;
  for (int l = 0; l < MAX_LAYERS; ++l) fr(h->tc_layer[l]);
}

size_t lstm_tc_packed_state_bytes(const dvg_lstm_s* h, int rows) {
  if (!h->tc_ok) return 0;
  return (size_t)h->dims.n_layers * ceil_div(rows, TC_ROWS) * (h->dims.hidden_size / 64) * 2 * TC_A_IMG;
}
size_t lstm_tc_scratch_bytes_xp(const dvg_lstm_s* h, int rows) {
  return (size_t)ceil_div(rows, TC_ROWS) * ceil_div(h->dims.input_size, 64) * 2 * TC_A_IMG;
}
size_t lstm_tc_scratch_bytes_ep(const dvg_lstm_s* h, int rows) {
  return (size_t)ceil_div(rows, TC_ROWS) * (h->dims.hidden_size / 64) * 2 * TC_A_IMG;
}

int lstm_tc_repack_state(dvg_lstm_s* h, int rows, const float* h_f32, uint8_t* hp, cudaStream_t stream) {
  const int H = h->dims.hidden_size, L = h->dims.n_layers, hk = H / 64, RT = ceil_div(rows, TC_ROWS);
  for (int l = 0; l < L; ++l) {
    tc_pack_rows_kernel<<<dim3(RT, hk), 256, 0, stream>>>(hp + (size_t)l * RT * hk * 2 * TC_A_IMG,
                                                          h_f32 + (size_t)l * rows * H, H, rows, H, hk);
    DVG_LAUNCH_CHECK();
  }
  return DVG_OK;
}

static bool use_pairs() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DVG_TC_PAIRS");     // developer switch: 0 forces the 1-CTA (multicast) kernel
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

template <int EPI>
static int launch_tc(const dvg_lstm_s* h, TcArgs& a, cudaStream_t stream) {
  const bool pairs = use_pairs() && a.row_tiles >= 2 && a.n_tile % 32 == 0;
  const size_t b_stage = pairs ? (size_t)a.n_tile * 128 : 2 * (size_t)a.n_tile * 128;
  const size_t stage_bytes = 2 * (size_t)TC_A_IMG + b_stage;
  const size_t tail = 128 /*barriers + tmem slot*/ + 2 * 256 * sizeof(float) /*bias*/;
  const size_t budget = 227 * 1024 - 1024 /*alignment slack*/ - tail;
  int stages = (int)(budget / stage_bytes);
  if (stages > 4) stages = 4;
  DVG_REQUIRE(stages >= 2, "tile too large for a 2-stage pipeline");
  a.stages = stages;
  const size_t smem = stages * stage_bytes + 1024 + tail;
  static bool configured = false;  // per template instantiation
  if (!configured) {
    DVG_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    DVG_CUDA(cudaFuncSetAttribute(tc_gemm2_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  int cm;
  if (pairs) {
    cm = 2;
  } else {
    // 1-CTA kernel: cluster along the row-tile axis multicasts the weight images
    cm = a.row_tiles >= 4 ? 4 : (a.row_tiles >= 2 ? 2 : 1);
    while (cm > 1 && (a.n_tile % cm) != 0) cm >>= 1;
  }
  a.cm = cm;
  const int groups = ceil_div(a.row_tiles, cm);
  int clusters = groups * a.n_tiles;
  const int max_clusters = h->sm_count / cm;
  if (clusters > max_clusters) clusters = max_clusters;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * cm);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cm;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cm > 1 ? 1 : 0;
#ifdef DVG_TRACE
  static unsigned long long* tbuf = nullptr;
  const bool tr = EPI == EPI_LSTM && getenv("DVG_TC_TRACE") != nullptr;
  if (tr) {
    if (!tbuf) cudaMalloc(&tbuf, 256 * TRACE_SLOTS * 8);
    cudaMemsetAsync(tbuf, 0, 256 * TRACE_SLOTS * 8, stream);
    a.trace = tbuf;
  }
#endif
  if (pairs) DVG_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm2_kernel<EPI>, (const TcArgs)a));
  else DVG_CUDA(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<EPI>, (const TcArgs)a));
#ifdef DVG_TRACE
  if (tr) {
    static int n_dump = 0;
    cudaStreamSynchronize(stream);
    if (n_dump++ == 6) {
      std::vector<unsigned long long> hbuf(256 * TRACE_SLOTS);
      cudaMemcpy(hbuf.data(), tbuf, 256 * TRACE_SLOTS * 8, cudaMemcpyDeviceToHost);
      unsigned long long t0 = ~0ull;
      for (int b = 0; b < (int)cfg.gridDim.x; ++b) if (hbuf[b * TRACE_SLOTS] && hbuf[b * TRACE_SLOTS] < t0) t0 = hbuf[b * TRACE_SLOTS];
      fprintf(stderr, "TRACE grid=%d cm=%d pairs=%d stages=%d (ns since first CTA start)\n", (int)cfg.gridDim.x, cm,
              (int)pairs, stages);
      for (int b = 0; b < (int)cfg.gridDim.x; b += (b < 8 ? 1 : 13)) {
        fprintf(stderr, "cta %3d:", b);
        for (int i = 0; i < 31; ++i) {
          unsigned long long v = hbuf[b * TRACE_SLOTS + i];
          if (i == 2 || i == 14 || i == 26) fprintf(stderr, " |");
          fprintf(stderr, " %lld", v ? (long long)(v - t0) : -1ll);
        }
        fprintf(stderr, "\n");
      }
    }
  }
#endif
  return DVG_OK;
}


int lstm_tc_step(dvg_lstm_s* h, int nsplit, int rows, const float* x, int ldx, const float* h_in,
                 const float* c_in, const uint8_t* hp_in, float* h_out, float* c_out, uint8_t* hp_out, float* y,
                 int ldy, const float* eps, float* z, float* mu, float* logvar, const uint8_t* hold,
                 int rows_per_flag, cudaStream_t stream) {
  const int G = h->dims.input_size, H = h->dims.hidden_size, L = h->dims.n_layers;
  const int hk = H / 64, RT = ceil_div(rows, TC_ROWS), kbx = ceil_div(G, 64);
  const size_t lsz = (size_t)rows * H;
  const size_t lpk = (size_t)RT * hk * 2 * TC_A_IMG;
  int rc;
  if (lstm_step_usable(h, rows))
    return lstm_step_launch(h, nullptr, nsplit, rows, x, ldx, h_in, c_in, hp_in, h_out, c_out, hp_out, y, ldy, eps, z,
                            mu, logvar, hold, rows_per_flag, stream, nullptr);
  if (lstm_small_usable(h, rows))
    return lstm_small_launch(h, nsplit, rows, x, ldx, h_in, c_in, hp_in, h_out, c_out, hp_out, y, ldy, hold, rows_per_flag,
                             stream);
  h->prof_mark(stream);
  tc_pack_rows_kernel<<<dim3(RT, kbx), 256, 0, stream>>>(h->tc_xp, x, ldx, rows, G, kbx);
  DVG_LAUNCH_CHECK();
  h->prof_mark(stream);
  {
    TcArgs a{};
    a.rows = rows; a.row_tiles = RT; a.nsplit = nsplit;
    a.a0 = h->tc_xp; a.kb0 = kbx; a.a1 = nullptr; a.kb1 = 0;
    a.w = h->tc_embed.w; a.bias = h->tc_embed.bias; a.n_tile = h->tc_embed.n_tile; a.n_tiles = h->tc_embed.n_tiles;
    a.out_packed = h->tc_ep; a.out_kb_total = hk;
    if ((rc = launch_tc<EPI_PACK>(h, a, stream))) return rc;
    h->prof_mark(stream);
  }
  const uint8_t* layer_in = h->tc_ep;
  for (int l = 0; l < L; ++l) {
    TcArgs a{};
    a.rows = rows; a.row_tiles = RT; a.nsplit = nsplit;
    a.a0 = layer_in; a.kb0 = hk; a.a1 = hp_in + l * lpk; a.kb1 = hk;
    a.w = h->tc_layer[l].w; a.bias = h->tc_layer[l].bias; a.n_tile = 256; a.n_tiles = hk;
    a.c_in = c_in + l * lsz; a.h_in = h_in + l * lsz; a.h_out = h_out + l * lsz; a.c_out = c_out + l * lsz;
    a.hp_out = hp_out + l * lpk; a.H = H; a.hold = hold; a.rows_per_flag = rows_per_flag > 0 ? rows_per_flag : 1;
    if ((rc = launch_tc<EPI_LSTM>(h, a, stream))) return rc;
    h->prof_mark(stream);
    layer_in = hp_out + l * lpk;
  }
  {
    TcArgs a{};
    a.rows = rows; a.row_tiles = RT; a.nsplit = nsplit;
    a.a0 = layer_in; a.kb0 = hk; a.a1 = nullptr; a.kb1 = 0;
    a.w = h->tc_head.w; a.bias = h->tc_head.bias; a.n_tile = h->tc_head.n_tile; a.n_tiles = 1;
    if (h->dims.kind == DVG_GAUSSIAN_LSTM) {
      a.eps = eps; a.z = z; a.mu = mu; a.logvar = logvar; a.Z = h->dims.output_size;
      if ((rc = launch_tc<EPI_GAUSS>(h, a, stream))) return rc;
    } else {
      a.y = y; a.ldy = ldy; a.n_valid = h->dims.output_size;
      if ((rc = launch_tc<EPI_TANH>(h, a, stream))) return rc;
    }
  }
  h->prof_mark(stream);
  return DVG_OK;
}

}  // namespace dvg
