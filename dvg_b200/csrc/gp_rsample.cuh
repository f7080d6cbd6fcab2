// GP posterior sample of one (rollout, latent dim) problem by 256 cooperating threads (gpytorch's .rsample():
// generate_frames.py:171,292; train.py:284).  Shared by the stand-alone rsample kernels (gp.cu) and the persistent
// step kernel (lstm_step.cu), which resamples fired rollouts at the end of the same launch.  `sync` is the barrier of
// the 256 participating threads, `smf` their shared-memory scratch (gp_rsample_smem_floats() floats).
#pragma once
#include "common.cuh"

namespace dvg {

__host__ __device__ inline size_t gp_rsample_smem_floats(int N, int mp) {
  return (size_t)2 * mp * mp + 3 * (size_t)N * (mp + 4) + (size_t)N * (N + 1) + 4 * (size_t)N + 2 * (size_t)mp;
}

// ---------------------------------------------------------------------------------------------------
// rsample: one CTA per (rollout s, latent dim d); full [N,N] predictive covariance in shared memory,
// Cholesky, y = mean + L eps.
// ---------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;

// One (rollout, dim) problem per call, NTHR cooperating threads (256 in the stand-alone kernels, 512 in the step kernel): full [N,N] predictive covariance in
// shared memory, Cholesky, y = mean + L eps.  Every global operand (factors, z, beta, eps, x) is staged into
// shared memory first with batched coalesced loads: the first version read beta / eps / z from global inside
// dependent inner loops and spent ~45 of its 70 us waiting on L2.
template <int NTHR, class Sync>
__device__ __forceinline__ void gp_rsample_body(float* smf, const int tid, Sync sync, int s_idx, int d, int N, int D,
                                                int mp, const float* __restrict__ x, int ldx,
                                                const float* __restrict__ eps,
                                                const float* __restrict__ zall, const float* __restrict__ linv_all,
                                                const float* __restrict__ lqt_all,
                                                const float* __restrict__ alpha_all, const float* __restrict__ hyp,
                                                float* __restrict__ out, int ldo) {
#ifdef DVG_TRACE
  long long tq[8]; int nq = 0;
#define RSQ() do { sync(); if (nq < 8) tq[nq++] = clock64(); } while (0)
#else
#define RSQ() do {} while (0)
#endif
  RSQ();
  const int MP = mp;
  // rows of K / U / R are read as float4: stride MP + 4 floats keeps them 16-byte aligned (MP % 4 == 0) and, for
  // MP / 4 even (MP = 40), makes the per-thread-row float4 reads bank-conflict free
  const int ldk = MP + 4, lds = N + 1;
  float* s_linv = smf;                 // [MP][MP]
  float* s_lqt = s_linv + MP * MP;     // [MP][MP]
  float* s_k = s_lqt + MP * MP;        // [N][MP+4]  K_xz
  float* s_u = s_k + N * ldk;          // [N][MP+4]  (Linv k_n)
  float* s_r = s_u + N * ldk;          // [N][MP+4]  (L_q^T k_n)
  float* s_sig = s_r + N * ldk;        // [N][N+1]   Sigma_y -> L
  float* s_x = s_sig + N * lds;        // [N]
  float* s_mean = s_x + N;             // [N]
  float* s_t = s_mean + N;             // [N]
  float* s_eps = s_t + N;              // [N]
  float* s_z = s_eps + N;              // [MP]
  float* s_beta = s_z + MP;            // [MP]
  // All global operands are requested before the first shared-memory store, so the problem pays ONE L2 round trip
  // (~1 k cycles) instead of three dependent ones (factors, then z / beta / x / eps, then the hyper-parameters).
  const float ell = __ldg(hyp + d * 4 + 0), sc = __ldg(hyp + d * 4 + 1), c = __ldg(hyp + d * 4 + 2),
              noise = __ldg(hyp + d * 4 + 3);
  {
    // N, MP <= 128 <= NTHR: one element per thread
    float zr = 0.f, betar = 0.f, xr = 0.f, er = 0.f;
    if (tid < MP) {
      zr = __ldg(zall + (size_t)d * MP + tid);
      betar = __ldg(alpha_all + (size_t)d * MP + tid);
    }
    if (tid < N) {
      xr = __ldg(x + (size_t)(s_idx * N + tid) * ldx + d);
      er = __ldg(eps + ((size_t)s_idx * D + d) * N + tid);
    }
    const float4* g1 = reinterpret_cast<const float4*>(linv_all + (size_t)d * MP * MP);
    const float4* g2 = reinterpret_cast<const float4*>(lqt_all + (size_t)d * MP * MP);
    const int n4 = MP * MP / 4;
    for (int e0 = 0; e0 < n4; e0 += NTHR * 2) {       // 4 independent 16-byte loads in flight per thread
      float4 t[2][2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int e = e0 + u * NTHR + tid;
        if (e < n4) { t[u][0] = __ldg(g1 + e); t[u][1] = __ldg(g2 + e); }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int e = e0 + u * NTHR + tid;
        if (e < n4) {
          reinterpret_cast<float4*>(s_linv)[e] = t[u][0];
          reinterpret_cast<float4*>(s_lqt)[e] = t[u][1];
        }
      }
    }
    if (tid < MP) { s_z[tid] = zr; s_beta[tid] = betar; }
    if (tid < N) { s_x[tid] = xr; s_eps[tid] = er; }
  }
  const float inv_ell = 1.0f / ell;
  sync();
  RSQ();
  for (int e = tid; e < N * MP; e += NTHR) {
    const int n = e / MP, m = e % MP;
    const float t = (s_x[n] - s_z[m]) * inv_ell;
    s_k[n * ldk + m] = sc * expf(-0.5f * t * t);
  }
  sync();
  RSQ();
  // Thread mapping for the O(N M^2) / O(N^3) phases: the point index (n or a) is the FAST thread index (row strides
  // ldk / lds -> conflict-free), the matrix row (j or b) is shared by a whole warp (broadcast reads), and there are
  // no runtime integer divisions.  All dot products read float4.  (History, cycles per problem on B200: first
  // version 140 k; scalar LDS + triangle skipping by `continue` 87 k; float4 dots + flat pair list 70 k.)
  const int lg = N <= 32 ? 5 : (N <= 64 ? 6 : 7);       // N <= 128 (shared-memory bound)
  const int fi = tid & ((1 << lg) - 1);                 // point index
  const int grp = tid >> lg, ngrp = NTHR >> lg;   // which matrix rows this thread visits
  const int MB = MP >> 2;
  if (fi < N) {
    const float4* kr4 = reinterpret_cast<const float4*>(s_k + fi * ldk);
    // two matrix rows (j, j + ngrp) at a time: twice the independent FMA chains per thread
    for (int j = grp; j < MP; j += 2 * ngrp) {
      const int j2 = j + ngrp < MP ? j + ngrp : j;      // (last odd row: computed twice, stored once)
      const float4* lr4 = reinterpret_cast<const float4*>(s_linv + j * MP);
      const float4* qr4 = reinterpret_cast<const float4*>(s_lqt + j * MP);
      const float4* lr4b = reinterpret_cast<const float4*>(s_linv + j2 * MP);
      const float4* qr4b = reinterpret_cast<const float4*>(s_lqt + j2 * MP);
      float v0 = 0.f, v1 = 0.f, w0 = 0.f, w1 = 0.f, x0 = 0.f, x1 = 0.f, y0 = 0.f, y1 = 0.f;
      const int jb = j >> 2, jb2 = j2 >> 2;
      for (int mb = 0; mb < MB; ++mb) {
        const float4 k = kr4[mb];
        if (mb <= jb) {                                 // Linv row j: entries m <= j (exact zeros beyond j)
          const float4 l = lr4[mb];
          v0 = fmaf(l.x, k.x, v0); v1 = fmaf(l.y, k.y, v1); v0 = fmaf(l.z, k.z, v0); v1 = fmaf(l.w, k.w, v1);
        }
        if (mb >= jb) {                                 // L_q^T row j: entries m >= j (exact zeros below j)
          const float4 q = qr4[mb];
          w0 = fmaf(q.x, k.x, w0); w1 = fmaf(q.y, k.y, w1); w0 = fmaf(q.z, k.z, w0); w1 = fmaf(q.w, k.w, w1);
        }
        if (mb <= jb2) {
          const float4 l = lr4b[mb];
          x0 = fmaf(l.x, k.x, x0); x1 = fmaf(l.y, k.y, x1); x0 = fmaf(l.z, k.z, x0); x1 = fmaf(l.w, k.w, x1);
        }
        if (mb >= jb2) {
          const float4 q = qr4b[mb];
          y0 = fmaf(q.x, k.x, y0); y1 = fmaf(q.y, k.y, y1); y0 = fmaf(q.z, k.z, y0); y1 = fmaf(q.w, k.w, y1);
        }
      }
      s_u[fi * ldk + j] = v0 + v1;
      s_r[fi * ldk + j] = w0 + w1;
      if (j2 != j) {
        s_u[fi * ldk + j2] = x0 + x1;
        s_r[fi * ldk + j2] = y0 + y1;
      }
    }
  }
  sync();
  RSQ();
  for (int n = tid; n < N; n += NTHR) {
    float mu0 = 0.f, mu1 = 0.f;
    for (int j = 0; j + 1 < MP; j += 2) {
      mu0 = fmaf(s_beta[j], s_u[n * ldk + j], mu0);
      mu1 = fmaf(s_beta[j + 1], s_u[n * ldk + j + 1], mu1);
    }
    s_mean[n] = c + (mu0 + mu1);       // MP is a multiple of 4
  }
  // Sigma_y, lower triangle: the N (N + 1) / 2 pairs (a, b <= a) are dealt out flat (b fastest: one row a is a
  // broadcast, consecutive rows b are conflict free), so every thread gets the same number of pairs.
  const int npairs = N * (N + 1) / 2;
  auto pair_of = [](int pidx, int& a, int& b) {
    a = (int)((sqrtf(8.f * (float)pidx + 1.f) - 1.f) * 0.5f);
    while (a * (a + 1) / 2 > pidx) --a;
    while ((a + 1) * (a + 2) / 2 <= pidx) ++a;
    b = pidx - a * (a + 1) / 2;
  };
  auto sigma_y = [&](int a, int b) -> float {
    const float4* ra4 = reinterpret_cast<const float4*>(s_r + a * ldk);
    const float4* rb4 = reinterpret_cast<const float4*>(s_r + b * ldk);
    const float4* ua4 = reinterpret_cast<const float4*>(s_u + a * ldk);
    const float4* ub4 = reinterpret_cast<const float4*>(s_u + b * ldk);
    float rr0 = 0.f, rr1 = 0.f, uu0 = 0.f, uu1 = 0.f;
#pragma unroll 2
    for (int mb = 0; mb < MB; ++mb) {
      const float4 r1 = ra4[mb], r2 = rb4[mb], u1 = ua4[mb], u2 = ub4[mb];
      rr0 = fmaf(r1.x, r2.x, rr0); uu0 = fmaf(u1.x, u2.x, uu0);
      rr1 = fmaf(r1.y, r2.y, rr1); uu1 = fmaf(u1.y, u2.y, uu1);
      rr0 = fmaf(r1.z, r2.z, rr0); uu0 = fmaf(u1.z, u2.z, uu0);
      rr1 = fmaf(r1.w, r2.w, rr1); uu1 = fmaf(u1.w, u2.w, uu1);
    }
    const float t = (s_x[a] - s_x[b]) * inv_ell;
    const float kxx = a == b ? sc : sc * expf(-0.5f * t * t);
    return (rr0 + rr1) + (kxx - (uu0 + uu1)) + (a == b ? noise : 0.f);
  };
  // Cholesky as LDL^T, right-looking with UNSCALED columns: after step j the trailing block holds
  // S - sum_{k<=j} c_k c_k^T / d_k with c_k the unscaled column k and d_k its diagonal; L[i][j] = c_j[i] / sqrt(d_j)
  // is applied in one pass at the end.
  constexpr int SLOTS = 2048 / NTHR;                     // matrix elements a thread can own in registers
  if (npairs <= SLOTS * NTHR) {
    // Register-resident variant (N <= 63): every thread OWNS its <= SLOTS elements of the triangle for the whole
    // factorisation; per column the owners of column j publish it to a double-buffered N-float vector, one barrier,
    // and every owner of a trailing element applies x -= (c[a] / d) c[b] in registers: no read-modify-write of the
    // trailing block through shared memory.  Same operations in the same order per element, so the result is
    // bit-identical to the shared-memory variant below.
    float xv[SLOTS];
    int ar[SLOTS], br[SLOTS];
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) {
      const int pidx = tid + k * NTHR;
      xv[k] = 0.f; ar[k] = 0; br[k] = -1;                // empty slot: never a column owner, never trailing
      if (pidx < npairs) {
        pair_of(pidx, ar[k], br[k]);
        xv[k] = sigma_y(ar[k], br[k]);
      }
    }
    RSQ();
    // (Measured, cycles for N = 50 on 16 warps: shared-memory variant ~18 k, this one 14.5 k = ~300 per column, the
    // barrier round trip dominating; a two-columns-per-barrier version that re-derives the second column locally
    // was slower, 24 k: two dependent reciprocals and seven LDS per element on the chain; dealing the elements
    // out in column-major order so that finished slots can be skipped by block-uniform branches measured no better
    // -- whole rollout 2.07 ms against 2.03 ms.)
    float* colbuf = s_k;                                 // K_xz is dead after the U / R phase: 2 x N floats of it
    for (int j = 0; j + 1 < N; ++j) {
      float* col = colbuf + (j & 1) * N;
#pragma unroll
      for (int k = 0; k < SLOTS; ++k)
        if (br[k] == j) col[ar[k]] = xv[k];
      sync();
      const float inv_d = __fdividef(1.0f, col[j]);
#pragma unroll
      for (int k = 0; k < SLOTS; ++k)
        if (br[k] > j) {
          const float ca = col[ar[k]] * inv_d;
          xv[k] = fmaf(-ca, col[br[k]], xv[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < SLOTS; ++k)
      if (ar[k] == br[k]) s_t[ar[k]] = xv[k];            // d_j
    sync();
#pragma unroll
    for (int k = 0; k < SLOTS; ++k)
      if (br[k] >= 0) s_sig[ar[k] * lds + br[k]] = ar[k] == br[k] ? sqrtf(xv[k]) : xv[k] * rsqrtf(s_t[br[k]]);
    sync();
  } else {
    for (int pidx = tid; pidx < npairs; pidx += NTHR) {
      int a, b;
      pair_of(pidx, a, b);
      s_sig[a * lds + b] = sigma_y(a, b);
    }
    sync();
    RSQ();
    // Shared-memory variant (one barrier per column).  ~41 k cycles for N = 50 with 256 threads; two alternatives
    // were measured and were no faster -- panels of 4 columns (two barriers per panel, 44.6 k: the read-modify-write
    // of the trailing block through shared memory serialises on LDS -> FMA -> STS latency either way) and a
    // left-looking one-thread-per-row variant (loads only, 64-thread barrier, 47.5 k: two lone warps issue at
    // ~0.25 IPC).
    for (int j = 0; j + 1 < N; ++j) {
      const float inv_d = __fdividef(1.0f, s_sig[j * lds + j]);
      const int a = j + 1 + fi;                           // this thread's ROW (consecutive threads -> stride lds, odd:
      if (a < N) {                                        // conflict-free; columns b are warp-uniform -> broadcast)
        const float ca = s_sig[a * lds + j] * inv_d;
        // batches of 4 with all loads issued before the first store: a plain read-modify-write loop serialises on
        // LDS -> FMA -> STS because the compiler must assume the store aliases the next loads
        for (int b0 = j + 1 + grp; b0 <= a; b0 += 4 * ngrp) {
          float xv[4], yv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int b = b0 + u * ngrp;
            xv[u] = b <= a ? s_sig[a * lds + b] : 0.f;
            yv[u] = b <= a ? s_sig[b * lds + j] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int b = b0 + u * ngrp;
            if (b <= a) s_sig[a * lds + b] = fmaf(-ca, yv[u], xv[u]);
          }
        }
      }
      sync();
    }
    for (int j = grp; j < N; j += ngrp) {
      const float rs = rsqrtf(s_sig[j * lds + j]);
      if (fi > j && fi < N) s_sig[fi * lds + j] *= rs;
    }
    sync();
    for (int j = tid; j < N; j += NTHR) s_sig[j * lds + j] = sqrtf(s_sig[j * lds + j]);
    sync();
  }
  RSQ();
  for (int n = tid; n < N; n += NTHR) {
    float a0 = 0.f, a1 = 0.f;
    int k = 0;
    for (; k + 1 <= n; k += 2) {
      a0 = fmaf(s_sig[n * lds + k], s_eps[k], a0);
      a1 = fmaf(s_sig[n * lds + k + 1], s_eps[k + 1], a1);
    }
    if (k <= n) a0 = fmaf(s_sig[n * lds + k], s_eps[k], a0);
    out[(size_t)(s_idx * N + n) * ldo + d] = s_mean[n] + (a0 + a1);
  }
  RSQ();
#ifdef DVG_TRACE
  if (tid == 0 && d == 0) printf("rsample phases (cycles): stage %lld K %lld UR %lld meanSig %lld chol %lld Leps %lld\n", tq[1]-tq[0], tq[2]-tq[1], tq[3]-tq[2], tq[4]-tq[3], tq[5]-tq[4], tq[6]-tq[5]);
#endif
#undef RSQ
}


}  // namespace dvg
