// Handle structures and cross-file launch entry points (internal; not part of the C ABI).
#pragma once
#include <vector>

#include "common.cuh"

namespace dvg {

constexpr int MAX_LAYERS = 8;

// ---- tensor-core GEMM tile plan ------------------------------------------------------------------
struct TcGemmPlan {
  int n_tile = 0;    // UMMA N (multiple of 16, <= 256)
  int n_tiles = 0;   // number of N tiles
  int kb0 = 0;       // k-blocks (of 64) from A source 0
  int kb1 = 0;       // k-blocks from A source 1
  uint8_t* w = nullptr;   // packed weights [n_tiles][kb0+kb1][2][n_tile*128 B]
  float* bias = nullptr;  // [n_tiles*n_tile] in packed column order
};

}  // namespace dvg

struct dvg_lstm_s {
  dvg_lstm_dims dims{};
  int device = 0;
  int sm_count = 0;
  int cc_major = 0;
  bool tc_ok = false;  // tensor-core variants usable (sm_100 and H % 64 == 0)

  // --- DVG_FP32 (FFMA) weight copies: W^T, K-major rows, padded N ---------------------------------
  int n_head = 0;      // head columns: G_out (lstm) or 2Z interleaved mu/logvar (gaussian)
  int hp = 0;          // H rounded up to 64
  int np_head = 0;     // n_head rounded up to 64
  float* f_embed_wt = nullptr;  // [G_in][hp]
  float* f_embed_b = nullptr;   // [hp]
  float* f_layer_wt[dvg::MAX_LAYERS] = {};  // [2H][4H], column = unit*4 + gate
  float* f_layer_b[dvg::MAX_LAYERS] = {};   // [4H] (b_ih + b_hh), same column order
  float* f_head_wt = nullptr;   // [H][np_head]
  float* f_head_b = nullptr;    // [np_head]

  // --- tensor-core packed weights -----------------------------------------------------------------
  dvg::TcGemmPlan tc_embed, tc_layer[dvg::MAX_LAYERS], tc_head;
  dvg::TcGemmPlan tc_layer0f;    // layer 0 with the embed Linear folded in (fused step kernel)
  uint8_t* small_w[2] = {nullptr, nullptr};   // lstm_small.cu: per-(32-unit tile, k-block) contiguous 128-row A images of layers 0 / 1
  float* fold_wx = nullptr;      // [4H][G]  W_ih0 W_e
  float* fold_bx = nullptr;      // [4H]     W_ih0 b_e + b_ih0 + b_hh0
  int* fused_flags = nullptr;    // dependency counters of the persistent step kernel (lstm_step.cu), self-resetting
  size_t flag_set_words = 0;     // three counter sets of this many words (chained launches rotate through them)
  size_t xp_stride = 0;          // two packed-x slabs, this many bytes apart
  // chained step launches (dvg_lstm_chain_begin / _end, lstm_step.cu): consecutive launches overlap
  bool chain_on = false, chain_ok = false, chain_dirty = false;
  int chain_idx = 0, chain_rows = 0, chain_nsplit = 0, chain_prev_trig = 0, chain_prev_restore = 0;
  const void* chain_out = nullptr;         // h_out of the last chained launch (must be the next one's h_in)
  const void* chain_gp = nullptr;
  cudaStream_t chain_stream = nullptr;
  int* sched_dev = nullptr;      // optional item order of the step kernel for (sched_rows, sched_pairs)
  int sched_len = 0, sched_rows = 0, sched_pairs = 0;

  // --- scratch, grown by reserve() ----------------------------------------------------------------
  int reserved_rows = 0;
  float* scratch_e = nullptr;    // fp32 [rows][H]: embed output (FFMA variant)
  uint8_t* tc_xp = nullptr;      // packed x            [RT][kbx][2][16 KB]
  uint8_t* tc_ep = nullptr;      // packed embed output [RT][H/64][2][16 KB]
  float* rs_buf = nullptr;       // [rows][G_out] side buffer of the step kernel's in-launch GP resample
  // Scratch outgrown by a later reserve(): CUDA graphs captured before the growth have the old device pointers baked
  // in, so the old allocations stay alive (and self-consistent: own flags, own packed-x slab) until destroy.
  std::vector<void*> retired;

  // --- optional per-kernel timing (dvg_lstm_profile): events recorded between launches ------------
  bool prof_on = false;
  int prof_n = 0;
  cudaEvent_t prof_ev[16] = {};
  void prof_mark(cudaStream_t s) {
    if (prof_on && prof_n < 16) cudaEventRecord(prof_ev[prof_n++], s);
  }
};

struct dvg_gp_s {
  dvg_gp_dims dims{};
  int device = 0;
  int mp = 0;                 // M rounded up to 4
  float* z = nullptr;         // [D][mp] inducing points
  float* linv = nullptr;      // [D][mp][mp] L_ZZ^-1 (lower triangular, zero padded)
  float* lqt = nullptr;       // [D][mp][mp] L_q^T  (upper triangular: lqt[j][i] = L_q[i][j], i >= j)
  float* linvT = nullptr;     // [D][mp][mp] transpose of linv  (lane-per-row trigger kernel reads [m][j])
  float* lq = nullptr;        // [D][mp][mp] masked L_q, row-major
  float* alpha = nullptr;     // [D][mp] beta = L_ZZ^-1 (m_q - c)
  float* hyp = nullptr;       // [D][4]  ell, s, c, noise
  double* work = nullptr;     // fp64 scratch for prepare [D][3][M][M]
  float* var_rows = nullptr;  // scratch [max_rollouts][D] for the trigger
  int var_rows_cap = 0;
  unsigned int* ticket = nullptr;   // [1 + max groups] last-CTA tickets of the fused trigger kernel (self-resetting)
  int* trig_list = nullptr;         // [max_rollouts] compacted rollouts that fired in the last trigger call
  int* trig_count = nullptr;
  // large inducing sets (gp_big.cu): factors loaded pre-computed, mp = M rounded up to 64, tiled FP32 GEMM kernels
  bool big = false;
  float* partial = nullptr;   // [D][mp/64][n_pad][3] row-block partial sums
  size_t partial_cap = 0;
  std::vector<void*> retired; // outgrown scratch kept alive for already captured graphs (freed by destroy)
  // tensor-core path of the large-M predictive (gp_tc.cu): bf16 hi/lo k-block images of Linv / L_q^T
  uint8_t* tc_img_v = nullptr; uint8_t* tc_img_w = nullptr;
  size_t tc_img_bytes = 0;
  int tc_JT = 0, tc_MT = 0;
  float* tc_alpha2 = nullptr;   // [D][mp] Linv^T beta
  size_t tc_alpha2_n = 0;
};

namespace dvg {

// lstm_fp32.cu
int lstm_fp32_pack(dvg_lstm_s* h, const float* embed_w, const float* embed_b, const float* const* w_ih,
                   const float* const* w_hh, const float* const* b_ih, const float* const* b_hh,
                   const float* head0_w, const float* head0_b, const float* head1_w, const float* head1_b,
                   cudaStream_t stream);
int lstm_fp32_step(dvg_lstm_s* h, int rows, const float* x, int ldx, const float* h_in, const float* c_in,
                   float* h_out, float* c_out, float* y, int ldy, const float* eps, float* z, float* mu,
                   float* logvar, const uint8_t* hold, int rows_per_flag, cudaStream_t stream);

// lstm_tc.cu
int lstm_tc_pack(dvg_lstm_s* h, const float* embed_w, const float* embed_b, const float* const* w_ih,
                 const float* const* w_hh, const float* const* b_ih, const float* const* b_hh,
                 const float* head0_w, const float* head0_b, const float* head1_w, const float* head1_b,
                 cudaStream_t stream);
size_t lstm_tc_packed_state_bytes(const dvg_lstm_s* h, int rows);
int lstm_tc_repack_state(dvg_lstm_s* h, int rows, const float* h_f32, uint8_t* hp, cudaStream_t stream);
int lstm_tc_step(dvg_lstm_s* h, int nsplit, int rows, const float* x, int ldx, const float* h_in,
                 const float* c_in, const uint8_t* hp_in, float* h_out, float* c_out, uint8_t* hp_out, float* y,
                 int ldy, const float* eps, float* z, float* mu, float* logvar, const uint8_t* hold,
                 int rows_per_flag, cudaStream_t stream);
void lstm_tc_free(dvg_lstm_s* h);
size_t lstm_tc_scratch_bytes_xp(const dvg_lstm_s* h, int rows);
size_t lstm_tc_scratch_bytes_ep(const dvg_lstm_s* h, int rows);

// lstm_step.cu: the persistent whole-step kernel
struct StepTrigHost {
  int S, W, warmup;
  float factor;
  const int32_t* stat_rows;
  float* window; int32_t* count; float* value; float* thr; uint8_t* mask;
  const float* rs_eps;   // non-null: fired rollouts are resampled inside the launch
};
// lstm_small.cu: one 16-CTA cluster per step for <= 64 rows (gate columns split over the cluster, DSMEM exchange)
bool lstm_small_usable(const dvg_lstm_s* h, int rows);
int lstm_small_pack(dvg_lstm_s* h, cudaStream_t stream);
void lstm_small_free(dvg_lstm_s* h);
bool lstm_small_can_fuse_trigger(const dvg_lstm_s* h, const dvg_gp_s* g, int rows, int S);
int lstm_small_launch(dvg_lstm_s* h, int nsplit, int rows, const float* x, int ldx, const float* h_in, const float* c_in,
                      const uint8_t* hp_in, float* h_out, float* c_out, uint8_t* hp_out, float* y, int ldy,
                      const uint8_t* hold, int rows_per_flag, cudaStream_t stream, dvg_gp_s* g = nullptr,
                      const StepTrigHost* trig = nullptr);

bool lstm_step_usable(const dvg_lstm_s* h, int rows);
size_t lstm_step_flag_words(const dvg_lstm_s* h, int rows);
int lstm_step_chain(dvg_lstm_s* h, bool begin, cudaStream_t stream);
// gp_factor.cu: blocked fp64 Cholesky + triangular inverse of K_ZZ + jitter I (large inducing sets, once per weight load)
size_t gp_factorize_workspace(int M, int batch);
int gp_factorize(int D, int M, double jitter, const float* inducing, const float* var_mean, const float* mean_const,
                 const float* raw_os, const float* raw_ls, float* linv, float* beta, void* workspace,
                 size_t workspace_bytes, cudaStream_t stream);
size_t lstm_step_xp_bytes(const dvg_lstm_s* h, int rows);
int lstm_step_build_schedule(dvg_lstm_s* h, int rows);
int lstm_step_launch(dvg_lstm_s* h, dvg_gp_s* g, int nsplit, int rows, const float* x, int ldx, const float* h_in,
                     const float* c_in, const uint8_t* hp_in, float* h_out, float* c_out, uint8_t* hp_out, float* y,
                     int ldy, const float* eps, float* z, float* mu, float* logvar, const uint8_t* hold,
                     int rows_per_flag, cudaStream_t stream, const StepTrigHost* trig);
bool lstm_tc_can_fuse_trigger(const dvg_lstm_s* h, const dvg_gp_s* g, int rows);
int lstm_tc_rollout_step(dvg_lstm_s* h, dvg_gp_s* g, int nsplit, int rows, const float* x, int ldx, const float* h_in,
                         const float* c_in, const uint8_t* hp_in, float* h_out, float* c_out, uint8_t* hp_out,
                         float* y, int ldy, int S, const int32_t* stat_rows, float* window, int W, int32_t* count,
                         int warmup, float factor, float* value, float* thr, uint8_t* mask, const float* rs_eps,
                         cudaStream_t stream);
bool lstm_tc_can_fuse_rsample(const dvg_gp_s* g, int n_points);

// gp.cu
int gp_prepare_launch(dvg_gp_s* h, const float* inducing, const float* var_mean, const float* chol_var,
                      const float* mean_const, const float* raw_os, const float* raw_ls, const float* raw_noise,
                      cudaStream_t stream);
int gp_predict_launch(dvg_gp_s* h, int n_rows, const float* x, int ldx, const int32_t* row_index, float* mean, int ldm,
                      float* var, int ldv, cudaStream_t stream);
int gp_trigger_launch(dvg_gp_s* h, int S, const float* x, int ldx, const int32_t* stat_rows, float* window, int W,
                      int32_t* count, int warmup, float factor, float* value, float* thr, uint8_t* mask,
                      cudaStream_t stream);
int gp_rsample_launch(dvg_gp_s* h, int S, int N, const float* x, int ldx, const float* eps, const uint8_t* mask,
                      float* out, int ldo, cudaStream_t stream);

// gp_tc.cu: tcgen05 version of the large-M predictive partial sums
bool gp_tc_enabled();
int gp_tc_pack(dvg_gp_s* h, cudaStream_t stream);
int gp_tc_partial_launch(dvg_gp_s* h, int n_rows, int n_pad, const float* x, int ldx, const int32_t* row_index,
                         int want_mean, cudaStream_t stream);

// gp_big.cu
int gp_big_load_factors(dvg_gp_s* h, const float* inducing, const float* linv, const float* lq, const float* beta,
                        const float* hyp, cudaStream_t stream);
int gp_big_predict_launch(dvg_gp_s* h, int n_rows, const float* x, int ldx, const int32_t* row_index, float* mean,
                          long long mean_sn, long long mean_sd, float* var, long long var_sn, long long var_sd,
                          cudaStream_t stream);
int gp_big_rsample_launch(dvg_gp_s* h, int S, int N, const float* x, int ldx, const float* eps, const uint8_t* mask,
                          float* out, int ldo, cudaStream_t stream);
int gp_big_trigger_launch(dvg_gp_s* h, int S, const float* x, int ldx, const int32_t* stat_rows, float* window, int W,
                          int32_t* count, int warmup, float factor, float* value, float* thr, uint8_t* mask,
                          cudaStream_t stream);

// rollout.cu
int eval_seq_finn_launch(int T, int S, int B, int C, int H, int W, const float* gt, const float* gen, float* ssim,
                         float* psnr, cudaStream_t stream);
int eval_seq_skimage_launch(int T, int S, int B, int C, int H, int W, const float* gt, const float* gen, float* ssim,
                            float* psnr, cudaStream_t stream);
int rollout_score_launch(int T, int S, int B, int G, const float* out, const float* target, float* scores,
                         cudaStream_t stream);

// moving_mnist.cu
int moving_mnist_launch(int B, int T, int W, int n_digits, int deterministic, const float* bank, int n_bank,
                        const uint32_t* draws, int draws_per_seq, int32_t* traj, float* frames, cudaStream_t stream);

}  // namespace dvg
