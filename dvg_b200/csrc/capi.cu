// C ABI of libdvg_b200.so (see include/dvg_b200.h).  Argument validation, handle lifetime, scratch
// management; all compute is in lstm_fp32.cu / lstm_tc.cu / gp.cu.
#include <stdarg.h>
#include <string.h>

#include <new>

#include "internal.cuh"

namespace dvg {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static bool capturing(cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &st) != cudaSuccess) { cudaGetLastError(); return false; }
  return st != cudaStreamCaptureStatusNone;
}
}  // namespace dvg

using namespace dvg;

extern "C" {

const char* dvg_last_error(void) { return g_err; }
int dvg_version(void) { return 100; }

int dvg_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  DVG_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  DVG_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return DVG_OK;
}

// ---------------------------------------------------------------------------------------------------
// LSTM
// ---------------------------------------------------------------------------------------------------
static void lstm_free_all(dvg_lstm_s* h) {
  auto fr = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
  fr(h->f_embed_wt); fr(h->f_embed_b); fr(h->f_head_wt); fr(h->f_head_b);
  for (int l = 0; l < MAX_LAYERS; ++l) { fr(h->f_layer_wt[l]); fr(h->f_layer_b[l]); }
  fr(h->scratch_e); fr(h->tc_xp); fr(h->tc_ep); fr(h->rs_buf);
  for (void* q : h->retired) cudaFree(q);
  h->retired.clear();
  for (int i = 0; i < 16; ++i)
    if (h->prof_ev[i]) { cudaEventDestroy(h->prof_ev[i]); h->prof_ev[i] = nullptr; }
  lstm_tc_free(h);
}

static int lstm_pack_all(dvg_lstm_s* h, const float* embed_w, const float* embed_b, const float* const* w_ih,
                         const float* const* w_hh, const float* const* b_ih, const float* const* b_hh,
                         const float* head0_w, const float* head0_b, const float* head1_w, const float* head1_b,
                         cudaStream_t stream) {
  DVG_REQUIRE(embed_w && embed_b && w_ih && w_hh && b_ih && b_hh && head0_w && head0_b, "null weight pointer");
  if (h->dims.kind == DVG_GAUSSIAN_LSTM) DVG_REQUIRE(head1_w && head1_b, "gaussian_lstm needs logvar_net weights");
  int rc = lstm_fp32_pack(h, embed_w, embed_b, w_ih, w_hh, b_ih, b_hh, head0_w, head0_b, head1_w, head1_b, stream);
  if (rc) return rc;
  if (h->tc_ok)
    rc = lstm_tc_pack(h, embed_w, embed_b, w_ih, w_hh, b_ih, b_hh, head0_w, head0_b, head1_w, head1_b, stream);
  return rc;
}

int dvg_lstm_prepare(dvg_lstm_t* out, const dvg_lstm_dims* dims, const float* embed_w, const float* embed_b,
                     const float* const* w_ih, const float* const* w_hh, const float* const* b_ih,
                     const float* const* b_hh, const float* head0_w, const float* head0_b, const float* head1_w,
                     const float* head1_b, dvg_stream_t stream) {
  DVG_REQUIRE(out && dims, "null argument");
  DVG_REQUIRE(dims->kind == DVG_LSTM || dims->kind == DVG_GAUSSIAN_LSTM, "bad kind %d", dims->kind);
  DVG_REQUIRE(dims->input_size > 0 && dims->output_size > 0 && dims->hidden_size > 0, "bad sizes");
  DVG_REQUIRE(dims->hidden_size % 16 == 0, "hidden_size must be a multiple of 16 (got %d)", dims->hidden_size);
  DVG_REQUIRE(dims->n_layers >= 1 && dims->n_layers <= MAX_LAYERS, "n_layers must be in [1,%d]", MAX_LAYERS);
  dvg_lstm_s* h = new (std::nothrow) dvg_lstm_s();
  DVG_REQUIRE(h, "out of host memory");
  h->dims = *dims;
  cudaDeviceProp prop;
  if (cudaGetDevice(&h->device) != cudaSuccess || cudaGetDeviceProperties(&prop, h->device) != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    delete h;
    return DVG_ERR_CUDA;
  }
  h->sm_count = prop.multiProcessorCount;
  h->cc_major = prop.major;
  h->tc_ok = prop.major == 10 && dims->hidden_size % 64 == 0;
  int rc = lstm_pack_all(h, embed_w, embed_b, w_ih, w_hh, b_ih, b_hh, head0_w, head0_b, head1_w, head1_b,
                         (cudaStream_t)stream);
  if (rc) { lstm_free_all(h); delete h; return rc; }
  *out = h;
  return DVG_OK;
}

int dvg_lstm_refresh(dvg_lstm_t h, const float* embed_w, const float* embed_b, const float* const* w_ih,
                     const float* const* w_hh, const float* const* b_ih, const float* const* b_hh,
                     const float* head0_w, const float* head0_b, const float* head1_w, const float* head1_b,
                     dvg_stream_t stream) {
  DVG_REQUIRE(h, "null handle");
  return lstm_pack_all(h, embed_w, embed_b, w_ih, w_hh, b_ih, b_hh, head0_w, head0_b, head1_w, head1_b,
                       (cudaStream_t)stream);
}

int dvg_lstm_destroy(dvg_lstm_t h) {
  if (!h) return DVG_OK;
  lstm_free_all(h);
  delete h;
  return DVG_OK;
}

int dvg_lstm_reserve(dvg_lstm_t h, int rows) {
  DVG_REQUIRE(h && rows > 0, "bad argument");
  if (rows <= h->reserved_rows) return DVG_OK;
  DVG_CUDA(cudaDeviceSynchronize());
  // Never free scratch that an already captured CUDA graph may reference: retire it until destroy.
  auto fr = [&](auto*& p) { if (p) h->retired.push_back((void*)p); p = nullptr; };
  fr(h->scratch_e); fr(h->tc_xp); fr(h->tc_ep); fr(h->rs_buf); fr(h->fused_flags); fr(h->sched_dev);
  h->sched_len = h->sched_rows = h->sched_pairs = 0;
  h->reserved_rows = 0;
  DVG_CUDA(cudaMalloc(&h->scratch_e, sizeof(float) * (size_t)rows * h->dims.hidden_size));
  if (h->tc_ok) {
    const size_t xpb = lstm_step_xp_bytes(h, rows) > lstm_tc_scratch_bytes_xp(h, rows) ? lstm_step_xp_bytes(h, rows)
                                                                                        : lstm_tc_scratch_bytes_xp(h, rows);
    h->xp_stride = align_up(xpb, 1024);            // two slabs: chained step launches alternate
    DVG_CUDA(cudaMalloc(&h->tc_xp, 2 * h->xp_stride));
    DVG_CUDA(cudaMemset(h->tc_xp, 0, 2 * h->xp_stride));
    DVG_CUDA(cudaMalloc(&h->rs_buf, sizeof(float) * (size_t)rows * h->dims.output_size));
    DVG_CUDA(cudaMalloc(&h->tc_ep, lstm_tc_scratch_bytes_ep(h, rows)));
    DVG_CUDA(cudaMemset(h->tc_ep, 0, lstm_tc_scratch_bytes_ep(h, rows)));
    h->flag_set_words = lstm_step_flag_words(h, rows);      // three sets: chained step launches rotate through them
    DVG_CUDA(cudaMalloc(&h->fused_flags, sizeof(int) * (3 * h->flag_set_words + 8)));       // + the chain's retired counter
    DVG_CUDA(cudaMemset(h->fused_flags, 0, sizeof(int) * (3 * h->flag_set_words + 8)));
    h->chain_on = h->chain_ok = false;
    h->chain_idx = 0;
    int rc = lstm_step_build_schedule(h, rows);
    if (rc) return rc;
  }
  h->reserved_rows = rows;
  return DVG_OK;
}

int dvg_lstm_chain_begin(dvg_lstm_t h, dvg_stream_t stream) {
  DVG_REQUIRE(h, "null handle");
  return lstm_step_chain(h, true, (cudaStream_t)stream);
}
int dvg_lstm_chain_end(dvg_lstm_t h, dvg_stream_t stream) {
  DVG_REQUIRE(h, "null handle");
  return lstm_step_chain(h, false, (cudaStream_t)stream);
}

static size_t state_f32_bytes(const dvg_lstm_s* h, int rows) {
  return (size_t)2 * h->dims.n_layers * rows * h->dims.hidden_size * sizeof(float);
}
size_t dvg_lstm_state_packed_offset(dvg_lstm_t h, int rows) {
  if (!h || rows <= 0) return 0;
  return align_up(state_f32_bytes(h, rows), 1024);
}
size_t dvg_lstm_state_bytes(dvg_lstm_t h, int rows) {
  if (!h || rows <= 0) return 0;
  return dvg_lstm_state_packed_offset(h, rows) + lstm_tc_packed_state_bytes(h, rows);
}

int dvg_lstm_state_repack(dvg_lstm_t h, int rows, void* state, dvg_stream_t stream) {
  DVG_REQUIRE(h && state && rows > 0, "bad argument");
  if (!h->tc_ok) return DVG_OK;
  return lstm_tc_repack_state(h, rows, (const float*)state,
                              (uint8_t*)state + dvg_lstm_state_packed_offset(h, rows), (cudaStream_t)stream);
}

static int lstm_step_common(dvg_lstm_t h, int variant, int rows, const float* x, int ldx, const void* state_in,
                            void* state_out, float* y, int ldy, const float* eps, float* z, float* mu, float* logvar,
                            const uint8_t* hold, int rows_per_flag, cudaStream_t stream) {
  DVG_REQUIRE(h && x && state_in && state_out, "null argument");
  DVG_REQUIRE(rows > 0 && ldx >= h->dims.input_size, "bad rows/ldx");
  DVG_REQUIRE(state_in != state_out, "state_in and state_out must be distinct blocks");
  DVG_REQUIRE(variant == DVG_FP32 || variant == DVG_BF16X3 || variant == DVG_BF16, "bad variant %d", variant);
  if (hold) DVG_REQUIRE(rows_per_flag > 0, "rows_per_flag must be positive when hold is given");
  if (variant != DVG_FP32 && !h->tc_ok) {
    if (h->cc_major != 10) {
      set_error("tensor-core variants need an sm_100 device (found sm_%d0)", h->cc_major);
      return DVG_ERR_ARCH;
    }
    set_error("tensor-core variants need hidden_size %% 64 == 0 (got %d)", h->dims.hidden_size);
    return DVG_ERR_ARG;
  }
  if (rows > h->reserved_rows) {
    if (capturing(stream)) {
      set_error("dvg_lstm_reserve(%d) must be called before stream capture", rows);
      return DVG_ERR_STATE;
    }
    int rc = dvg_lstm_reserve(h, rows);
    if (rc) return rc;
  }
  const size_t lsz = (size_t)h->dims.n_layers * rows * h->dims.hidden_size;
  const float* h_in = (const float*)state_in;
  const float* c_in = h_in + lsz;
  float* h_out = (float*)state_out;
  float* c_out = h_out + lsz;
  if (variant == DVG_FP32)
    return lstm_fp32_step(h, rows, x, ldx, h_in, c_in, h_out, c_out, y, ldy, eps, z, mu, logvar, hold, rows_per_flag,
                          stream);
  const size_t poff = dvg_lstm_state_packed_offset(h, rows);
  return lstm_tc_step(h, variant == DVG_BF16 ? 1 : 3, rows, x, ldx, h_in, c_in, (const uint8_t*)state_in + poff, h_out,
                      c_out, (uint8_t*)state_out + poff, y, ldy, eps, z, mu, logvar, hold, rows_per_flag, stream);
}

int dvg_lstm_step(dvg_lstm_t h, int variant, int rows, const float* x, int ldx, const void* state_in, void* state_out,
                  float* y, int ldy, const uint8_t* hold, int rows_per_flag, dvg_stream_t stream) {
  DVG_REQUIRE(h && h->dims.kind == DVG_LSTM, "handle is not an lstm");
  DVG_REQUIRE(y && ldy >= h->dims.output_size, "bad y/ldy");
  return lstm_step_common(h, variant, rows, x, ldx, state_in, state_out, y, ldy, nullptr, nullptr, nullptr, nullptr,
                          hold, rows_per_flag, (cudaStream_t)stream);
}

int dvg_lstm_profile(dvg_lstm_t h, int variant, int rows, const float* x, int ldx, const void* state_in,
                     void* state_out, float* y, int ldy, float* kernel_ms, int max_slots, dvg_stream_t stream) {
  DVG_REQUIRE(h && h->dims.kind == DVG_LSTM && kernel_ms && max_slots > 0, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  DVG_REQUIRE(!capturing(s), "dvg_lstm_profile cannot run inside stream capture");
  for (int i = 0; i < 16; ++i)
    if (!h->prof_ev[i]) DVG_CUDA(cudaEventCreate(&h->prof_ev[i]));
  h->prof_on = true;
  h->prof_n = 0;
  int rc = lstm_step_common(h, variant, rows, x, ldx, state_in, state_out, y, ldy, nullptr, nullptr, nullptr, nullptr,
                            nullptr, 0, s);
  h->prof_on = false;
  if (rc) return rc;
  DVG_CUDA(cudaStreamSynchronize(s));
  for (int i = 0; i < max_slots; ++i) kernel_ms[i] = 0.f;
  for (int i = 0; i + 1 < h->prof_n && i < max_slots; ++i)
    DVG_CUDA(cudaEventElapsedTime(&kernel_ms[i], h->prof_ev[i], h->prof_ev[i + 1]));
  return DVG_OK;
}

int dvg_gauss_lstm_step(dvg_lstm_t h, int variant, int rows, const float* x, int ldx, const void* state_in,
                        void* state_out, const float* eps, float* z, float* mu, float* logvar, dvg_stream_t stream) {
  DVG_REQUIRE(h && h->dims.kind == DVG_GAUSSIAN_LSTM, "handle is not a gaussian_lstm");
  DVG_REQUIRE(eps && z && mu && logvar, "null argument");
  return lstm_step_common(h, variant, rows, x, ldx, state_in, state_out, nullptr, 0, eps, z, mu, logvar, nullptr, 0,
                          (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------
// GP
// ---------------------------------------------------------------------------------------------------
static void gp_free_all(dvg_gp_s* h) {
  auto fr = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
  fr(h->z); fr(h->linv); fr(h->lqt); fr(h->alpha); fr(h->hyp); fr(h->work); fr(h->var_rows);
  fr(h->ticket); fr(h->trig_list); fr(h->trig_count); fr(h->linvT); fr(h->lq); fr(h->partial);
  fr(h->tc_img_v); fr(h->tc_img_w); fr(h->tc_alpha2);
  for (void* q : h->retired) cudaFree(q);
  h->retired.clear();
}

int dvg_gp_prepare(dvg_gp_t* out, const dvg_gp_dims* dims, const float* inducing, const float* var_mean,
                   const float* chol_var, const float* mean_const, const float* raw_outputscale,
                   const float* raw_lengthscale, const float* raw_noise, dvg_stream_t stream) {
  DVG_REQUIRE(out && dims, "null argument");
  DVG_REQUIRE(dims->num_dims > 0 && dims->num_inducing > 0, "bad sizes");
  DVG_REQUIRE(dims->num_inducing <= DVG_GP_MAX_INDUCING_ONDEVICE,
              "num_inducing=%d: the on-device fp64 factorisation handles at most %d inducing points; factorise on the "
              "host side and use dvg_gp_prepare_factors", dims->num_inducing, DVG_GP_MAX_INDUCING_ONDEVICE);
  dvg_gp_s* h = new (std::nothrow) dvg_gp_s();
  DVG_REQUIRE(h, "out of host memory");
  h->dims = *dims;
  h->mp = (int)align_up(dims->num_inducing, 4);
  const size_t D = dims->num_dims, M = dims->num_inducing, mp = h->mp;
  cudaError_t e = cudaGetDevice(&h->device);
  if (e == cudaSuccess) e = cudaMalloc(&h->z, sizeof(float) * D * mp);
  if (e == cudaSuccess) e = cudaMalloc(&h->linv, sizeof(float) * D * mp * mp);
  if (e == cudaSuccess) e = cudaMalloc(&h->lqt, sizeof(float) * D * mp * mp);
  if (e == cudaSuccess) e = cudaMalloc(&h->linvT, sizeof(float) * D * mp * mp);
  if (e == cudaSuccess) e = cudaMalloc(&h->lq, sizeof(float) * D * mp * mp);
  if (e == cudaSuccess) e = cudaMalloc(&h->alpha, sizeof(float) * D * mp);
  if (e == cudaSuccess) e = cudaMalloc(&h->hyp, sizeof(float) * D * 4);
  if (e == cudaSuccess) e = cudaMalloc(&h->work, sizeof(double) * D * M * M);
  h->var_rows_cap = 4096;
  if (e == cudaSuccess) e = cudaMalloc(&h->var_rows, sizeof(float) * 2 * D * h->var_rows_cap);   // two slots, like trig_list
  if (e == cudaSuccess) e = cudaMalloc(&h->trig_list, sizeof(int) * 2 * h->var_rows_cap);   // two slots: chained step launches alternate
  if (e == cudaSuccess) e = cudaMalloc(&h->trig_count, 2 * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&h->ticket, sizeof(unsigned int) * (4 + h->var_rows_cap / 8));
  if (e == cudaSuccess) e = cudaMemset(h->ticket, 0, sizeof(unsigned int) * (4 + h->var_rows_cap / 8));
  if (e == cudaSuccess) e = cudaMemset(h->trig_count, 0, 2 * sizeof(int));
  if (e != cudaSuccess) {
    set_error("GP handle allocation failed: %s", cudaGetErrorString(e));
    gp_free_all(h);
    delete h;
    return DVG_ERR_CUDA;
  }
  int rc = dvg_gp_refresh(h, inducing, var_mean, chol_var, mean_const, raw_outputscale, raw_lengthscale, raw_noise,
                          stream);
  if (rc) { gp_free_all(h); delete h; return rc; }
  *out = h;
  return DVG_OK;
}

int dvg_gp_refresh(dvg_gp_t h, const float* inducing, const float* var_mean, const float* chol_var,
                   const float* mean_const, const float* raw_outputscale, const float* raw_lengthscale,
                   const float* raw_noise, dvg_stream_t stream) {
  DVG_REQUIRE(h && inducing && var_mean && chol_var && mean_const && raw_outputscale && raw_lengthscale && raw_noise,
              "null argument");
  DVG_REQUIRE(!h->big, "handle holds pre-computed factors: use dvg_gp_refresh_factors");
  return gp_prepare_launch(h, inducing, var_mean, chol_var, mean_const, raw_outputscale, raw_lengthscale, raw_noise,
                           (cudaStream_t)stream);
}

size_t dvg_gp_factorize_workspace(const dvg_gp_dims* dims, int batch_dims) {
  if (!dims || dims->num_inducing <= 0) return 0;
  return gp_factorize_workspace(dims->num_inducing, batch_dims);
}

int dvg_gp_factorize(const dvg_gp_dims* dims, const float* inducing, const float* var_mean, const float* mean_const,
                     const float* raw_outputscale, const float* raw_lengthscale, float* linv, float* beta,
                     void* workspace, size_t workspace_bytes, dvg_stream_t stream) {
  DVG_REQUIRE(dims && inducing && var_mean && mean_const && raw_outputscale && raw_lengthscale && linv && beta,
              "null argument");
  DVG_REQUIRE(dims->num_dims > 0 && dims->num_inducing > 0 && dims->num_inducing <= 16384, "bad sizes");
  return gp_factorize(dims->num_dims, dims->num_inducing, (double)dims->jitter, inducing, var_mean, mean_const,
                      raw_outputscale, raw_lengthscale, linv, beta, workspace, workspace_bytes, (cudaStream_t)stream);
}

int dvg_gp_prepare_factors(dvg_gp_t* out, const dvg_gp_dims* dims, const float* inducing, const float* linv,
                           const float* lq, const float* beta, const float* hyp, dvg_stream_t stream) {
  DVG_REQUIRE(out && dims && inducing && linv && lq && beta && hyp, "null argument");
  DVG_REQUIRE(dims->num_dims > 0 && dims->num_inducing > 0, "bad sizes");
  dvg_gp_s* h = new (std::nothrow) dvg_gp_s();
  DVG_REQUIRE(h, "out of host memory");
  h->dims = *dims;
  h->big = true;
  h->mp = (int)align_up(dims->num_inducing, 64);
  const size_t D = dims->num_dims, mp = h->mp;
  cudaError_t e = cudaGetDevice(&h->device);
  if (e == cudaSuccess) e = cudaMalloc(&h->z, sizeof(float) * D * mp);
  if (e == cudaSuccess) e = cudaMalloc(&h->linv, sizeof(float) * D * mp * mp);
  if (e == cudaSuccess) e = cudaMalloc(&h->lqt, sizeof(float) * D * mp * mp);
  if (e == cudaSuccess) e = cudaMalloc(&h->alpha, sizeof(float) * D * mp);
  if (e == cudaSuccess) e = cudaMalloc(&h->hyp, sizeof(float) * D * 4);
  h->var_rows_cap = 4096;
  if (e == cudaSuccess) e = cudaMalloc(&h->var_rows, sizeof(float) * 2 * D * h->var_rows_cap);   // two slots, like trig_list
  if (e == cudaSuccess) e = cudaMalloc(&h->trig_list, sizeof(int) * 2 * h->var_rows_cap);   // two slots: chained step launches alternate
  if (e == cudaSuccess) e = cudaMalloc(&h->trig_count, 2 * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&h->ticket, sizeof(unsigned int) * (4 + h->var_rows_cap / 8));
  if (e == cudaSuccess) e = cudaMemset(h->ticket, 0, sizeof(unsigned int) * (4 + h->var_rows_cap / 8));
  if (e == cudaSuccess) e = cudaMemset(h->trig_count, 0, 2 * sizeof(int));
  if (e != cudaSuccess) {
    set_error("GP handle allocation failed: %s", cudaGetErrorString(e));
    gp_free_all(h);
    delete h;
    return DVG_ERR_CUDA;
  }
  int rc = gp_big_load_factors(h, inducing, linv, lq, beta, hyp, (cudaStream_t)stream);
  if (rc) { gp_free_all(h); delete h; return rc; }
  *out = h;
  return DVG_OK;
}

int dvg_gp_refresh_factors(dvg_gp_t h, const float* inducing, const float* linv, const float* lq, const float* beta,
                           const float* hyp, dvg_stream_t stream) {
  DVG_REQUIRE(h && h->big && inducing && linv && lq && beta && hyp, "bad argument (handle must come from dvg_gp_prepare_factors)");
  return gp_big_load_factors(h, inducing, linv, lq, beta, hyp, (cudaStream_t)stream);
}

int dvg_gp_destroy(dvg_gp_t h) {
  if (!h) return DVG_OK;
  gp_free_all(h);
  delete h;
  return DVG_OK;
}

int dvg_gp_predict(dvg_gp_t h, int n_rows, const float* x, int ldx, const int32_t* row_index, float* mean, int ldm,
                   float* var, int ldv, dvg_stream_t stream) {
  DVG_REQUIRE(h && x, "null argument");
  DVG_REQUIRE(n_rows >= 0 && ldx >= h->dims.num_dims, "bad n_rows/ldx");
  if (mean) DVG_REQUIRE(ldm >= h->dims.num_dims, "bad ldm");
  if (var) DVG_REQUIRE(ldv >= h->dims.num_dims, "bad ldv");
  if (h->big)
    return gp_big_predict_launch(h, n_rows, x, ldx, row_index, mean, ldm, 1, var, ldv, 1, (cudaStream_t)stream);
  return gp_predict_launch(h, n_rows, x, ldx, row_index, mean, ldm, var, ldv, (cudaStream_t)stream);
}

int dvg_gp_trigger(dvg_gp_t h, int n_rollouts, const float* x, int ldx, const int32_t* stat_rows, float* window,
                   int window_len, int32_t* count, int warmup, float factor, float* value, float* thr, uint8_t* mask,
                   dvg_stream_t stream) {
  DVG_REQUIRE(h && x && stat_rows && window && count, "null argument");
  DVG_REQUIRE(n_rollouts > 0 && ldx >= h->dims.num_dims, "bad n_rollouts/ldx");
  if (n_rollouts > h->var_rows_cap) {
    if (capturing((cudaStream_t)stream)) {
      set_error("trigger scratch too small for %d rollouts during stream capture", n_rollouts);
      return DVG_ERR_STATE;
    }
    DVG_CUDA(cudaDeviceSynchronize());
    if (h->var_rows) cudaFree(h->var_rows);
    h->var_rows = nullptr;
    h->var_rows_cap = 0;
    DVG_CUDA(cudaMalloc(&h->var_rows, sizeof(float) * 2 * (size_t)h->dims.num_dims * n_rollouts));
    if (h->trig_list) cudaFree(h->trig_list);
    h->trig_list = nullptr;
    DVG_CUDA(cudaMalloc(&h->trig_list, sizeof(int) * 2 * (size_t)n_rollouts));
    if (h->ticket) cudaFree(h->ticket);
    h->ticket = nullptr;
    DVG_CUDA(cudaMalloc(&h->ticket, sizeof(unsigned int) * (size_t)(4 + n_rollouts / 8)));
    DVG_CUDA(cudaMemset(h->ticket, 0, sizeof(unsigned int) * (size_t)(4 + n_rollouts / 8)));
    h->var_rows_cap = n_rollouts;
  }
  if (h->big)
    return gp_big_trigger_launch(h, n_rollouts, x, ldx, stat_rows, window, window_len, count, warmup, factor, value, thr,
                                 mask, (cudaStream_t)stream);
  return gp_trigger_launch(h, n_rollouts, x, ldx, stat_rows, window, window_len, count, warmup, factor, value, thr, mask,
                           (cudaStream_t)stream);
}

int dvg_gp_rsample(dvg_gp_t h, int n_rollouts, int n_points, const float* x, int ldx, const float* eps,
                   const uint8_t* mask, float* out, int ldo, dvg_stream_t stream) {
  DVG_REQUIRE(h && x && eps && out, "null argument");
  DVG_REQUIRE(n_rollouts > 0 && n_points > 0, "bad sizes");
  DVG_REQUIRE(n_points <= 128, "rsample correlates at most 128 points per call (got %d): the [N,N] covariance is factorised "
              "in shared memory", n_points);
  DVG_REQUIRE(ldx >= h->dims.num_dims && ldo >= h->dims.num_dims, "bad leading dimension");
  if (h->big) return gp_big_rsample_launch(h, n_rollouts, n_points, x, ldx, eps, mask, out, ldo, (cudaStream_t)stream);
  return gp_rsample_launch(h, n_rollouts, n_points, x, ldx, eps, mask, out, ldo, (cudaStream_t)stream);
}

int dvg_gp_export(dvg_gp_t h, float* linv, float* lq, float* alpha, float* hyp, dvg_stream_t stream) {
  DVG_REQUIRE(h, "null handle");
  const size_t D = h->dims.num_dims, mp = h->mp;
  cudaStream_t s = (cudaStream_t)stream;
  // raw padded copies: mp = M rounded up to 4
  if (linv) DVG_CUDA(cudaMemcpyAsync(linv, h->linv, sizeof(float) * D * mp * mp, cudaMemcpyDeviceToDevice, s));
  if (lq) DVG_CUDA(cudaMemcpyAsync(lq, h->lqt, sizeof(float) * D * mp * mp, cudaMemcpyDeviceToDevice, s));
  if (alpha) DVG_CUDA(cudaMemcpyAsync(alpha, h->alpha, sizeof(float) * D * mp, cudaMemcpyDeviceToDevice, s));
  if (hyp) DVG_CUDA(cudaMemcpyAsync(hyp, h->hyp, sizeof(float) * D * 4, cudaMemcpyDeviceToDevice, s));
  return DVG_OK;
}

int dvg_rollout_step(dvg_lstm_t h, dvg_gp_t g, int variant, int rows, const float* x, int ldx, const void* state_in,
                     void* state_out, float* y, int ldy, int n_rollouts, const int32_t* stat_rows, float* window,
                     int window_len, int32_t* count, int warmup, float factor, float* value, float* thr, uint8_t* mask,
                     const float* rs_eps, dvg_stream_t stream) {
  DVG_REQUIRE(h && g && x && state_in && state_out && y && stat_rows && window && count && mask, "null argument");
  DVG_REQUIRE(h->dims.kind == DVG_LSTM, "handle is not an lstm");
  DVG_REQUIRE(n_rollouts > 0 && rows % n_rollouts == 0, "rows must be a multiple of n_rollouts");
  DVG_REQUIRE(h->dims.input_size == g->dims.num_dims, "latent size mismatch between LSTM and GP");
  if (rs_eps) DVG_REQUIRE(h->dims.output_size == g->dims.num_dims, "rsample into y needs output_size == GP dims");
  cudaStream_t s = (cudaStream_t)stream;
  const bool tc = variant == DVG_BF16X3 || variant == DVG_BF16;
  const int n_points = rows / n_rollouts;
  if (tc && rows <= h->reserved_rows && n_rollouts <= g->var_rows_cap && window_len >= 1 && window_len <= 128 &&
      ldx >= h->dims.input_size && ldy >= h->dims.output_size && state_in != state_out &&
      lstm_tc_can_fuse_trigger(h, g, rows)) {
    const size_t lsz = (size_t)h->dims.n_layers * rows * h->dims.hidden_size;
    const size_t poff = dvg_lstm_state_packed_offset(h, rows);
    const float* h_in = (const float*)state_in;
    float* h_out = (float*)state_out;
    const bool rs_in_kernel = rs_eps != nullptr && !warmup && lstm_tc_can_fuse_rsample(g, n_points);
    int rc = lstm_tc_rollout_step(h, g, variant == DVG_BF16 ? 1 : 3, rows, x, ldx, h_in, h_in + lsz,
                                  (const uint8_t*)state_in + poff, h_out, h_out + lsz, (uint8_t*)state_out + poff, y, ldy,
                                  n_rollouts, stat_rows, window, window_len, count, warmup, factor, value, thr, mask,
                                  rs_in_kernel ? rs_eps : nullptr, s);
    if (rc || rs_in_kernel || rs_eps == nullptr || warmup) return rc;
    h->chain_ok = false;       // another kernel now sits between this step launch and the next
    return dvg_gp_rsample(g, n_rollouts, n_points, x, ldx, rs_eps, mask, y, ldy, stream);
  }
  if (tc && rows <= h->reserved_rows && window_len >= 1 && window_len <= 128 && ldx >= h->dims.input_size &&
      ldy >= h->dims.output_size && state_in != state_out && lstm_small_can_fuse_trigger(h, g, rows, n_rollouts)) {
    // small batches (<= 64 rows): the 16-CTA cluster kernel computes the trigger in its idle window and applies the
    // decision to its own cell updates; only the (rare) resample of fired rollouts is a second launch
    const size_t lsz = (size_t)h->dims.n_layers * rows * h->dims.hidden_size;
    const size_t poff = dvg_lstm_state_packed_offset(h, rows);
    const float* h_in = (const float*)state_in;
    float* h_out = (float*)state_out;
    StepTrigHost t{};
    t.S = n_rollouts; t.W = window_len; t.warmup = warmup; t.factor = factor; t.stat_rows = stat_rows; t.window = window;
    t.count = count; t.value = value; t.thr = thr; t.mask = mask;
    const bool rs_in_kernel = rs_eps != nullptr && !warmup && h->dims.output_size == g->dims.num_dims;
    t.rs_eps = rs_in_kernel ? rs_eps : nullptr;          // fired rollouts are resampled at the end of the same launch
    int rc = lstm_small_launch(h, variant == DVG_BF16 ? 1 : 3, rows, x, ldx, h_in, h_in + lsz, (const uint8_t*)state_in + poff,
                               h_out, h_out + lsz, (uint8_t*)state_out + poff, y, ldy, nullptr, n_points, s, g, &t);
    if (rc || rs_in_kernel || rs_eps == nullptr || warmup) return rc;
    return dvg_gp_rsample(g, n_rollouts, n_points, x, ldx, rs_eps, mask, y, ldy, stream);
  }
  // not fusable (fp32 variant, large inducing set, scratch not reserved): separate calls, same semantics
  int rc = dvg_gp_trigger(g, n_rollouts, x, ldx, stat_rows, window, window_len, count, warmup, factor, value, thr, mask,
                          stream);
  if (rc) return rc;
  rc = dvg_lstm_step(h, variant, rows, x, ldx, state_in, state_out, y, ldy, warmup ? nullptr : mask, n_points, stream);
  if (rc || rs_eps == nullptr || warmup) return rc;
  return dvg_gp_rsample(g, n_rollouts, n_points, x, ldx, rs_eps, mask, y, ldy, stream);
}

int dvg_eval_seq_finn(int n_frames, int n_samples, int n_seq, int channels, int height, int width, const float* gt,
                      const float* gen, float* ssim, float* psnr, dvg_stream_t stream) {
  DVG_REQUIRE(gt && gen && ssim && psnr, "null argument");
  DVG_REQUIRE(n_frames > 0 && n_samples > 0 && n_seq > 0 && channels > 0, "bad sizes");
  return eval_seq_finn_launch(n_frames, n_samples, n_seq, channels, height, width, gt, gen, ssim, psnr,
                              (cudaStream_t)stream);
}

int dvg_eval_seq(int n_frames, int n_samples, int n_seq, int channels, int height, int width, const float* gt,
                 const float* gen, float* ssim, float* psnr, dvg_stream_t stream) {
  DVG_REQUIRE(gt && gen && ssim && psnr, "null argument");
  DVG_REQUIRE(n_frames > 0 && n_samples > 0 && n_seq > 0 && channels > 0, "bad sizes");
  return eval_seq_skimage_launch(n_frames, n_samples, n_seq, channels, height, width, gt, gen, ssim, psnr,
                                 (cudaStream_t)stream);
}

int dvg_rollout_score(int n_steps, int n_rollouts, int n_points, int dim, const float* latents, const float* target,
                      float* scores, dvg_stream_t stream) {
  DVG_REQUIRE(latents && target && scores, "null argument");
  DVG_REQUIRE(n_steps > 0 && n_rollouts > 0 && n_points > 0 && dim > 0, "bad sizes");
  return rollout_score_launch(n_steps, n_rollouts, n_points, dim, latents, target, scores, (cudaStream_t)stream);
}

int dvg_moving_mnist_draws(int n_frames, int n_digits) { return n_digits * (5 + 4 * n_frames); }

int dvg_moving_mnist(int n_seq, int n_frames, int image_size, int n_digits, int deterministic, const float* digit_bank,
                     int n_bank, const uint32_t* draws, int draws_per_seq, int32_t* traj, float* frames,
                     dvg_stream_t stream) {
  DVG_REQUIRE(digit_bank && draws && traj && frames, "null argument");
  DVG_REQUIRE(n_seq > 0 && n_frames > 0 && n_digits > 0 && n_bank > 0, "bad sizes");
  DVG_REQUIRE(image_size > 33 && image_size % 4 == 0, "image_size must be a multiple of 4 and > 33 (got %d)", image_size);
  DVG_REQUIRE(draws_per_seq >= dvg_moving_mnist_draws(n_frames, n_digits),
              "draw stream too short: %d < %d words per sequence", draws_per_seq, dvg_moving_mnist_draws(n_frames, n_digits));
  return moving_mnist_launch(n_seq, n_frames, image_size, n_digits, deterministic, digit_bank, n_bank, draws,
                             draws_per_seq, traj, frames, (cudaStream_t)stream);
}

}  // extern "C"
