/*
 * dvg_b200.h -- C ABI of the B200-native DVG rollout hot path.
 *
 * The reference (shgaurav1/DVG) is pure Python/PyTorch and has no FFI of its own; the drop-in
 * boundary is the Python object protocol of models/lstm.py and models/gp_models.py (see
 * INTEGRATION.md).  This header is what that protocol binds to: every entry point below replaces the
 * library calls the cited reference lines issue on the GPU.  Plain C: raw device pointers, sizes and a
 * cudaStream_t; no torch types.  All calls are asynchronous on `stream` (no host sync, no D2H) unless
 * stated; they return 0 on success and a negative dvg_status on failure (never throw);
 * dvg_last_error() returns a thread-local description of the last failure.
 *
 * Ownership: every tensor passed in is owned by the caller (PyTorch).  A handle owns only its caches
 * (packed / padded weight copies, GP factors, scratch sized by the largest `rows` seen).  One handle
 * per (model, device); a handle is not re-entrant; independent handles are thread-safe.
 */
#ifndef DVG_B200_H
#define DVG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DVG_API __attribute__((visibility("default")))
#else
#define DVG_API
#endif

typedef struct CUstream_st* dvg_stream_t; /* == cudaStream_t */
typedef struct dvg_lstm_s* dvg_lstm_t;
typedef struct dvg_gp_s* dvg_gp_t;

enum dvg_status {
  DVG_OK = 0,
  DVG_ERR_ARG = -1,      /* bad argument / unsupported size for the requested variant */
  DVG_ERR_CUDA = -2,     /* a CUDA runtime call failed (see dvg_last_error) */
  DVG_ERR_ARCH = -3,     /* device is not sm_100 (tensor-core variants) */
  DVG_ERR_STATE = -4     /* handle / state block misuse */
};

/* GEMM arithmetic of the LSTM gate / embed / head contractions. */
enum dvg_variant {
  DVG_FP32 = 0,    /* CUDA-core FFMA, exact fp32 products (any sizes) */
  DVG_BF16X3 = 1,  /* tcgen05 kind::f16, operands split a = hi + lo (bf16), hi*hi + hi*lo + lo*hi,
                      fp32 accumulate in TMEM: fp32-grade (<=1e-4 rel.) -- needs hidden_size % 64 == 0 */
  DVG_BF16 = 2     /* tcgen05 kind::f16, single bf16 pass (2e-2 tolerance variant) */
};

enum dvg_lstm_kind {
  DVG_LSTM = 0,          /* models/lstm.py:42-72   embed -> L x LSTMCell -> Linear + Tanh */
  DVG_GAUSSIAN_LSTM = 1  /* models/lstm.py:140-175 embed -> L x LSTMCell -> mu_net, logvar_net, reparameterize */
};

DVG_API const char* dvg_last_error(void);
DVG_API int dvg_version(void);
/* sm count / compute capability of the current device (host query, synchronous). */
DVG_API int dvg_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * Frame-predictor LSTM / gaussian_lstm
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int kind;         /* dvg_lstm_kind */
  int input_size;   /* G_in  (embed.weight [H, G_in]) */
  int hidden_size;  /* H */
  int n_layers;     /* L */
  int output_size;  /* G_out (output.0.weight [G_out, H]) or Z (mu_net/logvar_net.weight [Z, H]) */
} dvg_lstm_dims;

/* Replaces the per-call weight reads of models/lstm.py:66,69,72 (and :172-173): packs the fp32
 * parameters once into kernel layouts (x||h concatenated along K, gate-interleaved N, padded, bf16
 * hi/lo tile images for the tensor-core variants).  Call again (dvg_lstm_refresh) after the parameters
 * change.  Pointers are device pointers in the reference's state_dict layout:
 *   embed_w [H,G_in], embed_b [H]; per layer l: w_ih[l] [4H,H], w_hh[l] [4H,H], b_ih[l] [4H], b_hh[l] [4H]
 *   (gate chunk order i,f,g,o); head0 = output.0 (or mu_net) weight [G_out,H] + bias; head1 = logvar_net
 *   (DVG_GAUSSIAN_LSTM only, else NULL).  The w_ih.. arrays are HOST arrays of L device pointers. */
DVG_API int dvg_lstm_prepare(dvg_lstm_t* out, const dvg_lstm_dims* dims,
                     const float* embed_w, const float* embed_b,
                     const float* const* w_ih, const float* const* w_hh,
                     const float* const* b_ih, const float* const* b_hh,
                     const float* head0_w, const float* head0_b,
                     const float* head1_w, const float* head1_b,
                     dvg_stream_t stream);
DVG_API int dvg_lstm_refresh(dvg_lstm_t h,
                     const float* embed_w, const float* embed_b,
                     const float* const* w_ih, const float* const* w_hh,
                     const float* const* b_ih, const float* const* b_hh,
                     const float* head0_w, const float* head0_b,
                     const float* head1_w, const float* head1_b,
                     dvg_stream_t stream);
DVG_API int dvg_lstm_destroy(dvg_lstm_t h);

/* Grow the handle's scratch so that steps with up to `rows` rows never allocate.  Synchronous (cudaMalloc,
 * cudaDeviceSynchronize).  Contract: reserve the largest row count BEFORE capturing step calls into a CUDA graph (a step
 * that would have to grow the scratch during capture fails with DVG_ERR_STATE); scratch that an already captured graph
 * may reference is retired, not freed, when the handle grows later, and released by dvg_lstm_destroy.  A handle is used
 * from one stream at a time: the dependency counters and scratch of the step kernels are per handle. */
DVG_API int dvg_lstm_reserve(dvg_lstm_t h, int rows);

/* Chained steps.  Between dvg_lstm_chain_begin and dvg_lstm_chain_end, consecutive dvg_lstm_step / dvg_rollout_step
 * calls on this handle whose state_in is the previous call's state_out (same rows, variant, stream) may OVERLAP on the
 * GPU: the next time step's first tiles start on the SMs that ran out of work while the previous step's last tiles
 * finish (the launches synchronise per tile through device counters instead of waiting for the whole previous grid).
 * The caller promises that between the two calls it enqueues NOTHING on the stream: every input of every step of the
 * chain other than the recurrent state (x, eps, rs_eps, stat_rows) is complete before dvg_lstm_chain_begin, and the
 * outputs (y, value, thr, mask) are only read by work enqueued after dvg_lstm_chain_end.  This is the situation of
 * a latent-space rollout whose inputs are known up front (teacher-forced context frames, the bench's hot-path loop);
 * with the encoder / decoder between the steps (generate_frames.py:266-298) do not open a chain.  Results are
 * identical with and without a chain.  Steps that cannot be chained (hold mask, small grids, the stand-alone resample
 * fallback) silently take the ordinary stream-ordered path.  Both calls may be captured in a CUDA graph. */
DVG_API int dvg_lstm_chain_begin(dvg_lstm_t h, dvg_stream_t stream);
DVG_API int dvg_lstm_chain_end(dvg_lstm_t h, dvg_stream_t stream);

/* Recurrent state block (replaces the list of (h,c) tuples of models/lstm.py:58-63).  One block holds
 *   h  fp32 [L][rows][H]   at byte offset 0
 *   c  fp32 [L][rows][H]   at byte offset L*rows*H*4
 *   hp bf16 hi/lo tile images of h (tensor-core A operand) at dvg_lstm_state_packed_offset()
 * A zero-filled block is a valid initial state (init_hidden).  The caller owns the memory. */
DVG_API size_t dvg_lstm_state_bytes(dvg_lstm_t h, int rows);
DVG_API size_t dvg_lstm_state_packed_offset(dvg_lstm_t h, int rows);
/* Rebuild the packed image from the fp32 h part (after the caller wrote h from outside). */
DVG_API int dvg_lstm_state_repack(dvg_lstm_t h, int rows, void* state, dvg_stream_t stream);

/* One time step of models/lstm.py:65-72 for `rows` independent rows.
 *   x [rows, G_in] row-major with leading dimension ldx (floats);  y [rows, G_out], ldy.
 *   state_in -> state_out (distinct blocks; state_in is not modified).
 *   hold (optional, may be NULL): u8 flags, one per group of `rows_per_flag` consecutive rows; rows whose
 *   flag is non-zero keep their state (state_out = state_in) -- the "LSTM is not advanced on a triggered
 *   step" semantics of generate_frames.py:289-295; y is still produced for them. */
DVG_API int dvg_lstm_step(dvg_lstm_t h, int variant, int rows,
                  const float* x, int ldx,
                  const void* state_in, void* state_out,
                  float* y, int ldy,
                  const uint8_t* hold, int rows_per_flag,
                  dvg_stream_t stream);

/* Measurement aid: dvg_lstm_step with CUDA events recorded on `stream` around the kernel launches; synchronises the
 * stream and returns the device time of each launch in kernel_ms.  The tensor-core variants issue ONE launch per step
 * (slot 0 = the whole step: lstm_step_kernel, or lstm_small_kernel for <= 64 rows at hidden size 256); the DVG_FP32
 * variant and the DVG_TC_FUSED=0 developer path issue one launch per GEMM (slots: 0 x-pack, 1 embed, 2..L+1 the LSTM
 * layers, L+2 head). */
DVG_API int dvg_lstm_profile(dvg_lstm_t h, int variant, int rows,
                     const float* x, int ldx,
                     const void* state_in, void* state_out,
                     float* y, int ldy,
                     float* kernel_ms, int max_slots,
                     dvg_stream_t stream);

/* One time step of models/lstm.py:166-175; eps [rows, Z] is the N(0,1) draw of :163.
 * Writes z, mu, logvar [rows, Z] (dense). */
DVG_API int dvg_gauss_lstm_step(dvg_lstm_t h, int variant, int rows,
                        const float* x, int ldx,
                        const void* state_in, void* state_out,
                        const float* eps, float* z, float* mu, float* logvar,
                        dvg_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Variational GP predictive + trigger  (models/gp_models.py:10-24 + gpytorch WhitenedVariationalStrategy
 * + GaussianLikelihood, call sites generate_frames.py:131,170,229,273,291; train.py:283)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int num_dims;      /* D (= g_dim) */
  int num_inducing;  /* M */
  float jitter;             /* 1e-3 (gpytorch add_jitter) */
  float noise_lower_bound;  /* 1e-4 (GaussianLikelihood GreaterThan constraint); 0 for the earliest 0.3.x */
} dvg_gp_dims;

/* Hoists everything that is constant in eval mode (the reference recomputes it per call): softplus
 * hyper-parameters, K_ZZ + jitter, its Cholesky factor (fp64 on device), L_ZZ^-1, beta = L_ZZ^-1(m_q - c),
 * masked L_q.  Pointers are device pointers in gpytorch 0.3.x state_dict layout:
 *   inducing [D,M,1], var_mean [D,M], chol_var [D,M,M] (raw, upper part ignored), mean_const [D,1],
 *   raw_outputscale [D], raw_lengthscale [D,1,1], raw_noise [D,1]. */
DVG_API int dvg_gp_prepare(dvg_gp_t* out, const dvg_gp_dims* dims,
                   const float* inducing, const float* var_mean, const float* chol_var,
                   const float* mean_const, const float* raw_outputscale, const float* raw_lengthscale,
                   const float* raw_noise, dvg_stream_t stream);
DVG_API int dvg_gp_refresh(dvg_gp_t h,
                   const float* inducing, const float* var_mean, const float* chol_var,
                   const float* mean_const, const float* raw_outputscale, const float* raw_lengthscale,
                   const float* raw_noise, dvg_stream_t stream);
DVG_API int dvg_gp_destroy(dvg_gp_t h);

/* Large inducing sets (BASELINE configs[4] sweeps M = 128 .. 4096).  dvg_gp_prepare factorises on the device in shared
 * memory and accepts up to DVG_GP_MAX_INDUCING_ONDEVICE points; beyond that -- and, for speed, from 65 points on, which
 * is what the Python host side does -- the eval-mode constants are computed by the caller (torch.linalg in fp64, the
 * library the reference itself relies on through gpytorch) and loaded here; dvg_gp_predict / dvg_gp_trigger then run
 * the tensor-core tiled kernels over them (gp_tc.cu: tcgen05, bf16x3 split operands, fp32 accumulation; DVG_GP_TC=0
 * selects the FP32 FFMA tiles of gp_big.cu) and dvg_gp_rsample the large-M resample kernel (n_points <= 128).  Device
 * pointers, fp32, dense row-major:
 *   inducing [D,M], linv [D,M,M] = chol(K_ZZ + jitter I)^-1 (lower), lq [D,M,M] = tril(chol_variational_covar),
 *   beta [D,M] = linv (m_q - c), hyp [D,4] = (lengthscale, outputscale, mean constant, noise incl. lower bound).
 */
#define DVG_GP_MAX_INDUCING_ONDEVICE 128
/* The factorisation for such a handle, natively: blocked fp64 Cholesky of K_ZZ + jitter I (64 x 64 blocks: diagonal block
 * factorised and inverted in shared memory, panel and trailing update by tiled FP64 GEMMs) and the triangular inverse by
 * block rows, batched over the latent dims -- what the reference recomputes with cuSOLVER potrf on every call.
 * Inputs as for dvg_gp_prepare (fp32 device pointers; dims->jitter is used); writes linv [D,M,M] and beta [D,M] (fp32)
 * for dvg_gp_prepare_factors.  The caller owns the workspace (device memory, 256-byte aligned):
 * dvg_gp_factorize_workspace(dims, b) bytes let b latent dims be processed per pass (b = 1 is the minimum; two
 * Mp x Mp fp64 matrices per dim).  Synchronises the stream once at the end to read the "not positive definite" flag
 * (DVG_ERR_ARG). */
DVG_API size_t dvg_gp_factorize_workspace(const dvg_gp_dims* dims, int batch_dims);
DVG_API int dvg_gp_factorize(const dvg_gp_dims* dims, const float* inducing, const float* var_mean, const float* mean_const,
                     const float* raw_outputscale, const float* raw_lengthscale, float* linv, float* beta,
                     void* workspace, size_t workspace_bytes, dvg_stream_t stream);
DVG_API int dvg_gp_prepare_factors(dvg_gp_t* out, const dvg_gp_dims* dims, const float* inducing, const float* linv,
                           const float* lq, const float* beta, const float* hyp, dvg_stream_t stream);
DVG_API int dvg_gp_refresh_factors(dvg_gp_t h, const float* inducing, const float* linv, const float* lq,
                           const float* beta, const float* hyp, dvg_stream_t stream);

/* likelihood(gp_layer(x)).mean / .variance for a set of latent rows.
 *   x [*, D] row-major (ldx floats): the [N,D] latent itself -- the reference's
 *   h.transpose(0,1).view(D,N,1) is only a strided view of it.
 *   row_index (optional): n_rows int32 row numbers to evaluate (gather); NULL = rows 0..n_rows-1.
 *   mean / var (either may be NULL): written at [i, d] for the i-th evaluated row, leading dims ldm / ldv. */
DVG_API int dvg_gp_predict(dvg_gp_t h, int n_rows, const float* x, int ldx, const int32_t* row_index,
                   float* mean, int ldm, float* var, int ldv, dvg_stream_t stream);

/* Fused trigger of generate_frames.py:227-232,275,283-289 for S independent rollouts, all on device:
 *   stat_rows [S] int32: the latent row whose variance drives rollout s (s*N + column);
 *   value[s] = || variance[stat_rows[s], :] ||_2 ; the window (fp32 [S, window_len]) slides;
 *   thr[s] = mean(window) + factor * std(window) (population std);  mask[s] = value > thr.
 *   warmup != 0: value is appended at window[s][*count] without a decision (mask = 0), the
 *   generate_frames.py:266-280 phase.  `count` is a device int32 (number of warm-up values stored). */
DVG_API int dvg_gp_trigger(dvg_gp_t h, int n_rollouts, const float* x, int ldx, const int32_t* stat_rows,
                   float* window, int window_len, int32_t* count, int warmup, float factor,
                   float* value, float* thr, uint8_t* mask, dvg_stream_t stream);

/* likelihood(gp_layer(x)).rsample() (generate_frames.py:171,292; train.py:284) for the rollouts whose
 * mask is set (mask NULL = all):  latent[s*N + n, d] = mean + (chol(Sigma_y) eps)[n] with
 * Sigma_y the full [N,N] predictive covariance of rollout s in dimension d.
 *   x [S*N, D] (ldx), eps [S, D, N] standard normal, out [S*N, D] (ldo) -- rows of unmasked rollouts are
 *   left untouched, so `out` can be the LSTM output buffer (on-device select).  N <= 128 (the batch size of one
 *   rollout; the reference correlates exactly the N points of one call).  The mask is read on the device at run time
 *   (no host copy, no stale state between calls). */
DVG_API int dvg_gp_rsample(dvg_gp_t h, int n_rollouts, int n_points, const float* x, int ldx,
                   const float* eps, const uint8_t* mask, float* out, int ldo, dvg_stream_t stream);

/* Debug / test access to the hoisted factors (device->device copies into caller buffers; any may be NULL),
 * with Mp = num_inducing rounded up to a multiple of 4 (zero padded):
 *   linv [D,Mp,Mp] (L_ZZ^-1, lower), lqt [D,Mp,Mp] (masked L_q, transposed), beta [D,Mp] (= L_ZZ^-1 (m_q - c)),
 *   hyp [D,4] = (ell, s, c, noise). */
DVG_API int dvg_gp_export(dvg_gp_t h, float* linv, float* lqt, float* alpha, float* hyp, dvg_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * N-diverse-futures bookkeeping
 * ---------------------------------------------------------------------------------------------- */
/* One GPtrigger_gen time step (generate_frames.py:266-298) for S = n_rollouts rollouts of rows/S points each:
 * dvg_gp_trigger(x), dvg_lstm_step(x, hold = warmup ? none : mask) and -- when rs_eps is given and the step is a
 * decision step -- dvg_gp_rsample(x, rs_eps, mask) into y: same arguments, same results, issued as ONE persistent
 * kernel when the tensor-core variants apply.  The trigger runs beside the gate GEMM tiles; nobody waits for its
 * mask: the LSTM advances every rollout and, in the rare step where rollouts fired, the end of the same launch
 * restores their state rows from state_in (a triggered rollout does not advance its LSTM, :289-295) and overwrites
 * their rows of y with the GP posterior sample (:291-292).  The mask never leaves the device.
 *   rs_eps  [n_rollouts, D, rows/n_rollouts] standard normal noise, or NULL (then follow with dvg_gp_rsample(mask)). */
DVG_API int dvg_rollout_step(dvg_lstm_t h, dvg_gp_t g, int variant, int rows,
                     const float* x, int ldx, const void* state_in, void* state_out, float* y, int ldy,
                     int n_rollouts, const int32_t* stat_rows, float* window, int window_len, int32_t* count,
                     int warmup, float factor, float* value, float* thr, uint8_t* mask, const float* rs_eps,
                     dvg_stream_t stream);

/* finn_eval_seq on the device (utils.py:237-301): channel-mean SSIM (11x11 Gaussian window,
 * sigma 1.5, K1=.01, K2=.03, L=1, NaN -> -1) and PSNR = 10 log10(1/mse) of every generated frame against the
 * ground truth, without the per-frame D2H copy of generate_frames.py:175-176.
 *   gt [T, B, C, H, W], gen [T, S, B, C, H, W] dense fp32  ->  ssim, psnr [S, B, T]. */
DVG_API int dvg_eval_seq_finn(int n_frames, int n_samples, int n_seq, int channels, int height, int width,
                      const float* gt, const float* gen, float* ssim, float* psnr, dvg_stream_t stream);

/* utils.eval_seq on the device (utils.py:220-234; the metric make_gifs selects on, generate_frames.py:178,188-189):
 * channel-mean SSIM and PSNR with the legacy skimage.measure.compare_ssim / compare_psnr defaults for float images --
 * 7x7 uniform window, sample covariance, K1=.01, K2=.03, data_range 2, mean over the interior cropped by 3 pixels;
 * PSNR = 10 log10(R^2/mse), R = 1 when min(gt) >= 0 else 2.  Same layouts as dvg_eval_seq_finn. */
DVG_API int dvg_eval_seq(int n_frames, int n_samples, int n_seq, int channels, int height, int width,
                 const float* gt, const float* gen, float* ssim, float* psnr, dvg_stream_t stream);

/* Device-side scoring pass of the best-of-N selection (the reference scores every sample on the host after a
 * D2H copy per frame, generate_frames.py:175-178,185-190): scores[s, b] = mean over (t, g) of
 * (latents[t, s*B + b, g] - target[t, b, g])^2.  latents [T, S*B, dim] dense, target [T, B, dim], scores [S, B]. */
DVG_API int dvg_rollout_score(int n_steps, int n_rollouts, int n_points, int dim, const float* latents,
                      const float* target, float* scores, dvg_stream_t stream);

/* Bouncing-digit batches on the device (data/moving_mnist.py:38-91, MovingMNIST.__getitem__, for n_seq samples at
 * once).  digit_bank [n_bank, 32, 32] fp32 in [0,1] (the 32x32-scaled MNIST digits, :22-25).
 * draws [n_seq, draws_per_seq] raw 32-bit integers: the k-th np.random.randint(lo, hi) call the reference makes for a
 * sample (digit index, sx, sy, dx, dy, then the velocity redraws at bounces, digit after digit) is
 * lo + draws[seq][k] % (hi - lo); draws_per_seq >= dvg_moving_mnist_draws(n_frames, n_digits) (the worst case).
 * deterministic != 0 mirrors velocities at the walls instead (:58,65,72,79).  traj is workspace and by-product:
 * int32 [n_seq, n_digits, 1 + 2*n_frames] = {digit index, (sx, sy) per frame}.  frames [n_frames, n_seq, 1, W, W] fp32
 * -- the list-of-frames layout utils.normalize_data produces (utils.py:86-95) -- = min(1, sum of the digits). */
DVG_API int dvg_moving_mnist_draws(int n_frames, int n_digits);
DVG_API int dvg_moving_mnist(int n_seq, int n_frames, int image_size, int n_digits, int deterministic,
                     const float* digit_bank, int n_bank, const uint32_t* draws, int draws_per_seq, int32_t* traj,
                     float* frames, dvg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DVG_B200_H */
