"""Sample-batched rollout drivers: the N-diverse-futures bookkeeping of the reference, re-designed so the
S futures of a batch run as S*B rows of ONE fused step instead of a sequential python loop.

Reference loops mirrored here (all sizes config-driven instead of the hard-coded 90/50/12/105/15):

* ``generate_frames.py:138-178``  make_gifs pass B: ``for s in range(nsample)`` x ``for i in range(1, n_eval)``
* ``train.py:262-289``            plot(): the same with nsample=5 and a resample only at i == 10
* ``generate_frames.py:249-300``  GPtrigger_gen: variance trigger, LSTM not advanced on a triggered step
* ``generate_frames.py:111-134``  make_gifs pass A: GP mean on the LSTM output

Design
------
``RolloutEngine`` owns every buffer of a rollout (two ping-pong LSTM state blocks, the latent output, the
per-rollout trigger window / value / threshold / mask) and issues, per time step, a fixed sequence of
C-ABI calls on the current stream with no host synchronisation and no D2H copy:

    trigger:  dvg_gp_trigger   (variance at the statistic row of each rollout -> window -> mask, on device)
    advance:  dvg_lstm_step    (all S*B rows; ``hold = mask`` keeps the state of triggered rollouts)
    resample: dvg_gp_rsample   (only CTAs of masked rollouts do work; overwrites their rows of the output)

Because the sequence is static it can be captured in a CUDA graph (``capture_latent_rollout``).
Rows are laid out rollout-major: row = s * B + b, so one rollout's B points are contiguous (rsample
correlates exactly those B points, generate_frames.py:171).
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _capi

TRIGGER_FACTOR = float(np.float32(2 + 0.01 * 1))   # generate_frames.py:288, depth == 1 (:254)


@dataclass
class RolloutConfig:
    n_points: int                 # B: batch sequences per rollout (N of the GP call)
    n_rollouts: int               # S: diverse futures resident on this GPU
    window: int = 12              # generate_frames.py:266 (warm-up length == window length)
    stat_col: int = 3             # generate_frames.py:230 hard-coded column
    stat_col_warmup: Optional[Sequence[int]] = None   # generate_frames.py:275 uses ``index``; default = stat_col
    variant: str = "bf16x3"
    trigger: bool = True          # False: manual-resample drivers only (make_gifs / plot); no trigger scratch, no stat_col


class RolloutEngine:
    def __init__(self, frame_predictor, gp_layer, likelihood, cfg: RolloutConfig):
        self.fp, self.gp, self.lik, self.cfg = frame_predictor, gp_layer, likelihood, cfg
        self.lib = _capi.load()
        self.S, self.B = cfg.n_rollouts, cfg.n_points
        self.R = self.S * self.B
        self.G = frame_predictor.output_size
        self.D = gp_layer.num_dims
        assert frame_predictor.input_size == self.D == self.G, "rollout needs g_dim in == out == GP dims"
        p = next(frame_predictor.parameters())
        if not p.is_cuda:
            raise _capi.DvgError("RolloutEngine needs CUDA modules; there is no CPU fallback")
        self.dev = p.device
        self.lrt = frame_predictor._runtime()
        self.grt = gp_layer._runtime(likelihood)
        self.variant = _capi.VARIANTS[cfg.variant if frame_predictor.hidden_size % 64 == 0 else "fp32"]
        _capi.check(self.lib.dvg_lstm_reserve(self.lrt.handle, self.R), "dvg_lstm_reserve")
        S, B, dev = self.S, self.B, self.dev
        self.blocks = [self.lrt.new_block(self.R, zero=True), self.lrt.new_block(self.R, zero=True)]
        self.cur = 0
        self.window = torch.zeros(S, cfg.window, device=dev)
        self.count = torch.zeros(1, dtype=torch.int32, device=dev)
        self.value = torch.zeros(S, device=dev)
        self.thr = torch.zeros(S, device=dev)
        self.mask = torch.zeros(S, dtype=torch.uint8, device=dev)
        self._mask_buf, self._value_buf = self.mask, self.value     # where the trigger writes (see latent_rollout)
        base = torch.arange(S, dtype=torch.int32) * B
        if not cfg.trigger:
            self.stat_rows = self.stat_rows_warmup = None
            self.reset()
            return
        # the reference indexes ``variance[:, 3]`` / ``[:, index]`` (generate_frames.py:230,275) and raises IndexError
        # for a batch that is too small; here an out-of-range column would read another rollout's row
        wcols = list(cfg.stat_col_warmup) if cfg.stat_col_warmup is not None else [cfg.stat_col] * S
        if not 0 <= cfg.stat_col < B or len(wcols) != S or any(not 0 <= int(c) < B for c in wcols):
            raise IndexError(f"trigger statistic column out of range for n_points={B}: stat_col={cfg.stat_col}, "
                             f"stat_col_warmup={cfg.stat_col_warmup}")
        self.stat_rows = (base + cfg.stat_col).to(dev)
        self.stat_rows_warmup = (base + torch.tensor(wcols, dtype=torch.int32)).to(dev)
        # trigger scratch must exist before any graph capture
        _capi.check(self.lib.dvg_gp_trigger(self.grt.handle, S, _capi.ptr(torch.zeros(self.R, self.D, device=dev)),
                                            self.D, _capi.ptr(self.stat_rows), _capi.ptr(self.window), cfg.window,
                                            _capi.ptr(self.count), 1, TRIGGER_FACTOR, None, None, None,
                                            _capi.stream_ptr()), "dvg_gp_trigger")
        self.reset()

    # ---- state -----------------------------------------------------------------------------------
    def reset(self):
        """frame_predictor.hidden = init_hidden() + empty trigger window for every rollout."""
        self.blocks[0].zero_()      # the other block is fully overwritten by the first step (pad rows are never stored)
        self.cur = 0
        self.window.zero_()
        self.count.zero_()
        self.mask.zero_()

    def hidden(self):
        """Current (h, c) views, reference layout (list of L tuples of [S*B, H])."""
        return self.lrt.views(self.blocks[self.cur], self.R)

    def load_broadcast_state(self, hidden_B):
        """Broadcast a B-row state (context phase computed once) to all S rollouts."""
        views = self.lrt.views(self.blocks[self.cur], self.R)
        for l, (h, c) in enumerate(hidden_B):
            views[l][0].copy_(h.repeat(self.S, 1))
            views[l][1].copy_(c.repeat(self.S, 1))
        _capi.check(self.lib.dvg_lstm_state_repack(self.lrt.handle, self.R, _capi.ptr(self.blocks[self.cur]),
                                                   _capi.stream_ptr()), "dvg_lstm_state_repack")

    # ---- per-step primitives (async, current stream) ---------------------------------------------
    def _ld(self, t):
        assert t.is_cuda and t.dtype == torch.float32 and t.shape == (self.R, self.G) and t.stride(1) == 1
        return t.stride(0)

    def trigger(self, h, warmup: bool):
        if self.stat_rows is None:
            raise _capi.DvgError("this RolloutEngine was built with trigger=False")
        rows = self.stat_rows_warmup if warmup else self.stat_rows
        _capi.check(self.lib.dvg_gp_trigger(self.grt.handle, self.S, _capi.ptr(h), self._ld(h), _capi.ptr(rows),
                                            _capi.ptr(self.window), self.cfg.window, _capi.ptr(self.count),
                                            1 if warmup else 0, TRIGGER_FACTOR, _capi.ptr(self.value),
                                            _capi.ptr(self.thr), _capi.ptr(self._mask_buf), _capi.stream_ptr()),
                    "dvg_gp_trigger")

    def advance(self, h, out, hold: bool):
        nxt = 1 - self.cur
        _capi.check(self.lib.dvg_lstm_step(self.lrt.handle, self.variant, self.R, _capi.ptr(h), self._ld(h),
                                           _capi.ptr(self.blocks[self.cur]), _capi.ptr(self.blocks[nxt]),
                                           _capi.ptr(out), self._ld(out), _capi.ptr(self._mask_buf) if hold else None,
                                           self.B, _capi.stream_ptr()), "dvg_lstm_step")
        self.cur = nxt

    def resample(self, h, eps, out, masked: bool):
        assert eps.shape == (self.S, self.D, self.B) and eps.is_contiguous()
        _capi.check(self.lib.dvg_gp_rsample(self.grt.handle, self.S, self.B, _capi.ptr(h), self._ld(h),
                                            _capi.ptr(eps), _capi.ptr(self._mask_buf) if masked else None, _capi.ptr(out),
                                            self._ld(out), _capi.stream_ptr()), "dvg_gp_rsample")

    # ---- fused steps -----------------------------------------------------------------------------
    def step_trigger_mode(self, h, eps, out, warmup: bool, resample: bool = True):
        """One GPtrigger_gen step (generate_frames.py:266-298) for all rollouts: ``out`` [S*B, G] receives
        the decoder input (LSTM prediction, or the GP sample for triggered rollouts)."""
        if self.stat_rows is None:
            raise _capi.DvgError("this RolloutEngine was built with trigger=False")
        rows = self.stat_rows_warmup if warmup else self.stat_rows
        nxt = 1 - self.cur
        rs = None
        if not warmup and resample:
            assert eps.shape == (self.S, self.D, self.B) and eps.is_contiguous()
            rs = _capi.ptr(eps)
        # ONE launch: trigger + LSTM step + (fired rollouts only) state restore and GP resample into `out`
        _capi.check(self.lib.dvg_rollout_step(self.lrt.handle, self.grt.handle, self.variant, self.R, _capi.ptr(h),
                                              self._ld(h), _capi.ptr(self.blocks[self.cur]),
                                              _capi.ptr(self.blocks[nxt]), _capi.ptr(out), self._ld(out), self.S,
                                              _capi.ptr(rows), _capi.ptr(self.window), self.cfg.window,
                                              _capi.ptr(self.count), 1 if warmup else 0, TRIGGER_FACTOR,
                                              _capi.ptr(self._value_buf), _capi.ptr(self.thr),
                                              _capi.ptr(self._mask_buf), rs, _capi.stream_ptr()), "dvg_rollout_step")
        self.cur = nxt

    def step_manual_mode(self, h, eps, out, resample: bool):
        """One make_gifs / plot step (generate_frames.py:166-174): LSTM always advances; on a resample step
        every rollout's decoder input is the GP sample of the *encoder* latent."""
        self.advance(h, out, hold=False)
        if resample:
            self.resample(h, eps, out, masked=False)

    @contextlib.contextmanager
    def chained(self):
        """Steps issued inside this block may overlap on the GPU (dvg_lstm_chain_begin / _end, include/dvg_b200.h: step
        t+1 starts on the SMs step t no longer needs).  The caller's promise: all step inputs other than the recurrent
        state exist before the block, nothing else is enqueued on the stream inside it, and each step writes its own
        output rows.  Results are identical to the unchained sequence."""
        _capi.check(self.lib.dvg_lstm_chain_begin(self.lrt.handle, _capi.stream_ptr()), "dvg_lstm_chain_begin")
        try:
            yield self
        finally:
            _capi.check(self.lib.dvg_lstm_chain_end(self.lrt.handle, _capi.stream_ptr()), "dvg_lstm_chain_end")

    # ---- latent-space rollout (hot path only; the bench / CUDA-graph unit) -------------------------
    def latent_rollout(self, lat, eps, out, warmup_steps: Optional[int] = None, masks=None, values=None):
        """Run T trigger-mode steps on pre-computed encoder latents ``lat`` [T, S*B, G] (stand-ins for
        encoder(x_in)), writing decoder inputs to ``out`` [T, S*B, G].  ``eps`` [T, S, D, B].
        Optional ``masks`` [T, S] u8 / ``values`` [T, S] record the trigger trace (device copies)."""
        W = self.cfg.window if warmup_steps is None else warmup_steps
        # every input of every step exists up front and nothing else is enqueued between the steps: chain the launches
        with self.chained():
            for t in range(lat.shape[0]):
                # the trigger writes its mask / value straight into row t of the trace buffers (no copy kernels)
                self._mask_buf = masks[t] if masks is not None else self.mask
                self._value_buf = values[t] if values is not None else self.value
                self.step_trigger_mode(lat[t], eps[t], out[t], warmup=t < W)
        self._mask_buf, self._value_buf = self.mask, self.value

    def capture_latent_rollout(self, lat, eps, out, masks=None, values=None, post=None):
        """CUDA-graph the whole T-step latent rollout (static launch sequence, zero host work per replay).
        ``reset()`` is captured too (the graph zeroes state and window first), so every replay starts from block 0
        whatever the parity of T."""
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.reset()
            self.latent_rollout(lat[:2], eps[:2], out[:2])     # warm the allocator / lazy init outside capture
            self.reset()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        if post is not None:
            with torch.cuda.stream(s):
                post()                                          # warm outside capture
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
        with torch.cuda.graph(g):
            self.reset()
            self.latent_rollout(lat, eps, out, masks=masks, values=values)
            if post is not None:
                post()          # e.g. the scoring pass of the best-of-N selection: same graph, no host launches
        self.cur = 0 if lat.shape[0] % 2 == 0 else 1
        return g


class LatentRolloutPipeline:
    """Host-facing streaming driver for ``RolloutEngine``: ``submit(lat_host, eps_host)`` enqueues one complete
    latent rollout whose inputs live in pinned HOST memory and returns a ticket; ``result(ticket)`` yields the
    pinned host results.  Three CUDA streams and two complete buffer sets -- each with its own captured CUDA graph, so
    the H2D copy lands directly in the buffer the graph reads and the D2H copy leaves from the buffer it wrote, no
    staging copies -- overlap the H2D copy of rollout k+1 and the D2H copy of rollout k-1 with the compute graph of
    rollout k (the per-step D2H + host numpy of the reference, generate_frames.py:175-176,230, is what this replaces).

    ``post(out_dev)`` (optional) is captured at the end of each graph (best-of-N scoring / selection) and returns a
    tuple of device tensors; they travel to the host with the trigger masks.  ``full_output=False`` skips the D2H
    copy of the complete [T, S*B, G] decoder-input tensor: as in the reference's own output stage only the selected
    futures leave the device (generate_frames.py:185-217; SURVEY 8e "only winner frames travel")."""

    def __init__(self, engine: "RolloutEngine", T: int, post=None, full_output: bool = True, capture_post: bool = True):
        self.eng = engine
        self.post, self.capture_post = post, capture_post
        dev, R, G, S, D, B = engine.dev, engine.R, engine.G, engine.S, engine.D, engine.B
        self.T, self.full_output = T, full_output
        self.lat = [torch.empty(T, R, G, device=dev) for _ in range(2)]
        self.lat16 = [None, None]
        self.eps = [torch.empty(T, S, D, B, device=dev) for _ in range(2)]
        self.out = [torch.empty(T, R, G, device=dev) for _ in range(2)]
        self.masks = [torch.zeros(T, S, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.extra = [None, None]
        self.graph = []
        for k in range(2):
            def hook(k=k):
                self.extra[k] = tuple(post(self.out[k]))
            self.graph.append(engine.capture_latent_rollout(self.lat[k], self.eps[k], self.out[k], masks=self.masks[k],
                                                            post=hook if post is not None and capture_post else None))
            if post is not None and not capture_post:      # (e.g. post holds NCCL collectives: enqueue it eagerly)
                hook()
        self.host_out = [torch.empty(T, R, G).pin_memory() if full_output else None for _ in range(2)]
        self.host_masks = [torch.empty(T, S, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.host_extra = [tuple(torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in (self.extra[k] or ()))
                           for k in range(2)]
        self.s_in, self.s_cmp, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_cmp = [torch.cuda.Event() for _ in range(2)]
        self.ev_out = [torch.cuda.Event() for _ in range(2)]
        self.n = 0
        for e in self.ev_out + self.ev_cmp:
            e.record()

    def d2h_bytes(self) -> int:
        n = self.host_masks[0].numel() + sum(t.numel() * t.element_size() for t in self.host_extra[0])
        return n + (self.host_out[0].numel() * 4 if self.full_output else 0)

    def submit(self, lat_host: torch.Tensor, eps_host: Optional[torch.Tensor] = None):
        """Enqueue one rollout.  ``eps_host`` None: the rsample noise is drawn on the device (torch.randn on the
        compute stream, like gpytorch's rsample does in the reference) instead of being shipped from the host."""
        k = self.n & 1
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.ev_cmp[k])               # the graph that read buffer set k two rollouts ago is done
            if lat_host.dtype == torch.bfloat16:
                # half-width transport: bf16 on the wire, expanded to the fp32 the kernels read on the device (exact
                # when the caller's latents are bf16-representable; the H2D copy is what bounds 8 ranks on one host)
                if self.lat16[k] is None:
                    self.lat16[k] = torch.empty(self.lat[k].shape, dtype=torch.bfloat16, device=self.lat[k].device)
                self.lat16[k].copy_(lat_host, non_blocking=True)
                self.lat[k].copy_(self.lat16[k])
            else:
                self.lat[k].copy_(lat_host, non_blocking=True)
            if eps_host is not None:
                self.eps[k].copy_(eps_host, non_blocking=True)
            self.ev_in[k].record()
        with torch.cuda.stream(self.s_cmp):
            self.s_cmp.wait_event(self.ev_in[k])
            self.s_cmp.wait_event(self.ev_out[k])              # out[k] / masks[k] / extra[k] drained by the D2H stream
            if eps_host is None:
                self.eps[k].normal_()
            self.graph[k].replay()
            if self.post is not None and not self.capture_post:
                for dst, src in zip(self.extra[k], self.post(self.out[k])):
                    dst.copy_(src)
            self.ev_cmp[k].record()
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_cmp[k])
            if self.full_output:
                self.host_out[k].copy_(self.out[k], non_blocking=True)
            self.host_masks[k].copy_(self.masks[k], non_blocking=True)
            for h, d in zip(self.host_extra[k], self.extra[k] or ()):
                h.copy_(d, non_blocking=True)
            self.ev_out[k].record()
        self.n += 1
        return self.n - 1

    def result(self, ticket):
        """(host outputs [T, S*B, G] or None, host masks [T, S], tuple of host copies of ``post``'s tensors)."""
        k = ticket & 1
        self.ev_out[k].synchronize()
        return self.host_out[k], self.host_masks[k], self.host_extra[k]

    def drain(self):
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.synchronize()


def score_rollouts(out: torch.Tensor, target: torch.Tensor, n_rollouts: int, n_points: int) -> torch.Tensor:
    """Latent-space best-of-N scoring on the device (one pass, ``dvg_rollout_score``): ``out`` [T, S*B, G],
    ``target`` [T, B, G] -> mean-squared-error scores [S, B] (lower is better)."""
    T, R, G = out.shape
    assert R == n_rollouts * n_points and target.shape == (T, n_points, G)
    assert out.is_cuda and out.is_contiguous() and target.is_contiguous()
    scores = torch.empty(n_rollouts, n_points, device=out.device, dtype=torch.float32)
    _capi.check(_capi.load().dvg_rollout_score(T, n_rollouts, n_points, G, _capi.ptr(out), _capi.ptr(target),
                                               _capi.ptr(scores), _capi.stream_ptr()), "dvg_rollout_score")
    return scores


def eval_seq_finn(gt: torch.Tensor, gen: torch.Tensor):
    """``utils.finn_eval_seq`` for all S samples at once, on the device: ``gt`` [T, B, C, H, W], ``gen``
    [T, S, B, C, H, W] -> (ssim, psnr) each [S, B, T].  Best-of-N as in generate_frames.py:188-189,207:
    ``shard.select_best(ssim.mean(2), higher_is_better=True)``."""
    T, S, B, C, H, W = gen.shape
    assert gt.shape == (T, B, C, H, W) and gt.is_cuda and gen.is_cuda
    gt, gen = gt.contiguous().float(), gen.contiguous().float()
    out = torch.empty(2, S, B, T, device=gen.device, dtype=torch.float32)
    _capi.check(_capi.load().dvg_eval_seq_finn(T, S, B, C, H, W, _capi.ptr(gt), _capi.ptr(gen), _capi.ptr(out[0]),
                                               _capi.ptr(out[1]), _capi.stream_ptr()), "dvg_eval_seq_finn")
    return out[0], out[1]


def eval_seq(gt: torch.Tensor, gen: torch.Tensor):
    """``utils.eval_seq`` (utils.py:220-234: legacy skimage ``compare_ssim`` / ``compare_psnr`` defaults -- the metric
    ``make_gifs`` ranks the samples by, generate_frames.py:178,188-189) for all S samples at once, on the device.
    Same layouts as :func:`eval_seq_finn`."""
    T, S, B, C, H, W = gen.shape
    assert gt.shape == (T, B, C, H, W) and gt.is_cuda and gen.is_cuda
    gt, gen = gt.contiguous().float(), gen.contiguous().float()
    out = torch.empty(2, S, B, T, device=gen.device, dtype=torch.float32)
    _capi.check(_capi.load().dvg_eval_seq(T, S, B, C, H, W, _capi.ptr(gt), _capi.ptr(gen), _capi.ptr(out[0]),
                                          _capi.ptr(out[1]), _capi.stream_ptr()), "dvg_eval_seq")
    return out[0], out[1]


# -------------------------------------------------------------------------------------------------------
# Pixel-space drivers (encoder / decoder are the reference conv nets on the stock PyTorch path)
# -------------------------------------------------------------------------------------------------------
def _rep(t, S):
    return t.repeat(S, *([1] * (t.dim() - 1)))


@torch.no_grad()
def generation(frame_predictor, encoder, decoder, x_in, skip):
    """``generation`` of generate_frames.py:220-224: one autoregressive pixel step, ``decoder([frame_predictor(
    encoder(x_in)[0]), skip])``; advances ``frame_predictor.hidden``."""
    h = encoder(x_in)[0]
    return decoder([frame_predictor(h), skip])


@torch.no_grad()
def var_value(gp_layer, likelihood, encoder, x_in, context_array, stat_col: int = 3):
    """``var_value`` of generate_frames.py:227-232 without the host round trip: the trigger statistic
    ``||variance[:, stat_col]||_2`` over the latent dims of the encoder latent of ``x_in`` and the window slid by
    one.  ``context_array`` is a 1-D float tensor (any device); both results stay on the GPU.  The batched,
    fused form of the same arithmetic is ``RolloutEngine.step_trigger_mode``."""
    h = encoder(x_in)[0]
    D = gp_layer.num_dims
    var = likelihood(gp_layer(h.transpose(0, 1).view(D, h.shape[0], 1))).variance          # [D, N]
    value = var[:, stat_col].float().norm()
    ctx = torch.as_tensor(context_array, dtype=torch.float32, device=value.device)
    return value, torch.cat([ctx[1:], value.reshape(1)])


@torch.no_grad()
def posterior_rollout(frame_predictor, gp_layer, likelihood, encoder, decoder, x, n_past, n_eval,
                      last_frame_skip=False):
    """generate_frames.py:111-134 (B rows).  Returns the list of n_eval frames."""
    frame_predictor.hidden = frame_predictor.init_hidden()
    gen = [x[0]]
    x_in = x[0]
    skip = None
    D = gp_layer.num_dims
    for i in range(1, n_eval):
        h, sk = encoder(x_in)
        if last_frame_skip or i < n_past:
            skip = sk
        if i < n_past:
            frame_predictor(h)
            x_in = x[i]
        else:
            h_pred = frame_predictor(h)
            pred = likelihood(gp_layer(h_pred.transpose(0, 1).view(D, h_pred.shape[0], 1)))
            x_in = decoder([pred.mean.transpose(0, 1), skip])
        gen.append(x_in)
    return gen


@torch.no_grad()
def diverse_rollout(frame_predictor, gp_layer, likelihood, encoder, decoder, x, n_past, n_eval, nsample,
                    eps: Optional[Dict] = None, resample_every: Optional[int] = 15,
                    resample_at: Optional[Sequence[int]] = None, last_frame_skip=False, variant="bf16x3",
                    record_latents=False, codec=None, engine: Optional["RolloutEngine"] = None,
                    eps_dev: Optional[torch.Tensor] = None, frames_out: Optional[torch.Tensor] = None):
    """generate_frames.py:138-178 / train.py:262-289 with the S samples batched.

    The context phase (i < n_past) is identical for every sample (teacher forcing, no randomness), so it
    is computed once with B rows; its LSTM state is then broadcast to the S*B rows of the engine.
    ``eps[(s, i)]`` ([D,B]) injects the rsample noise; missing entries are drawn with torch.randn.
    Returns frames ``gen[t]`` of shape [S, B, C, H, W] (t < n_past: ground truth broadcast).

    ``codec`` (a ``dvg_b200.codec.BatchedCodec`` of the same encoder / decoder) switches the conv stacks to the
    sample-batched execution of SURVEY §8f rank 2 (folded BatchNorm, channels-last, row chunks, skip half of the
    decoder convs computed once per sequence).  ``engine`` reuses a prebuilt ``RolloutEngine``; ``eps_dev``
    [n_hits, S, D, B] is the device-resident noise of the resample steps in order; ``frames_out``
    [n_eval - n_ctx, S*B, C, H, W] receives the generated frames -- with all three the call is a fixed launch
    sequence without host transfers (``PixelRollout`` captures it into one CUDA graph)."""
    B = x[0].shape[0]
    S = nsample
    D = gp_layer.num_dims
    dev = x[0].device
    enc = encoder if codec is None else codec.encode
    frame_predictor.hidden = frame_predictor.init_hidden()
    skip = None
    x_in = x[0]
    i = 1
    while i < min(n_past, n_eval):
        h, skip = enc(x_in)
        frame_predictor(h)
        x_in = x[i]
        i += 1
    eng = engine if engine is not None else RolloutEngine(
        frame_predictor, gp_layer, likelihood, RolloutConfig(n_points=B, n_rollouts=S, variant=variant, trigger=False))
    eng.cur = 0
    eng.load_broadcast_state(frame_predictor.hidden)
    n_ctx = i
    gens = [_rep(x[t], S).view(S, B, *x[t].shape[1:]) for t in range(n_ctx)] if frames_out is None else []
    lats = [None] * n_ctx
    xs = _rep(x_in, S)
    shared = codec is not None and not last_frame_skip and skip is not None
    if shared:
        codec.set_shared_skips(skip)
        skip_s = None
    else:
        skip_s = [_rep(sk, S) for sk in skip] if skip is not None else None
    out = torch.empty(S * B, D, device=dev)
    n_hit = 0
    for i in range(n_ctx, n_eval):
        if shared:
            h, sk = codec.encode(xs, want_skips=False)
        else:
            h, sk = enc(xs)
            if last_frame_skip or skip_s is None:
                skip_s = sk
        hit = (resample_every is not None and i % resample_every == 0) or (resample_at is not None and i in resample_at)
        e = None
        if hit and eps_dev is not None:
            e = eps_dev[n_hit]
            n_hit += 1
        elif hit:
            e = torch.stack([(eps[(s, i)] if eps is not None and (s, i) in eps else torch.randn(D, B)).to(dev)
                             for s in range(S)]).float().contiguous()
        eng.step_manual_mode(h.contiguous(), e, out, resample=hit)
        dst = frames_out[i - n_ctx] if frames_out is not None else None
        if shared:
            xs = codec.decode_shared(out, out=dst)
        else:
            xs = decoder([out, skip_s]) if codec is None else codec.decode(out, skip_s)
            if dst is not None:
                dst.copy_(xs)
                xs = dst
        if frames_out is None:
            gens.append(xs.view(S, B, *xs.shape[1:]))
        if record_latents:
            lats.append(out.clone().view(S, B, D))
    if frames_out is not None:
        return (frames_out, lats) if record_latents else frames_out
    return (gens, lats) if record_latents else gens


def resample_steps(n_past, n_eval, resample_every: Optional[int] = 15, resample_at: Optional[Sequence[int]] = None):
    """The steps of the sampling phase at which ``diverse_rollout`` draws a GP sample (generate_frames.py:167)."""
    return [i for i in range(min(n_past, n_eval), n_eval)
            if (resample_every is not None and i % resample_every == 0) or (resample_at is not None and i in resample_at)]


class PixelRollout:
    """Pass B of ``make_gifs`` (generate_frames.py:138-178) in pixel space as ONE fixed launch sequence: context
    phase on B rows, S*B-row sampling phase through ``BatchedCodec`` + ``RolloutEngine``; with ``graph=True`` the
    whole thing (all conv launches included) is captured into a CUDA graph and ``run`` is a replay -- what makes the
    small configurations (SM-MNIST, B=16: ~25 launches of a few microseconds per time step) launch-latency free.

    ``x`` static [n_ctx, B, C, W, W] (context frames), ``eps`` static [n_hits, S, D, B]; ``frames``
    [n_eval - n_ctx, S*B, C, W, W] is the result buffer."""

    def __init__(self, frame_predictor, gp_layer, likelihood, encoder, decoder, frame_shape, n_points, nsample, n_past,
                 n_eval, resample_every: Optional[int] = 15, resample_at: Optional[Sequence[int]] = None,
                 variant="bf16x3", codec_dtype=torch.float32, chunk_rows: Optional[int] = None, graph: bool = True):
        from .codec import BatchedCodec
        self.args = (frame_predictor, gp_layer, likelihood, encoder, decoder)
        self.B, self.S, self.n_past, self.n_eval = n_points, nsample, n_past, n_eval
        self.kw = dict(resample_every=resample_every, resample_at=resample_at, variant=variant)
        dev = next(frame_predictor.parameters()).device
        D = gp_layer.num_dims
        self.n_ctx = min(n_past, n_eval)
        self.hits = resample_steps(n_past, n_eval, resample_every, resample_at)
        self.codec = BatchedCodec(encoder, decoder, n_points, dtype=codec_dtype, chunk_rows=chunk_rows)
        self.engine = RolloutEngine(frame_predictor, gp_layer, likelihood,
                                    RolloutConfig(n_points=n_points, n_rollouts=nsample, variant=variant, trigger=False))
        self.x = torch.zeros(self.n_ctx, n_points, *frame_shape, device=dev)
        self.eps = torch.zeros(max(1, len(self.hits)), nsample, D, n_points, device=dev)
        self.frames = torch.empty(n_eval - self.n_ctx, nsample * n_points, *frame_shape, device=dev)
        self.graph = None
        if graph:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._launch()                                  # cuDNN autotune / lazy allocations outside capture
                self._launch()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._launch()

    def _launch(self):
        fp, gp, lik, enc, dec = self.args
        diverse_rollout(fp, gp, lik, enc, dec, list(self.x.unbind(0)), self.n_past, self.n_eval, self.S,
                        codec=self.codec, engine=self.engine, eps_dev=self.eps, frames_out=self.frames, **self.kw)

    def run(self, x=None, eps=None):
        """Copy the inputs into the static buffers (when given) and launch; returns ``frames`` (async)."""
        if x is not None:
            self.x.copy_(torch.stack(list(x[:self.n_ctx])) if not torch.is_tensor(x) else x[:self.n_ctx])
        if eps is not None:
            self.eps.copy_(eps)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._launch()
        return self.frames


@torch.no_grad()
def trigger_rollout(frame_predictor, gp_layer, likelihood, encoder, decoder, x0, n_rollouts, eps: Optional[Dict] = None,
                    warmup=12, n_steps=105, stat_col=3, stat_cols_warmup: Optional[Sequence[int]] = None,
                    skip_until=5, variant="bf16x3"):
    """generate_frames.py:249-300 with the outer ``for index in range(batch_size)`` loop batched: rollout s
    plays ``index = stat_cols_warmup[s]``.  ``eps[(s, i)]`` ([D,B]) is the rsample noise of rollout s at a
    triggered step i.  Returns dict(gen_seq [n_steps][S,B,...], values [n_steps,S], triggers [n_steps,S],
    latents [n_steps][S,B,D])."""
    B = x0.shape[0]
    S = n_rollouts
    D = gp_layer.num_dims
    dev = x0.device
    cols = list(stat_cols_warmup) if stat_cols_warmup is not None else [stat_col] * S
    eng = RolloutEngine(frame_predictor, gp_layer, likelihood,
                        RolloutConfig(n_points=B, n_rollouts=S, window=warmup, stat_col=stat_col,
                                      stat_col_warmup=cols, variant=variant))
    xs = _rep(x0, S)
    skip = None
    out = torch.empty(S * B, D, device=dev)
    gen_seq, latents = [], []
    values = torch.empty(n_steps, S, device=dev)
    trig = torch.zeros(n_steps, S, dtype=torch.uint8, device=dev)
    zero_eps = torch.zeros(S, D, B, device=dev)
    for i in range(n_steps):
        h, sk = encoder(xs)
        if i < skip_until:
            skip = sk
        h = h.contiguous()
        if i < warmup:
            e = zero_eps
        else:
            e = torch.stack([(eps[(s, i)] if eps is not None and (s, i) in eps else torch.randn(D, B)).to(dev)
                             for s in range(S)]).float().contiguous()
        eng.step_trigger_mode(h, e, out, warmup=i < warmup)
        values[i].copy_(eng.value)
        trig[i].copy_(eng.mask)
        xs = decoder([out, skip])
        gen_seq.append(xs.view(S, B, *xs.shape[1:]))
        latents.append(out.clone().view(S, B, D))
    return {"gen_seq": gen_seq, "values": values, "triggers": trig, "latents": latents}


@torch.no_grad()
def plot_rollout(frame_predictor, gp_layer, likelihood, encoder, decoder, x, n_past, n_eval, nsample=5,
                 eps: Optional[Dict] = None, resample_at: Sequence[int] = (10,), last_frame_skip=False, variant="bf16x3"):
    """The computational part of ``plot`` (train.py:256-335), the sampling pass run during training: ``nsample`` futures
    with ONE GP resample (at step 10, train.py:283-285), and per sequence the sample with the smallest summed squared
    pixel error over all ``n_eval`` frames (strict ``<`` scan of train.py:303-310 == first minimum).

    Returns dict(samples [n_eval][S,B,...], sse [S,B], best [B])."""
    samples = diverse_rollout(frame_predictor, gp_layer, likelihood, encoder, decoder, x, n_past, n_eval, nsample, eps=eps,
                              resample_every=None, resample_at=list(resample_at), last_frame_skip=last_frame_skip,
                              variant=variant)
    S, B = nsample, x[0].shape[0]
    sse = torch.zeros(S, B, device=x[0].device)
    for t in range(n_eval):
        d = samples[t].float() - x[t].float().unsqueeze(0)
        sse += d.reshape(S, B, -1).pow(2).sum(2)
    return {"samples": samples, "sse": sse, "best": sse.argmin(0)}


def make_gifs(frame_predictor, gp_layer, likelihood, encoder, decoder, x, n_past, n_eval, nsample,
              eps: Optional[Dict] = None, resample_every: Optional[int] = 15, last_frame_skip=False,
              variant="bf16x3", metric="skimage", codec=None):
    """The computational part of ``make_gifs`` (generate_frames.py:107-217) with everything on the device:
    pass A (approximate posterior, GP mean), pass B (``nsample`` diverse futures, batched), SSIM / PSNR of every
    generated frame against the ground truth and the best-of-N choice per sequence (the gif writing is out of scope).

    Returns dict(posterior [n_eval][B,...], samples [n_eval][S,B,...], ssim [B,S,T_f], psnr [B,S,T_f], best [B]).
    ``metric="skimage"`` (default) is ``utils.eval_seq`` -- what the script calls (generate_frames.py:178) -- restated
    from the documented legacy skimage defaults (skimage itself is absent here: unpinned); ``metric="finn"`` is the
    self-contained ``finn_eval_seq`` variant (utils.py:237-301), pinned to the reference's own functions.
    ``codec``: optional ``BatchedCodec`` for the S*B-row sampling pass (see ``diverse_rollout``)."""
    from . import shard
    posterior = posterior_rollout(frame_predictor, gp_layer, likelihood, encoder, decoder, x, n_past, n_eval,
                                  last_frame_skip)
    samples = diverse_rollout(frame_predictor, gp_layer, likelihood, encoder, decoder, x, n_past, n_eval, nsample,
                              eps=eps, resample_every=resample_every, last_frame_skip=last_frame_skip, variant=variant,
                              codec=codec)
    gt = torch.stack([x[t] for t in range(n_past, n_eval)])                    # [T_f, B, C, H, W]
    gen = torch.stack([samples[t] for t in range(n_past, n_eval)])             # [T_f, S, B, C, H, W]
    ssim, psnr = (eval_seq if metric == "skimage" else eval_seq_finn)(gt, gen)  # [S, B, T_f]
    best = shard.select_best(ssim.mean(2), higher_is_better=True)              # generate_frames.py:188-189,207
    return {"posterior": posterior, "samples": samples, "ssim": ssim.permute(1, 0, 2), "psnr": psnr.permute(1, 0, 2),
            "best": best}
