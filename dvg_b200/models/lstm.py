"""Drop-in ``lstm`` / ``gaussian_lstm`` (reference: models/lstm.py:42-72 and :140-175).

Same constructor signature, sub-module / parameter names (``embed``, ``lstm.{i}``, ``output.0``,
``mu_net``, ``logvar_net``), ``init_hidden()``, public ``hidden`` attribute (list of ``(h, c)`` tuples of
``[rows, H]`` tensors that callers assign from outside) and ``forward`` return values, so state_dicts and
whole-module pickles (train.py:380-383) interchange with the reference.

Under ``torch.no_grad()`` (the rollout loops of generate_frames.py / train.py ``plot``) ``forward`` is ONE
call into the C ABI (``dvg_lstm_step`` / ``dvg_gauss_lstm_step``), which runs the fused sm_100a kernels on
the caller's current CUDA stream.  With autograd enabled (train.py:175-248) the step is computed by the
same torch modules on the GPU so training keeps working; that path is not the product and carries no
parity claim.  CPU tensors are rejected: there is no CPU fallback.

Instances restored by unpickling never ran ``__init__``; every runtime attribute is created lazily.
"""
from __future__ import annotations

import contextlib
import os

import torch
import torch.nn as nn

from .. import _capi

_DEFAULT_VARIANT = os.environ.get("DVG_B200_VARIANT", "bf16x3")


def _device_for_state(mod: nn.Module):
    # models/lstm.py:61-62 allocates the state with .cuda() regardless of where the parameters are.
    if torch.cuda.is_available():
        p = next(mod.parameters(), None)
        if p is not None and p.is_cuda:
            return p.device
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


class _Runtime:
    """Per-module C-ABI handle + state-block bookkeeping (never pickled)."""

    def __init__(self, mod: nn.Module, kind: int):
        self.lib = _capi.load()
        self.kind = kind
        self.handle = None
        self.sig = None
        self.device = None
        self.refresh(mod)

    # -- weights -----------------------------------------------------------------------------------
    @staticmethod
    def _params(mod):
        ps = [mod.embed.weight, mod.embed.bias]
        for cell in mod.lstm:
            ps += [cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh]
        if hasattr(mod, "mu_net"):
            ps += [mod.mu_net.weight, mod.mu_net.bias, mod.logvar_net.weight, mod.logvar_net.bias]
        else:
            ps += [mod.output[0].weight, mod.output[0].bias]
        return ps

    def signature(self, mod):
        return tuple((p.data_ptr(), p._version) for p in self._params(mod))

    def refresh(self, mod):
        ps = self._params(mod)
        for p in ps:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise _capi.DvgError("dvg_b200 needs contiguous fp32 CUDA parameters (call .cuda() first); "
                                     "there is no CPU fallback")
        L = len(mod.lstm)
        cells = list(mod.lstm)
        arrs = [_capi.ptr_array([getattr(c, n).data for c in cells])
                for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        gauss = self.kind == _capi.DVG_GAUSSIAN_LSTM
        h0w, h0b = (mod.mu_net.weight, mod.mu_net.bias) if gauss else (mod.output[0].weight, mod.output[0].bias)
        h1w, h1b = (mod.logvar_net.weight, mod.logvar_net.bias) if gauss else (None, None)
        args = [_capi.ptr(mod.embed.weight), _capi.ptr(mod.embed.bias), arrs[0][0], arrs[1][0], arrs[2][0],
                arrs[3][0], _capi.ptr(h0w), _capi.ptr(h0b), _capi.ptr(h1w), _capi.ptr(h1b), _capi.stream_ptr()]
        with torch.cuda.device(ps[0].device):
            if self.handle is None:
                dims = _capi.LstmDims(self.kind, mod.embed.in_features, mod.embed.out_features, L,
                                      h0w.shape[0])
                hd = _capi.c_void_p()
                _capi.check(self.lib.dvg_lstm_prepare(_capi.ctypes.byref(hd), _capi.ctypes.byref(dims), *args),
                            "dvg_lstm_prepare")
                self.handle = hd
                self.device = ps[0].device
            else:
                _capi.check(self.lib.dvg_lstm_refresh(self.handle, *args), "dvg_lstm_refresh")
        self.sig = self.signature(mod)
        self.H = mod.embed.out_features
        self.L = L

    def __del__(self):
        try:
            if self.handle is not None:
                self.lib.dvg_lstm_destroy(self.handle)
        except Exception:
            pass

    # -- state blocks ------------------------------------------------------------------------------
    def new_block(self, rows, zero):
        nbytes = self.lib.dvg_lstm_state_bytes(self.handle, rows)
        mk = torch.zeros if zero else torch.empty
        return mk(nbytes, dtype=torch.uint8, device=self.device)

    def views(self, block, rows):
        n = self.L * rows * self.H
        f = block[: 2 * n * 4].view(torch.float32)
        h = f[:n].view(self.L, rows, self.H)
        c = f[n:].view(self.L, rows, self.H)
        hidden = [(h[l], c[l]) for l in range(self.L)]
        hidden[0][0]._dvg_block = block          # lets forward() recognise its own state without a copy
        hidden[0][0]._dvg_versions = tuple(t._version for hc in hidden for t in hc)
        return hidden

    def block_of(self, hidden, rows):
        """State block behind ``hidden``; imports (copy + repack) foreign tensors."""
        h0 = hidden[0][0]
        block = getattr(h0, "_dvg_block", None)
        if block is not None and block.numel() == self.lib.dvg_lstm_state_bytes(self.handle, rows):
            base = block.data_ptr()
            n = rows * self.H * 4
            ok = len(hidden) == self.L
            for l in range(self.L if ok else 0):
                h, c = hidden[l]
                ok = ok and h.data_ptr() == base + l * n and c.data_ptr() == base + (self.L + l) * n \
                    and h.shape == (rows, self.H) and c.shape == (rows, self.H)
            vers = tuple(t._version for hc in hidden for t in hc)
            if ok:
                if h0._dvg_versions != vers:         # caller wrote into our views in place
                    _capi.check(self.lib.dvg_lstm_state_repack(self.handle, rows, _capi.ptr(block),
                                                               _capi.stream_ptr()), "dvg_lstm_state_repack")
                    h0._dvg_versions = vers
                return block
        if len(hidden) != self.L:
            raise _capi.DvgError(f"hidden has {len(hidden)} layers, model has {self.L}")
        block = self.new_block(rows, zero=False)
        views = self.views(block, rows)
        for l in range(self.L):
            views[l][0].copy_(hidden[l][0].detach().to(self.device, torch.float32).reshape(rows, self.H))
            views[l][1].copy_(hidden[l][1].detach().to(self.device, torch.float32).reshape(rows, self.H))
        views[0][0]._dvg_versions = tuple(t._version for hc in views for t in hc)
        _capi.check(self.lib.dvg_lstm_state_repack(self.handle, rows, _capi.ptr(block), _capi.stream_ptr()),
                    "dvg_lstm_state_repack")
        return block


class _FastLstmBase(nn.Module):
    _dvg_kind = _capi.DVG_LSTM

    # ---- pickling / lazy runtime -----------------------------------------------------------------
    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_dvg_rt", None)
        hid = state.get("hidden")
        if hid is not None:
            state["hidden"] = [(h.detach().clone(), c.detach().clone()) for h, c in hid]
        return state

    def _runtime(self) -> _Runtime:
        rt = self.__dict__.get("_dvg_rt")
        if rt is None:
            rt = _Runtime(self, self._dvg_kind)
            self.__dict__["_dvg_rt"] = rt
        elif rt.sig != rt.signature(self):
            rt.refresh(self)     # parameters were updated (optimizer step, load_state_dict, .to())
        return rt

    @property
    def gemm_variant(self) -> str:
        v = self.__dict__.get("_dvg_variant", _DEFAULT_VARIANT)
        if v != "fp32" and self.hidden_size % 64 != 0:
            return "fp32"          # tensor-core tiles need H % 64 == 0; the FFMA variant covers the rest
        return v

    @gemm_variant.setter
    def gemm_variant(self, v: str):
        if v not in _capi.VARIANTS:
            raise ValueError(f"unknown variant {v!r}; choose from {sorted(_capi.VARIANTS)}")
        self.__dict__["_dvg_variant"] = v

    @contextlib.contextmanager
    def chained(self):
        """Steps issued inside this block may overlap on the GPU (dvg_lstm_chain_begin / _end, include/dvg_b200.h).  Only for
        loops whose inputs all exist before the block and that enqueue nothing else on the stream between the steps --
        e.g. the teacher-forced context frames of a sequence; NOT the generation loop with the decoder / encoder between
        the steps.  Results are identical."""
        rt = self._runtime()
        _capi.check(rt.lib.dvg_lstm_chain_begin(rt.handle, _capi.stream_ptr()), "dvg_lstm_chain_begin")
        try:
            yield self
        finally:
            _capi.check(rt.lib.dvg_lstm_chain_end(rt.handle, _capi.stream_ptr()), "dvg_lstm_chain_end")

    def init_hidden(self):
        """models/lstm.py:58-63 -- L tuples of zeros [batch_size, H] (views of one zeroed state block
        when the CUDA runtime is up, so the first forward needs no import copy)."""
        dev = _device_for_state(self)
        if dev.type == "cuda" and next(self.parameters()).is_cuda:
            rt = self._runtime()
            return rt.views(rt.new_block(self.batch_size, zero=True), self.batch_size)
        return [(torch.zeros(self.batch_size, self.hidden_size, device=dev),
                 torch.zeros(self.batch_size, self.hidden_size, device=dev)) for _ in range(self.n_layers)]

    def _use_fast_path(self, x):
        if not x.is_cuda:
            raise _capi.DvgError("dvg_b200 hot path is CUDA-only (no CPU fallback); got a CPU tensor")
        # eval() mode always takes the kernels: every rollout loop of the reference that runs in eval mode without
        # torch.no_grad() (GPtrigger_gen, generate_frames.py:249-300; plot(), train.py:256-289 after :372) detaches
        # the prediction at once (generate_frames.py:222, train.py:280), so the autograd graph a stock module would
        # build there is never used.  train() mode with autograd on delegates to the stock torch ops on the GPU.
        # (``eval_uses_kernels = False`` on an instance restores grad-in-eval.)
        if not self.training and getattr(self, "eval_uses_kernels", True):
            return True
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return False
        return True

    def _prep_input(self, input):
        x = input.reshape(-1, self.input_size)
        if x.dtype != torch.float32:
            x = x.float()
        if x.stride(-1) != 1 or (x.shape[0] > 1 and x.stride(0) < self.input_size):
            x = x.contiguous()
        return x

    def _trunk_autograd(self, input):
        embedded = self.embed(input.view(-1, self.input_size))
        h_in = embedded
        for i in range(self.n_layers):
            self.hidden[i] = self.lstm[i](h_in, self.hidden[i])
            h_in = self.hidden[i][0]
        return h_in


class lstm(_FastLstmBase):
    """Frame predictor (models/lstm.py:42-72)."""
    _dvg_kind = _capi.DVG_LSTM

    def __init__(self, input_size, output_size, hidden_size, n_layers, batch_size):
        super().__init__()
        self.input_size = input_size
        self.output_size = output_size
        self.hidden_size = hidden_size
        self.batch_size = batch_size
        self.n_layers = n_layers
        self.embed = nn.Linear(input_size, hidden_size)
        self.lstm = nn.ModuleList([nn.LSTMCell(hidden_size, hidden_size) for _ in range(self.n_layers)])
        self.output = nn.Sequential(nn.Linear(hidden_size, output_size), nn.Tanh())
        self.hidden = self.init_hidden()

    def forward(self, input, hold=None, rows_per_flag=1):
        """``hold`` (optional u8 CUDA tensor, one flag per ``rows_per_flag`` rows) keeps the state of the
        flagged rows -- the not-advanced-on-trigger semantics of generate_frames.py:289-295."""
        if not self._use_fast_path(input):
            return self.output(self._trunk_autograd(input))
        rt = self._runtime()
        x = self._prep_input(input)
        rows = x.shape[0]
        blk_in = rt.block_of(self.hidden, rows)
        blk_out = rt.new_block(rows, zero=False)
        y = torch.empty(rows, self.output_size, dtype=torch.float32, device=x.device)
        _capi.check(rt.lib.dvg_lstm_step(rt.handle, _capi.VARIANTS[self.gemm_variant], rows, _capi.ptr(x),
                                         x.stride(0) if rows > 1 else self.input_size, _capi.ptr(blk_in),
                                         _capi.ptr(blk_out), _capi.ptr(y), self.output_size, _capi.ptr(hold),
                                         rows_per_flag, _capi.stream_ptr()), "dvg_lstm_step")
        self.hidden = rt.views(blk_out, rows)
        return y


class gaussian_lstm(_FastLstmBase):
    """Prior / posterior cell (models/lstm.py:140-175)."""
    _dvg_kind = _capi.DVG_GAUSSIAN_LSTM

    def __init__(self, input_size, output_size, hidden_size, n_layers, batch_size):
        super().__init__()
        self.input_size = input_size
        self.output_size = output_size
        self.hidden_size = hidden_size
        self.n_layers = n_layers
        self.batch_size = batch_size
        self.embed = nn.Linear(input_size, hidden_size)
        self.lstm = nn.ModuleList([nn.LSTMCell(hidden_size, hidden_size) for _ in range(self.n_layers)])
        self.mu_net = nn.Linear(hidden_size, output_size)
        self.logvar_net = nn.Linear(hidden_size, output_size)
        self.hidden = self.init_hidden()

    def reparameterize(self, mu, logvar):
        """models/lstm.py:161-164 (torch ops; used by the autograd path)."""
        logvar = logvar.mul(0.5).exp_()
        eps = logvar.data.new(logvar.size()).normal_()
        return eps.mul(logvar).add_(mu)

    def forward(self, input, eps=None):
        """Returns (z, mu, logvar).  ``eps`` injects the N(0,1) draw of models/lstm.py:163; when omitted it
        is drawn exactly like the reference (``.normal_()`` on the default generator)."""
        if not self._use_fast_path(input):
            h_in = self._trunk_autograd(input)
            mu = self.mu_net(h_in)
            logvar = self.logvar_net(h_in)
            if eps is not None:
                return eps.mul(logvar.mul(0.5).exp()).add(mu), mu, logvar
            return self.reparameterize(mu, logvar), mu, logvar
        rt = self._runtime()
        x = self._prep_input(input)
        rows = x.shape[0]
        Z = self.output_size
        if eps is None:
            eps = torch.empty(rows, Z, dtype=torch.float32, device=x.device).normal_()
        else:
            eps = eps.to(x.device, torch.float32).reshape(rows, Z).contiguous()
        blk_in = rt.block_of(self.hidden, rows)
        blk_out = rt.new_block(rows, zero=False)
        out = torch.empty(3, rows, Z, dtype=torch.float32, device=x.device)
        _capi.check(rt.lib.dvg_gauss_lstm_step(rt.handle, _capi.VARIANTS[self.gemm_variant], rows, _capi.ptr(x),
                                               x.stride(0) if rows > 1 else self.input_size, _capi.ptr(blk_in),
                                               _capi.ptr(blk_out), _capi.ptr(eps), _capi.ptr(out[0]),
                                               _capi.ptr(out[1]), _capi.ptr(out[2]), _capi.stream_ptr()),
                    "dvg_gauss_lstm_step")
        self.hidden = rt.views(blk_out, rows)
        return out[0], out[1], out[2]
