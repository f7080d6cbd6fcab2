"""Drop-in ``GPRegressionLayer1`` + ``GaussianLikelihood`` (reference: models/gp_models.py:10-24 and the
``gpytorch.likelihoods.GaussianLikelihood(batch_size=g_dim)`` of generate_frames.py:67 / train.py:102).

gpytorch-free: the parameters live in plain ``nn.Module`` containers whose ``state_dict`` keys are the
gpytorch 0.3.x names the reference checkpoints hold (train.py:384-386):

    gp_layer:   variational_strategy.inducing_points [D,M,1]
                variational_strategy.variational_distribution.variational_mean [D,M]
                variational_strategy.variational_distribution.chol_variational_covar [D,M,M]
                variational_strategy.variational_params_initialized (0-d buffer)
                mean_module.constant [D,1]
                covar_module.raw_outputscale [D]
                covar_module.base_kernel.raw_lengthscale [D,1,1]
    likelihood: noise_covar.raw_noise [D,1]

Call protocol (generate_frames.py:131,170,229,273,291; train.py:283):
``pred = likelihood(gp_layer(h.transpose(0,1).view(D,N,1)))`` then ``pred.mean`` / ``pred.variance``
(``[D,N]``) or ``pred.rsample()`` (``[D,N]``).  Evaluation is lazy: nothing is computed until one of the
three is read, and each read is a single C-ABI call (``dvg_gp_predict`` / ``dvg_gp_rsample``).  The
eval-mode constants the reference recomputes on every call (K_ZZ, its Cholesky factor, K_ZZ^-1(m-c)) are
hoisted into ``dvg_gp_prepare`` and refreshed only when a parameter changes.

The kernels own the eval-mode predictive.  In ``train()`` mode ``forward`` delegates to ``gp_train.train_forward``
(plain torch ops with autograd on the GPU: diag-only data covariance, KL memo, VariationalELBO), so the reference's
training steps (train.py:146-248) run against these classes (minimal form; no backward kernels).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import _capi

JITTER = 1e-3
NOISE_LOWER_BOUND = 1e-4
MAX_INDUCING_ONDEVICE = 128      # include/dvg_b200.h: DVG_GP_MAX_INDUCING_ONDEVICE (limit of the on-device fp64 factorisation)
# Inducing sets above this size take the tiled path (pre-computed factors + tensor-core tiles, gp_tc.cu / gp_big.cu): the
# shared-memory kernels, built for the reference's M = 40, reach 4 TFLOP/s at M = 128 where the tiled path reaches > 100.
TILED_FROM = 64


class _Holder(nn.Module):
    """Plain parameter container (keeps gpytorch's nested state_dict names)."""


class GaussianLikelihood(nn.Module):
    def __init__(self, batch_size=1, noise_lower_bound=NOISE_LOWER_BOUND):
        super().__init__()
        self.noise_covar = _Holder()
        self.noise_covar.raw_noise = nn.Parameter(torch.zeros(batch_size, 1))
        self.noise_lower_bound = noise_lower_bound

    @property
    def noise(self):
        return nn.functional.softplus(self.noise_covar.raw_noise) + self.noise_lower_bound

    def forward(self, pred):
        """``likelihood(gp_layer(x))``: eval -> lazy GPPrediction with the noise folded in; train -> the autograd
        prediction with ``noise`` added to its variance."""
        return pred.with_likelihood(self)


class _GpRuntime:
    def __init__(self, layer, likelihood):
        self.lib = _capi.load()
        self.handle = None
        self.sig = None
        self.refresh(layer, likelihood)

    @staticmethod
    def _tensors(layer, likelihood):
        vs = layer.variational_strategy
        ts = [vs.inducing_points, vs.variational_distribution.variational_mean,
              vs.variational_distribution.chol_variational_covar, layer.mean_module.constant,
              layer.covar_module.raw_outputscale, layer.covar_module.base_kernel.raw_lengthscale]
        if likelihood is not None:
            ts.append(likelihood.noise_covar.raw_noise)
        return ts

    def signature(self, layer, likelihood):
        return tuple((t.data_ptr(), t._version) for t in self._tensors(layer, likelihood))

    def refresh(self, layer, likelihood):
        ts = self._tensors(layer, likelihood)
        for t in ts:
            if not t.is_cuda or t.dtype != torch.float32:
                raise _capi.DvgError("dvg_b200 GP needs fp32 CUDA parameters (call .cuda() first); no CPU fallback")
        dev = ts[0].device
        D, M = ts[1].shape
        if likelihood is None:     # latent f (no observation noise): softplus(-100) == 0, lower bound 0
            raw_noise = torch.full((D, 1), -100.0, device=dev)
            lb = 0.0
        else:
            raw_noise = ts[6]
            lb = float(likelihood.noise_lower_bound)
        with torch.cuda.device(dev):
            if M > TILED_FROM:
                self._refresh_factors(ts, raw_noise, lb, D, M)
            else:
                args = [_capi.ptr(t.detach().contiguous()) for t in ts[:6]] + [_capi.ptr(raw_noise.detach().contiguous())]
                if self.handle is None:
                    dims = _capi.GpDims(D, M, JITTER, lb)
                    hd = _capi.c_void_p()
                    _capi.check(self.lib.dvg_gp_prepare(_capi.ctypes.byref(hd), _capi.ctypes.byref(dims), *args,
                                                        _capi.stream_ptr()), "dvg_gp_prepare")
                    self.handle = hd
                else:
                    _capi.check(self.lib.dvg_gp_refresh(self.handle, *args, _capi.stream_ptr()), "dvg_gp_refresh")
        self._keep = raw_noise
        self.D, self.M, self.device = D, M, dev
        self.sig = self.signature(layer, likelihood)

    def _refresh_factors(self, ts, raw_noise, lb, D, M):
        """Large inducing sets (M > 64, BASELINE configs[4]): the eval-mode constants -- L = chol(K_ZZ + jitter I),
        Linv = L^-1, beta = Linv (m_q - c), L_q = tril(chol_variational_covar) -- are computed once per weight load and
        handed to the C ABI, whose tiled kernels own the per-call work.  The factorisation itself is the library's
        blocked fp64 kernels (``dvg_gp_factorize``, csrc/gp_factor.cu); ``DVG_GP_NATIVE_FACTOR=0`` takes torch.linalg
        (cuSOLVER, what gpytorch itself calls for this step) instead -- tests/test_gpu_gp.py holds one against the other."""
        Z, m_q, chol_var, c_raw, raw_os, raw_ls = [t.detach() for t in ts[:6]]
        f64 = torch.float64
        sp = nn.functional.softplus
        ell = sp(raw_ls.to(f64).reshape(D))
        s = sp(raw_os.to(f64).reshape(D))
        c = c_raw.to(f64).reshape(D)
        noise = sp(raw_noise.detach().to(f64).reshape(D)) + lb
        hyp = torch.stack([ell, s, c, noise], dim=1).float().contiguous()
        z = Z.reshape(D, M)
        linv = torch.empty(D, M, M, dtype=torch.float32, device=z.device)
        beta = torch.empty(D, M, dtype=torch.float32, device=z.device)
        if os.environ.get("DVG_GP_NATIVE_FACTOR", "1") != "0":
            dims = _capi.GpDims(D, M, JITTER, lb)
            keep = [z.float().contiguous(), m_q.float().contiguous(), c_raw.float().reshape(D).contiguous(),
                    raw_os.float().reshape(D).contiguous(), raw_ls.float().reshape(D).contiguous()]
            batch = max(1, min(D, (1 << 28) // (M * M)))           # ~2 GB per fp64 operand at most
            ws = torch.empty(self.lib.dvg_gp_factorize_workspace(_capi.ctypes.byref(dims), batch), dtype=torch.uint8,
                             device=z.device)
            _capi.check(self.lib.dvg_gp_factorize(_capi.ctypes.byref(dims), *[_capi.ptr(t) for t in keep], _capi.ptr(linv),
                                                  _capi.ptr(beta), _capi.ptr(ws), ws.numel(), _capi.stream_ptr()),
                        "dvg_gp_factorize")
            del ws
        else:
            eye = torch.eye(M, dtype=f64, device=z.device)
            step = max(1, min(D, (1 << 28) // (M * M)))           # bound the fp64 temporaries (~2 GB per operand)
            for d0 in range(0, D, step):
                d1 = min(D, d0 + step)
                zz = z[d0:d1].to(f64)
                t = (zz[:, :, None] - zz[:, None, :]) / ell[d0:d1, None, None]
                K = s[d0:d1, None, None] * torch.exp(-0.5 * t * t) + JITTER * eye
                L = torch.linalg.cholesky(K)
                Li = torch.linalg.solve_triangular(L, eye.expand(d1 - d0, M, M), upper=False)
                linv[d0:d1] = Li.float()
                beta[d0:d1] = torch.einsum("dij,dj->di", Li, m_q[d0:d1].to(f64) - c[d0:d1, None]).float()
                del t, K, L, Li
        lq = torch.tril(chol_var).float().contiguous()
        zc = z.float().contiguous()
        args = [_capi.ptr(zc), _capi.ptr(linv), _capi.ptr(lq), _capi.ptr(beta), _capi.ptr(hyp)]
        if self.handle is None:
            dims = _capi.GpDims(D, M, JITTER, lb)
            hd = _capi.c_void_p()
            _capi.check(self.lib.dvg_gp_prepare_factors(_capi.ctypes.byref(hd), _capi.ctypes.byref(dims), *args,
                                                        _capi.stream_ptr()), "dvg_gp_prepare_factors")
            self.handle = hd
        else:
            _capi.check(self.lib.dvg_gp_refresh_factors(self.handle, *args, _capi.stream_ptr()),
                        "dvg_gp_refresh_factors")
        torch.cuda.current_stream().synchronize()      # the staging tensors die with this frame

    def __del__(self):
        try:
            if self.handle is not None:
                self.lib.dvg_gp_destroy(self.handle)
        except Exception:
            pass


class GPPrediction:
    """Lazy stand-in for the gpytorch ``MultivariateNormal`` the rollout consumes."""

    def __init__(self, layer, x, likelihood=None):
        if x.dim() != 3 or x.shape[-1] != 1 or x.shape[0] != layer.num_dims:
            raise ValueError(f"expected x of shape [{layer.num_dims}, N, 1], got {tuple(x.shape)}")
        if not x.is_cuda:
            raise _capi.DvgError("dvg_b200 GP is CUDA-only (no CPU fallback); got a CPU tensor")
        self._layer, self._lik = layer, likelihood
        lat = x.detach()[..., 0].transpose(0, 1)        # [N, D]; the original latent when x was its view
        if lat.dtype != torch.float32:
            lat = lat.float()
        if lat.stride(1) != 1 or (lat.shape[0] > 1 and lat.stride(0) < lat.shape[1]):
            lat = lat.contiguous()
        self._lat = lat
        self._mv = None

    def with_likelihood(self, likelihood):
        p = GPPrediction.__new__(GPPrediction)
        p._layer, p._lik, p._lat, p._mv = self._layer, likelihood, self._lat, None
        return p

    def _rt(self):
        return self._layer._runtime(self._lik)

    def _ld(self):
        return self._lat.stride(0) if self._lat.shape[0] > 1 else self._lat.shape[1]

    def _mean_var(self):
        if self._mv is None:
            rt = self._rt()
            N, D = self._lat.shape
            out = torch.empty(2, N, D, dtype=torch.float32, device=self._lat.device)
            _capi.check(rt.lib.dvg_gp_predict(rt.handle, N, _capi.ptr(self._lat), self._ld(), None,
                                              _capi.ptr(out[0]), D, _capi.ptr(out[1]), D, _capi.stream_ptr()),
                        "dvg_gp_predict")
            self._mv = out
        return self._mv

    @property
    def mean(self):
        """[D, N] (a transposed view of the [N, D] buffer the decoder wants)."""
        return self._mean_var()[0].transpose(0, 1)

    @property
    def variance(self):
        return self._mean_var()[1].transpose(0, 1)

    def rsample(self, eps=None):
        """mean + chol(Sigma_y) eps, eps ~ N(0,I) drawn as ``randn[D,N]`` unless injected.  Returns [D, N]."""
        rt = self._rt()
        N, D = self._lat.shape
        if eps is None:
            eps = torch.randn(D, N, device=self._lat.device, dtype=torch.float32)
        eps = eps.to(self._lat.device, torch.float32).reshape(1, D, N).contiguous()
        out = torch.empty(N, D, dtype=torch.float32, device=self._lat.device)
        _capi.check(rt.lib.dvg_gp_rsample(rt.handle, 1, N, _capi.ptr(self._lat), self._ld(), _capi.ptr(eps), None,
                                          _capi.ptr(out), D, _capi.stream_ptr()), "dvg_gp_rsample")
        return out.transpose(0, 1)


class GPRegressionLayer1(nn.Module):
    def __init__(self, num_dims=90, num_inducing_points=40):
        super().__init__()
        D, M = num_dims, num_inducing_points
        self.num_dims, self.num_inducing_points = D, M
        vs = _Holder()
        vs.inducing_points = nn.Parameter(torch.rand(D, M, 1))                       # models/gp_models.py:13
        vs.register_buffer("variational_params_initialized", torch.tensor(0))    # gpytorch: set on the first call
        vd = _Holder()
        vd.variational_mean = nn.Parameter(torch.zeros(D, M))
        vd.chol_variational_covar = nn.Parameter(torch.eye(M).repeat(D, 1, 1))
        vs.variational_distribution = vd
        self.variational_strategy = vs
        self.mean_module = _Holder()
        self.mean_module.constant = nn.Parameter(torch.zeros(D, 1))
        self.covar_module = _Holder()
        self.covar_module.raw_outputscale = nn.Parameter(torch.zeros(D))
        self.covar_module.base_kernel = _Holder()
        self.covar_module.base_kernel.raw_lengthscale = nn.Parameter(torch.zeros(D, 1, 1))

    def __getstate__(self):
        state = dict(self.__dict__)
        for k in ("_dvg_rts", "_dvg_vinit", "_dvg_kl_memo"):
            state.pop(k, None)
        return state

    def _runtime(self, likelihood) -> _GpRuntime:
        self._ensure_variational_init()
        rts = self.__dict__.setdefault("_dvg_rts", {})
        key = id(likelihood) if likelihood is not None else 0
        rt = rts.get(key)
        if rt is None:
            rt = _GpRuntime(self, likelihood)
            rts[key] = rt
        elif rt.sig != rt.signature(self, likelihood):
            rt.refresh(self, likelihood)
        return rt

    def _load_from_state_dict(self, *args, **kwargs):
        self.__dict__.pop("_dvg_vinit", None)          # the loaded flag decides again
        return super()._load_from_state_dict(*args, **kwargs)

    def _ensure_variational_init(self):
        """gpytorch's ``VariationalStrategy.__call__`` initialises the variational distribution from the prior on the
        first call of a layer whose ``variational_params_initialized`` buffer is still 0 (a freshly constructed layer;
        checkpoints carry 1).  The flag is read from the device once per (construction / load_state_dict)."""
        if self.__dict__.get("_dvg_vinit"):
            return
        vs = self.variational_strategy
        if int(vs.variational_params_initialized.item()) == 0:
            from . import gp_train
            gp_train.initialize_variational_dist(self)
        self.__dict__["_dvg_vinit"] = True

    def forward(self, x):
        self._ensure_variational_init()
        if self.training:
            from . import gp_train
            return gp_train.train_forward(self, x)
        return GPPrediction(self, x)
