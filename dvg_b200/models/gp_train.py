"""Training-mode GP for the drop-in ``GPRegressionLayer1`` (minimal form): the branch gpytorch 0.3.x's
``WhitenedVariationalStrategy.forward`` takes when ``self.training`` is set, as plain torch ops WITH autograd on the
GPU, so the reference's training steps (train.py:146-172 ``train_GP_Frame_predictor``, :200-248 ``train_model``) run
against the drop-in classes.  No hand-written kernels here: backward passes are out of the hot path's scope; the
sm_100a kernels own the eval-mode rollout.

What training mode computes (per latent dim d, all D batched):

    K_ZZ = k(Z,Z) + 1e-3 I ;  L = chol(K_ZZ) ;  A = L^-1 K_ZX ;  b = L^-1 (m_q - c)
    mean      = c + A^T b
    variance  = sum_m (K_XZ L_q)^2 + clamp(s - sum_m A^2, 0)          (DIAGONAL data covariance only)
    KL        = 0.5 [ -log|K_ZZ| - log|L_q L_q^T| + tr(L_q L_q^T K_ZZ) + |b|^2 - M ]      (memoised for the ELBO)
    ELBO      = E_q[log N(y | f, noise)] / N  -  KL / num_data          (VariationalELBO, combine_terms=True)

``gpytorch_shim()`` provides the few ``gpytorch.*`` names train.py touches (likelihoods.GaussianLikelihood,
mlls.VariationalELBO, settings.max_cg_iterations / use_toeplitz) for hosts without gpytorch.
"""
from __future__ import annotations

import contextlib
import math
import types

import torch
import torch.nn.functional as F

JITTER = 1e-3


def _rbf(x1, x2, ell, s):
    """ScaleKernel(RBFKernel) on [D,n1,1] x [D,n2,1] -> [D,n1,n2] (direct squared distance: differentiable and within
    1e-6 of gpytorch's quadratic-expansion form)."""
    d = (x1 - x2.transpose(-1, -2)) / ell
    return s * torch.exp(-0.5 * d * d)


class GPTrainPrediction:
    """q(f) at the training inputs: ``mean`` / ``variance`` [D,N] with autograd; ``covariance_matrix`` is diagonal in
    training mode (gpytorch builds a DiagLazyTensor data term there)."""

    def __init__(self, layer, mean, variance, kl):
        self._layer, self.mean, self.variance, self.kl = layer, mean, variance, kl

    @property
    def stddev(self):
        return self.variance.sqrt()

    def with_likelihood(self, likelihood):
        return GPTrainPrediction(self._layer, self.mean, self.variance + likelihood.noise, self.kl)

    def rsample(self, eps=None):
        eps = torch.randn_like(self.mean) if eps is None else eps.to(self.mean)
        return self.mean + self.variance.sqrt() * eps


def effective(layer):
    vs = layer.variational_strategy
    D = layer.num_dims
    ell = F.softplus(layer.covar_module.base_kernel.raw_lengthscale).reshape(D, 1, 1)
    s = F.softplus(layer.covar_module.raw_outputscale).reshape(D, 1, 1)
    c = layer.mean_module.constant.reshape(D, 1)
    return vs.inducing_points, vs.variational_distribution.variational_mean, \
        torch.tril(vs.variational_distribution.chol_variational_covar), ell, s, c


def train_forward(layer, x):
    """``gp_layer(x)`` in training mode.  x [D,N,1] (any strides; gradients flow to x as well)."""
    Z, m_q, L_q, ell, s, c = effective(layer)
    D, M = m_q.shape
    x = x.to(Z.dtype)
    K_zz = _rbf(Z, Z, ell, s) + JITTER * torch.eye(M, dtype=Z.dtype, device=Z.device)
    K_zx = _rbf(Z, x, ell, s)                                       # [D,M,N]
    L = torch.linalg.cholesky(K_zz)
    A = torch.linalg.solve_triangular(L, K_zx, upper=False)         # L^-1 K_ZX
    b = torch.linalg.solve_triangular(L, (m_q - c).unsqueeze(-1), upper=False)      # L^-1 (m - c)
    mean = c + (A.transpose(-1, -2) @ b).squeeze(-1)
    root = K_zx.transpose(-1, -2) @ L_q                             # K_XZ L_q
    variance = root.pow(2).sum(-1) + (s.reshape(D, 1) - A.pow(2).sum(-2)).clamp_min(0)
    logdet_K = 2.0 * L.diagonal(dim1=-2, dim2=-1).log().sum(-1)
    logdet_V = L_q.diagonal(dim1=-2, dim2=-1).pow(2).log().sum(-1)
    covar_trace = ((L_q @ L_q.transpose(-1, -2)) * K_zz).reshape(D, -1).sum(-1)
    kl = 0.5 * (-logdet_K - logdet_V + covar_trace + b.pow(2).sum((-1, -2)) - M)
    layer.__dict__["_dvg_kl_memo"] = kl
    return GPTrainPrediction(layer, mean, variance, kl)


@torch.no_grad()
def initialize_variational_dist(layer):
    """First call of a freshly constructed layer (VariationalStrategy.__call__ ->
    WhitenedVariationalStrategy.initialize_variational_dist): m_q <- prior mean at Z, L_q <- chol((K_ZZ + 1e-3 I)^-1)."""
    Z, m_q, _, ell, s, c = effective(layer)
    M = m_q.shape[1]
    K = _rbf(Z.double(), Z.double(), ell.double(), s.double()) + JITTER * torch.eye(M, dtype=torch.float64, device=Z.device)
    vd = layer.variational_strategy.variational_distribution
    vd.variational_mean.copy_(c.expand_as(m_q))
    vd.chol_variational_covar.copy_(torch.linalg.cholesky(torch.linalg.inv(K)).to(m_q.dtype))
    layer.variational_strategy.variational_params_initialized.fill_(1)


class VariationalELBO:
    """``gpytorch.mlls.VariationalELBO(likelihood, model, num_data, combine_terms=True)`` (train.py:112):
    ``mll(gp_layer(x), target)`` -> [D]."""

    def __init__(self, likelihood, model, num_data, combine_terms=True):
        self.likelihood, self.model, self.num_data, self.combine_terms = likelihood, model, num_data, combine_terms

    def __call__(self, pred: GPTrainPrediction, target):
        noise = self.likelihood.noise                                # [D,1]
        res = -0.5 * ((target - pred.mean) ** 2 + pred.variance) / noise
        res = res + (-0.5 * noise.log() - 0.5 * math.log(2 * math.pi))
        log_likelihood = res.sum(-1).div(pred.mean.shape[-1])
        kl = pred.kl.div(self.num_data)
        if self.combine_terms:
            return log_likelihood - kl
        return log_likelihood, kl, torch.zeros_like(kl)


def gpytorch_shim():
    """A stand-in ``gpytorch`` module exposing exactly what train.py / generate_frames.py touch."""
    from . import gp_models
    g = types.ModuleType("gpytorch")
    g.likelihoods = types.ModuleType("gpytorch.likelihoods")
    g.likelihoods.GaussianLikelihood = gp_models.GaussianLikelihood
    g.mlls = types.ModuleType("gpytorch.mlls")
    g.mlls.VariationalELBO = VariationalELBO
    g.settings = types.ModuleType("gpytorch.settings")

    @contextlib.contextmanager
    def _noop(*a, **k):
        yield
    g.settings.max_cg_iterations = _noop
    g.settings.use_toeplitz = _noop
    g.__dvg_shim__ = True
    return g
