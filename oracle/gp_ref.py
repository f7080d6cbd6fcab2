"""Oracle: plain-torch CPU restatement of the GP stage.  **PARITY UNPINNED.**

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

The reference's GP arithmetic is not in its tree: ``models/gp_models.py:10-24``
subclasses ``gpytorch.models.AbstractVariationalGP`` with a
``CholeskyVariationalDistribution`` + ``WhitenedVariationalStrategy`` and a
``ScaleKernel(RBFKernel)`` / ``ConstantMean`` prior, and the rollout wraps it in
``gpytorch.likelihoods.GaussianLikelihood`` (``generate_frames.py:67``,
``train.py:102``).  gpytorch is an un-vendored, un-pinned third-party dependency
(API used exists only in gpytorch >=0.3.0,<1.0) that is not installed and not
installable here, and the reference holds no tests or golden vectors for this
stage.  What follows restates the published eval-mode algorithm of
``WhitenedVariationalStrategy.forward`` (marginalising branch) from gpytorch
0.3.x, anchored on the reference call sites
``generate_frames.py:131,170,229,273,291`` and ``train.py:283``.

Per latent dimension d (all D in parallel; x = column d of the [N,D] latent,
i.e. the reference's ``h.transpose(0,1).view(D,N,1)``):

    ell = softplus(raw_lengthscale)   s = softplus(raw_outputscale)   c = constant
    noise = softplus(raw_noise) + noise_lower_bound (GreaterThan(1e-4); 0 in the earliest 0.3.x)
    k(a,b) = s * exp(-0.5 * ((a-b)/ell)^2)
    K_ZZ = k(Z,Z) + 1e-3 I          (add_jitter default)
    L_ZZ = chol(K_ZZ)
    L_q  = tril(chol_variational_covar)
    mean = c + K_XZ K_ZZ^-1 (m_q - c)
    Sigma_f = (K_XZ L_q)(K_XZ L_q)^T + K_XX - K_XZ K_ZZ^-1 K_ZX
    Sigma_y = Sigma_f + noise I     (likelihood(...))
    variance = diag(Sigma_y)
    rsample  = mean + chol(Sigma_y) eps,   eps ~ N(0, I_N)   (generate_frames.py:171,292)

``mode="gpytorch"`` forms squared distances like gpytorch's ``Kernel._sq_dist``
(mean-centred quadratic expansion, clamp at 0, exact zero diagonal when
x1 is x2); ``mode="direct"`` uses (a-b)^2.  ``dtype=torch.float64`` is the
"truth" bracket.
"""
from __future__ import annotations

import math
import torch
import torch.nn.functional as F

JITTER = 1e-3           # gpytorch LazyTensor.add_jitter default
NOISE_LOWER_BOUND = 1e-4  # GaussianLikelihood noise_constraint GreaterThan(1e-4)

K_INDUCING = "variational_strategy.inducing_points"
K_VMEAN = "variational_strategy.variational_distribution.variational_mean"
K_VCHOL = "variational_strategy.variational_distribution.chol_variational_covar"
K_VINIT = "variational_strategy.variational_params_initialized"
K_CONST = "mean_module.constant"
K_OSCALE = "covar_module.raw_outputscale"
K_LSCALE = "covar_module.base_kernel.raw_lengthscale"
K_NOISE = "noise_covar.raw_noise"


def random_gp_state_dicts(D=90, M=40, seed=1, trained_like=False, dtype=torch.float32, smooth_mean=False):
    """Parameter sets of SURVEY 8d.  ``trained_like=False`` reproduces
    ``GPRegressionLayer1.__init__`` (models/gp_models.py:11-19): Z ~ U(0,1),
    m_q = 0, L_q = I, c = 0, raw scales = 0 (softplus -> ln 2), raw_noise = 0.
    ``trained_like=True``: m_q ~ 0.3 N(0,1), L_q = tril(0.5 I + 0.05 N(0,1)),
    perturbed hyper-parameters, Z spread over the tanh range (-1,1).  The i.i.d. m_q is adversarial
    for the mean: K_ZZ^-1 (m_q - c) amplifies by ~cond(K_ZZ), so fp32 evaluations (the reference's own
    included) sit 1e-4..1e-2 from fp64.  ``smooth_mean=True`` uses m_q = 0.3 sin(3 z + phase_d), what a
    fitted GP looks like, for which fp32 is ~1e-5 accurate."""
    g = torch.Generator().manual_seed(seed)
    if not trained_like:
        gp = {
            K_INDUCING: torch.rand(D, M, 1, generator=g),
            K_VMEAN: torch.zeros(D, M),
            K_VCHOL: torch.eye(M).repeat(D, 1, 1),
            K_VINIT: torch.tensor(1),
            K_CONST: torch.zeros(D, 1),
            K_OSCALE: torch.zeros(D),
            K_LSCALE: torch.zeros(D, 1, 1),
        }
        lik = {K_NOISE: torch.zeros(D, 1)}
    else:
        gp = {
            K_INDUCING: torch.rand(D, M, 1, generator=g) * 2 - 1,
            K_VMEAN: 0.3 * torch.randn(D, M, generator=g),
            K_VCHOL: torch.tril(0.5 * torch.eye(M).repeat(D, 1, 1) + 0.05 * torch.randn(D, M, M, generator=g))
            + torch.triu(torch.randn(D, M, M, generator=g), 1),  # garbage above the diagonal: must be masked
            K_VINIT: torch.tensor(1),
            K_CONST: 0.1 * torch.randn(D, 1, generator=g),
            K_OSCALE: 0.5 * torch.randn(D, generator=g),
            K_LSCALE: 0.5 * torch.randn(D, 1, 1, generator=g) - 1.0,
        }
        lik = {K_NOISE: torch.randn(D, 1, generator=g) - 2.0}
        if smooth_mean:
            phase = torch.rand(D, 1, generator=g) * 6.283
            gp[K_VMEAN] = 0.3 * torch.sin(3.0 * gp[K_INDUCING][..., 0] + phase)
    cast = lambda sd: {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    return cast(gp), cast(lik)


def effective_hypers(gp_sd, lik_sd, dtype=torch.float32, noise_lower_bound=NOISE_LOWER_BOUND):
    """softplus transforms -> (ell [D], s [D], c [D], noise [D])."""
    ell = F.softplus(gp_sd[K_LSCALE].to(dtype)).reshape(-1)
    s = F.softplus(gp_sd[K_OSCALE].to(dtype)).reshape(-1)
    c = gp_sd[K_CONST].to(dtype).reshape(-1)
    noise = F.softplus(lik_sd[K_NOISE].to(dtype)).reshape(-1) + noise_lower_bound
    return ell, s, c, noise


def _sq_dist_gpytorch(x1, x2, x1_eq_x2):
    """gpytorch 0.3.x Kernel._sq_dist on [D,n,1] inputs."""
    adjustment = x1.mean(-2, keepdim=True)
    x1 = x1 - adjustment
    x2 = x2 - adjustment
    x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
    x1_pad = torch.ones_like(x1_norm)
    x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
    x2_pad = torch.ones_like(x2_norm)
    x1_ = torch.cat([-2.0 * x1, x1_norm, x1_pad], dim=-1)
    x2_ = torch.cat([x2, x2_pad, x2_norm], dim=-1)
    res = x1_.matmul(x2_.transpose(-2, -1))
    if x1_eq_x2:
        res.diagonal(dim1=-2, dim2=-1).fill_(0)
    return res.clamp_min_(0)


def kernel(x1, x2, ell, s, mode="gpytorch", x1_eq_x2=False):
    """ScaleKernel(RBFKernel) on [D,n1,1] x [D,n2,1] -> [D,n1,n2]."""
    ell_ = ell.reshape(-1, 1, 1)
    a, b = x1 / ell_, x2 / ell_
    if mode == "gpytorch":
        d2 = _sq_dist_gpytorch(a, b, x1_eq_x2)
    else:
        d2 = (a - b.transpose(-2, -1)).pow(2)
    return d2.div(-2).exp() * s.reshape(-1, 1, 1)


def predictive(gp_sd, lik_sd, x, dtype=torch.float32, mode="gpytorch",
               noise_lower_bound=NOISE_LOWER_BOUND, full_cov=True):
    """``likelihood(gp_layer(x))`` in eval mode.  x: [D,N,1] (any strides).

    Returns dict(mean [D,N], variance [D,N], covar [D,N,N] or None)."""
    x = x.to(dtype)
    ell, s, c, noise = effective_hypers(gp_sd, lik_sd, dtype, noise_lower_bound)
    Z = gp_sd[K_INDUCING].to(dtype)
    m_q = gp_sd[K_VMEAN].to(dtype)
    L_q = torch.tril(gp_sd[K_VCHOL].to(dtype))
    D, M = m_q.shape
    N = x.shape[-2]
    K_zz = kernel(Z, Z, ell, s, mode, x1_eq_x2=True) + JITTER * torch.eye(M, dtype=dtype)
    K_zx = kernel(Z, x, ell, s, mode)                      # [D,M,N]
    L_zz = torch.linalg.cholesky(K_zz)
    mean_diff = (m_q - c.reshape(-1, 1)).unsqueeze(-1)     # [D,M,1]
    rhs = torch.cat([K_zx, mean_diff], -1)
    solve = torch.cholesky_solve(rhs, L_zz)                # K_ZZ^-1 [K_ZX, m-c]
    K_xz = K_zx.transpose(-1, -2)
    mean = c.reshape(-1, 1) + (K_xz @ solve[..., -1:]).squeeze(-1)
    root = K_xz @ L_q                                      # [D,N,M]
    neg = (K_xz * -1) @ solve[..., :-1]                    # -K_XZ K_ZZ^-1 K_ZX
    if full_cov:
        K_xx = kernel(x, x, ell, s, mode, x1_eq_x2=True)
        covar = root @ root.transpose(-1, -2) + (K_xx + neg)
        covar = covar + torch.diag_embed(noise.reshape(-1, 1).expand(D, N))
        var = covar.diagonal(dim1=-2, dim2=-1)
    else:
        covar = None
        var = root.pow(2).sum(-1) + s.reshape(-1, 1) + neg.diagonal(dim1=-2, dim2=-1) + noise.reshape(-1, 1)
    return {"mean": mean, "variance": var, "covar": covar}


def rsample(mean, covar, eps):
    """MultivariateNormal.rsample for N <= max_cholesky_size: mean + chol(Sigma_y) eps.
    mean [D,N], covar [D,N,N], eps [D,N] (gpytorch draws randn[D,N,1])."""
    L = torch.linalg.cholesky(covar)
    return mean + (L @ eps.to(covar.dtype).unsqueeze(-1)).squeeze(-1)


def latent_to_gp_input(h):
    """``h.transpose(0,1).view(D,N,1)`` (generate_frames.py:131,170,229,273,291) for an
    [N,D] latent -- written with reshape so non-contiguous inputs work on CPU."""
    return h.transpose(0, 1).reshape(h.shape[1], h.shape[0], 1)
