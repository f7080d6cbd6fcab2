"""Oracle, second opinion: an INDEPENDENT restatement of the GP stage that follows the class structure and the
un-hoisted control flow of gpytorch 0.3.x instead of the closed-form equations of ``oracle/gp_ref.py``.
**PARITY UNPINNED** (gpytorch is absent from the reference tree, this image, the wheelhouse and the pip cache --
searched again in round 2: ``find / -iname '*gpytorch*'`` and ``pip download gpytorch`` both come back empty).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Why a second restatement: ``gp_ref.predictive`` was written from the equations of SURVEY 8c; this file was written
from the published structure of the library code the reference actually runs (``models/gp_models.py:10-24`` builds
``ConstantMean`` + ``ScaleKernel(RBFKernel)`` + ``CholeskyVariationalDistribution`` + ``WhitenedVariationalStrategy``,
``train.py:102,112`` add ``GaussianLikelihood`` and ``VariationalELBO``): one small class per gpytorch class, each
method named after and commented with the gpytorch method it restates, evaluated in gpytorch's order -- the prior is
formed on the JOINT input ``cat([Z, x])`` by ``model.forward`` and then sliced, ``add_jitter`` is applied to the
inducing block only, ``cholesky_solve`` acts on ``cat([K_ZX, m - mu_Z])``, the predictive covariance is the sum of a
root term and a data term, and the likelihood adds its noise last.  ``tests/test_oracle_gp.py`` cross-checks the two
restatements (eval mode: mean, covariance, variance, rsample) -- a transcription slip in either shows up as a
mismatch -- and this file additionally restates the TRAINING branch (diagonal data covariance, KL divergence memo,
``VariationalELBO`` with ``combine_terms=True``; train.py:146-172, 200-248), which checks the drop-in's autograd path
(``dvg_b200/models/gp_train.py``).  What neither can prove is that both agree with gpytorch itself.

Version notes (gpytorch 0.3.0 .. 0.3.6): ``GaussianLikelihood`` has the ``GreaterThan(1e-4)`` noise constraint from
0.3.3 on (``noise_lower_bound`` = 0 reproduces 0.3.0-0.3.2); ``variational_log_probability`` was renamed
``expected_log_prob`` in 0.3.3 with the same arithmetic.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import gp_ref

JITTER = 1e-3            # LazyTensor.add_jitter(jitter_val=1e-3)


class ConstantMean:
    """gpytorch.means.ConstantMean(batch_size=D): ``forward(x) = constant.expand(x.shape[:-1])``."""

    def __init__(self, constant):            # [D,1]
        self.constant = constant

    def __call__(self, x):                   # x [D,n,1] -> [D,n]
        return self.constant.expand(x.shape[0], x.shape[1])


class RBFKernel:
    """gpytorch.kernels.RBFKernel(batch_size=D): ``x_ = x.div(lengthscale)``; squared distance by
    ``Kernel._covar_dist(square_dist=True)`` (mean-centred quadratic expansion, diagonal zero-filled when x1 is x2,
    clamp at 0); ``postprocess_rbf = dist.div(-2).exp()``."""

    def __init__(self, raw_lengthscale):     # [D,1,1]
        self.lengthscale = F.softplus(raw_lengthscale)

    @staticmethod
    def _sq_dist(x1, x2, x1_eq_x2):
        adjustment = x1.mean(-2, keepdim=True)
        x1 = x1 - adjustment
        x2 = x2 - adjustment
        x1_norm = x1.pow(2).sum(dim=-1, keepdim=True)
        x2_norm = x2.pow(2).sum(dim=-1, keepdim=True)
        x1_ = torch.cat([-2.0 * x1, x1_norm, torch.ones_like(x1_norm)], dim=-1)
        x2_ = torch.cat([x2, torch.ones_like(x2_norm), x2_norm], dim=-1)
        res = x1_.matmul(x2_.transpose(-2, -1))
        if x1_eq_x2:
            res = res - torch.diag_embed(res.diagonal(dim1=-2, dim2=-1))     # diagonal().fill_(0), out of place
        return res.clamp_min(0)

    def __call__(self, x1, x2):
        x1_eq_x2 = x1.shape == x2.shape and torch.equal(x1, x2)
        return self._sq_dist(x1.div(self.lengthscale), x2.div(self.lengthscale), x1_eq_x2).div(-2).exp()


class ScaleKernel:
    """gpytorch.kernels.ScaleKernel(base, batch_size=D): ``outputscale.view(D,1,1) * base(x1, x2)``."""

    def __init__(self, base, raw_outputscale):   # [D]
        self.base = base
        self.outputscale = F.softplus(raw_outputscale)

    def __call__(self, x1, x2):
        return self.base(x1, x2) * self.outputscale.reshape(-1, 1, 1)


class CholeskyVariationalDistribution:
    """``variational_distribution``: mean = variational_mean, covariance = CholLazyTensor(chol * tril-mask)."""

    def __init__(self, variational_mean, chol_variational_covar):
        self.mean = variational_mean                                    # [D,M]
        M = chol_variational_covar.shape[-1]
        lower_mask = torch.ones(M, M, dtype=chol_variational_covar.dtype).tril(0)
        self.root = chol_variational_covar.mul(lower_mask)              # [D,M,M]

    def covariance_matrix(self):
        return self.root @ self.root.transpose(-1, -2)

    def logdet(self):                                                   # CholLazyTensor.logdet
        return self.root.diagonal(dim1=-2, dim2=-1).pow(2).log().sum(-1)


class GPModel:
    """``GPRegressionLayer1`` (models/gp_models.py:10-24) + ``WhitenedVariationalStrategy`` built from a state_dict."""

    def __init__(self, gp_sd, dtype=torch.float64):
        c = lambda k: gp_sd[k].detach().to(dtype)
        self.dtype = dtype
        self.inducing_points = c(gp_ref.K_INDUCING)                     # [D,M,1]
        self.mean_module = ConstantMean(c(gp_ref.K_CONST))
        self.covar_module = ScaleKernel(RBFKernel(c(gp_ref.K_LSCALE)), c(gp_ref.K_OSCALE))
        self.q = CholeskyVariationalDistribution(c(gp_ref.K_VMEAN), c(gp_ref.K_VCHOL))
        self.memo = {}

    # models/gp_models.py:21-24
    def forward(self, x):
        return self.mean_module(x), (lambda a, b: self.covar_module(a, b))

    # WhitenedVariationalStrategy.forward, marginalising branch
    def __call__(self, x, training=False):
        x = x.to(self.dtype)
        Z = self.inducing_points
        num_induc = Z.shape[-2]
        full_inputs = torch.cat([Z, x], dim=-2)
        full_mean, kern = self.forward(full_inputs)
        test_mean = full_mean[..., num_induc:]
        induc_mean = full_mean[..., :num_induc]
        mean_diff = (self.q.mean - induc_mean).unsqueeze(-1)
        # lazily sliced covariance blocks: each block is evaluated by the kernel on the sliced inputs
        zi, xi = full_inputs[..., :num_induc, :], full_inputs[..., num_induc:, :]
        induc_induc_covar = kern(zi, zi) + JITTER * torch.eye(num_induc, dtype=self.dtype)        # .add_jitter()
        induc_data_covar = kern(zi, xi)
        chol = torch.linalg.cholesky(induc_induc_covar)                                            # CholLazyTensor
        eager_rhs = torch.cat([induc_data_covar, mean_diff], -1)
        solve = torch.cholesky_solve(eager_rhs, chol)
        predictive_mean = test_mean + (induc_data_covar.transpose(-1, -2) @ solve[..., -1:]).squeeze(-1)
        root = induc_data_covar.transpose(-1, -2) @ self.q.root                                    # RootLazyTensor
        if training:
            # inv_quad_logdet(cat([K_ZX, mean_diff]), reduce_inv_quad=False): column-wise quadratic forms
            inv_quad = (eager_rhs * solve).sum(-2)
            interp_data_data_var, mean_diff_inv_quad = inv_quad[..., :-1], inv_quad[..., -1]
            logdet = 2.0 * chol.diagonal(dim1=-2, dim2=-1).log().sum(-1)
            data_diag = kern(xi, xi).diagonal(dim1=-2, dim2=-1)
            data_var = (data_diag - interp_data_data_var).clamp(0, math.inf)                       # DiagLazyTensor
            variance = root.pow(2).sum(-1) + data_var
            self.memo = {"prior_covar": induc_induc_covar, "logdet_memo": -logdet,
                         "mean_diff_inv_quad_memo": mean_diff_inv_quad}
            return {"mean": predictive_mean, "variance": variance, "covar": None}
        neg_induc_data_data_covar = (induc_data_covar.transpose(-1, -2) * -1) @ solve[..., :-1]
        data_covariance = kern(xi, xi) + neg_induc_data_data_covar
        covar = root @ root.transpose(-1, -2) + data_covariance                                    # PsdSumLazyTensor
        return {"mean": predictive_mean, "variance": covar.diagonal(dim1=-2, dim2=-1), "covar": covar}

    # WhitenedVariationalStrategy.kl_divergence (after a training-mode call)
    def kl_divergence(self):
        M = self.inducing_points.shape[-2]
        covar_trace = (self.q.covariance_matrix() * self.memo["prior_covar"]).reshape(self.q.mean.shape[0], -1).sum(-1)
        return 0.5 * (self.memo["logdet_memo"] - self.q.logdet() + covar_trace + self.memo["mean_diff_inv_quad_memo"] - M)

    # VariationalStrategy.initialize_variational_dist of the whitened strategy (first call of a fresh layer)
    def initial_variational_params(self):
        Z = self.inducing_points
        _, kern = self.forward(Z)
        K = kern(Z, Z) + JITTER * torch.eye(Z.shape[-2], dtype=self.dtype)
        return self.mean_module(Z), torch.linalg.cholesky(torch.linalg.inv(K))      # mean, scale_tril of N(mu, K^-1)


class GaussianLikelihood:
    """gpytorch.likelihoods.GaussianLikelihood(batch_size=D)."""

    def __init__(self, lik_sd, dtype=torch.float64, noise_lower_bound=gp_ref.NOISE_LOWER_BOUND):
        self.noise = F.softplus(lik_sd[gp_ref.K_NOISE].detach().to(dtype)) + noise_lower_bound      # [D,1]

    def __call__(self, pred):                # likelihood(mvn): covariance + noise I
        out = dict(pred)
        out["variance"] = pred["variance"] + self.noise
        if pred.get("covar") is not None:
            out["covar"] = pred["covar"] + torch.diag_embed(self.noise.expand(pred["variance"].shape))
        return out

    def variational_log_probability(self, pred, target):     # (= expected_log_prob from 0.3.3 on); pred is q(f)
        mean, variance = pred["mean"], pred["variance"]
        res = -0.5 * ((target - mean) ** 2 + variance) / self.noise
        res = res + (-0.5 * self.noise.log() - 0.5 * math.log(2 * math.pi))
        return res.sum(-1)


def variational_elbo(model: GPModel, lik: GaussianLikelihood, x, target, num_data):
    """``VariationalELBO(likelihood, gp_layer, num_data, combine_terms=True)(gp_layer(x), target)`` as used by
    train.py:112,165,226 -> [D]."""
    pred = model(x, training=True)
    num_batch = pred["mean"].shape[-1]
    log_likelihood = lik.variational_log_probability(pred, target.to(model.dtype)).div(num_batch)
    kl = model.kl_divergence().div(num_data)
    return log_likelihood - kl, pred


def predictive(gp_sd, lik_sd, x, dtype=torch.float64, noise_lower_bound=gp_ref.NOISE_LOWER_BOUND):
    """``likelihood(gp_layer(x))`` in eval mode through the class-structured path; same return as gp_ref.predictive."""
    return GaussianLikelihood(lik_sd, dtype, noise_lower_bound)(GPModel(gp_sd, dtype)(x, training=False))
