"""Self-consistency of the (unpinned) GP oracle: fp32 gpytorch-op-order vs fp64 direct, the two independent
restatements against each other (closed-form gp_ref vs class-structured gp_ref2), analytic properties of the
training branch (KL against torch.distributions, expected log-likelihood), the drop-in's autograd training path
against gp_ref2, trigger arithmetic.  CPU only."""
import math

import numpy as np
import pytest
import torch

from oracle import gp_ref, gp_ref2, trigger_ref


@pytest.mark.parametrize("trained", [False, True])
def test_fp32_vs_fp64(trained):
    gp, lik = gp_ref.random_gp_state_dicts(90, 40, seed=5, trained_like=trained)
    h = torch.tanh(torch.randn(50, 90, generator=torch.Generator().manual_seed(1)))
    x = gp_ref.latent_to_gp_input(h)
    p32 = gp_ref.predictive(gp, lik, x, torch.float32, "gpytorch")
    p64 = gp_ref.predictive(gp, lik, x, torch.float64, "direct")
    rel = lambda a, b: ((a.double() - b).norm() / b.norm().clamp_min(1e-30)).item()
    assert rel(p32["variance"], p64["variance"]) < 1e-4
    assert rel(p32["mean"], p64["mean"]) < 2e-3        # alpha = K_ZZ^-1 (m-c) is ill-conditioned in fp32
    assert rel(p32["covar"], p64["covar"]) < 1e-4
    # diag-only path agrees with the full covariance path
    pd = gp_ref.predictive(gp, lik, x, torch.float64, "direct", full_cov=False)
    torch.testing.assert_close(pd["variance"], p64["variance"], rtol=1e-9, atol=1e-12)


def test_init_values_and_bounds():
    gp, lik = gp_ref.random_gp_state_dicts(8, 40, seed=2)
    ell, s, c, noise = gp_ref.effective_hypers(gp, lik, torch.float64)
    assert torch.allclose(ell, torch.full_like(ell, np.log(2.0)))
    assert torch.allclose(noise, torch.full_like(noise, np.log(2.0) + 1e-4))
    h = torch.rand(11, 8)
    p = gp_ref.predictive(gp, lik, gp_ref.latent_to_gp_input(h), torch.float64, "direct")
    assert torch.all(p["mean"].abs() < 1e-12)               # m_q = 0, c = 0
    # L_q = I:  var = s + noise + k^T(I - K^-1)k  stays within [noise, s + noise + |k|^2]
    assert torch.all(p["variance"] > noise.reshape(-1, 1))
    ev = torch.linalg.eigvalsh(p["covar"])
    assert ev.min() > 0


def test_upper_triangle_of_chol_param_is_masked():
    gp, lik = gp_ref.random_gp_state_dicts(6, 12, seed=3, trained_like=True)
    h = torch.tanh(torch.randn(9, 6))
    x = gp_ref.latent_to_gp_input(h)
    a = gp_ref.predictive(gp, lik, x, torch.float64, "direct")
    gp2 = dict(gp)
    gp2[gp_ref.K_VCHOL] = torch.tril(gp[gp_ref.K_VCHOL])
    b = gp_ref.predictive(gp2, lik, x, torch.float64, "direct")
    torch.testing.assert_close(a["covar"], b["covar"])


def test_rsample_moments():
    gp, lik = gp_ref.random_gp_state_dicts(3, 10, seed=4, trained_like=True)
    h = torch.tanh(torch.randn(5, 3))
    p = gp_ref.predictive(gp, lik, gp_ref.latent_to_gp_input(h), torch.float64, "direct")
    g = torch.Generator().manual_seed(0)
    samp = torch.stack([gp_ref.rsample(p["mean"], p["covar"], torch.randn(3, 5, generator=g, dtype=torch.float64))
                        for _ in range(20000)])
    assert (samp.mean(0) - p["mean"]).abs().max() < 0.05
    assert (samp.var(0) - p["variance"]).abs().max() < 0.08


def test_trigger_arithmetic():
    rng = np.random.default_rng(0)
    var = rng.random((90, 50), dtype=np.float32)
    v = trigger_ref.trigger_value(var, 3)
    assert v.dtype == np.float32
    assert abs(float(v) - float(np.sqrt((var[:, 3].astype(np.float64) ** 2).sum()))) < 1e-5
    ctx = rng.random(12, dtype=np.float32)
    ctx2 = trigger_ref.slide(ctx, v)
    assert ctx2.shape == (12,) and ctx2[-1] == v and np.all(ctx2[:-1] == ctx[1:])
    thr = trigger_ref.threshold(ctx2)
    assert thr.dtype == np.float32
    assert abs(float(thr) - (ctx2.astype(np.float64).mean() + 2.01 * ctx2.astype(np.float64).std())) < 1e-5
    assert trigger_ref.decide(ctx2, v) == (v > thr)


# ---- second, independent restatement (gpytorch's class structure / un-hoisted op order) ----------------------------
@pytest.mark.parametrize("trained,smooth", [(False, False), (True, False), (True, True)])
@pytest.mark.parametrize("D,M,N", [(90, 40, 50), (7, 12, 5), (5, 40, 64)])
def test_two_restatements_agree(trained, smooth, D, M, N):
    gp, lik = gp_ref.random_gp_state_dicts(D, M, seed=11, trained_like=trained, smooth_mean=smooth)
    h = torch.tanh(torch.randn(N, D, generator=torch.Generator().manual_seed(2)))
    x = gp_ref.latent_to_gp_input(h)
    a = gp_ref.predictive(gp, lik, x, torch.float64, "gpytorch")
    b = gp_ref2.predictive(gp, lik, x, torch.float64)
    for k in ("mean", "variance", "covar"):
        torch.testing.assert_close(a[k], b[k], rtol=1e-9, atol=1e-11)
    eps = torch.randn(D, N, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    torch.testing.assert_close(gp_ref.rsample(a["mean"], a["covar"], eps), gp_ref.rsample(b["mean"], b["covar"], eps),
                               rtol=1e-8, atol=1e-10)
    # fp32 evaluation of both (what the reference itself runs) stays within the fp32 conditioning of the problem
    a32 = gp_ref.predictive(gp, lik, x, torch.float32, "gpytorch")
    b32 = gp_ref2.predictive(gp, lik, x, torch.float32)
    rel = lambda u, v: ((u.double() - v.double()).norm() / v.double().norm().clamp_min(1e-30)).item()
    assert rel(a32["variance"], b32["variance"]) < 1e-5
    assert rel(a32["variance"], b["variance"]) < 1e-4


def test_training_branch_analytic_properties():
    D, M, N = 6, 12, 9
    gp, lik = gp_ref.random_gp_state_dicts(D, M, seed=4, trained_like=True, smooth_mean=True)
    h = torch.tanh(torch.randn(N, D, generator=torch.Generator().manual_seed(5)))
    x = gp_ref.latent_to_gp_input(h)
    model = gp_ref2.GPModel(gp, torch.float64)
    ev = model(x, training=False)
    tr = model(x, training=True)
    # training mode keeps only the diagonal of the data covariance: same mean, same marginal variances
    torch.testing.assert_close(tr["mean"], ev["mean"], rtol=1e-10, atol=1e-12)
    torch.testing.assert_close(tr["variance"], ev["variance"], rtol=1e-8, atol=1e-10)
    # the whitened KL is KL( N(m, K V K) || N(c, K) ), V = L_q L_q^T, K = K_ZZ + jitter
    K = model.memo["prior_covar"]
    V = model.q.covariance_matrix()
    q = torch.distributions.MultivariateNormal(model.q.mean, covariance_matrix=K @ V @ K)
    p = torch.distributions.MultivariateNormal(model.mean_module(model.inducing_points), covariance_matrix=K)
    torch.testing.assert_close(model.kl_divergence(), torch.distributions.kl_divergence(q, p), rtol=1e-6, atol=1e-8)
    # E_q[log N(y | f, noise)] = log N(y | mean, noise) - var / (2 noise)
    likelihood = gp_ref2.GaussianLikelihood(lik, torch.float64)
    y = torch.randn(D, N, dtype=torch.float64, generator=torch.Generator().manual_seed(6))
    want = (torch.distributions.Normal(tr["mean"], likelihood.noise.sqrt()).log_prob(y) - tr["variance"] / (2 * likelihood.noise)).sum(-1)
    torch.testing.assert_close(likelihood.variational_log_probability(tr, y), want, rtol=1e-10, atol=1e-12)


def test_dropin_training_path_matches_second_restatement():
    """dvg_b200/models/gp_train.py (torch ops with autograd; what GPRegressionLayer1.forward does in train() mode)
    against the class-structured restatement: ELBO value in fp64, gradients by finite differences."""
    from dvg_b200.models.gp_models import GaussianLikelihood, GPRegressionLayer1
    from dvg_b200.models.gp_train import VariationalELBO
    D, M, N = 5, 10, 8
    gp, lik = gp_ref.random_gp_state_dicts(D, M, seed=9, trained_like=True, smooth_mean=True, dtype=torch.float64)
    layer, likelihood = GPRegressionLayer1(D, M).double(), GaussianLikelihood(D).double()
    layer.load_state_dict(gp)
    likelihood.load_state_dict(lik)
    layer.train(); likelihood.train()
    h = torch.tanh(torch.randn(N, D, dtype=torch.float64, generator=torch.Generator().manual_seed(1)))
    y = torch.tanh(torch.randn(N, D, dtype=torch.float64, generator=torch.Generator().manual_seed(2)))
    x = h.transpose(0, 1).reshape(D, N, 1)
    mll = VariationalELBO(likelihood, layer, num_data=N, combine_terms=True)

    def elbo_sum():
        return mll(layer(x), y.transpose(0, 1)).sum()

    got = elbo_sum()
    want, pred = gp_ref2.variational_elbo(gp_ref2.GPModel(layer.state_dict(), torch.float64),
                                          gp_ref2.GaussianLikelihood(likelihood.state_dict(), torch.float64), x, y.transpose(0, 1), N)
    torch.testing.assert_close(got, want.sum(), rtol=1e-7, atol=1e-9)     # direct vs quadratic-expansion distances
    p2 = layer(x)
    torch.testing.assert_close(p2.mean, pred["mean"], rtol=1e-7, atol=1e-9)
    torch.testing.assert_close(p2.variance, pred["variance"], rtol=1e-7, atol=1e-9)
    torch.testing.assert_close(likelihood(p2).variance, pred["variance"] + likelihood.noise, rtol=1e-7, atol=1e-9)
    got.backward()
    for prm in (layer.covar_module.base_kernel.raw_lengthscale, layer.variational_strategy.variational_distribution.variational_mean,
                layer.variational_strategy.inducing_points, likelihood.noise_covar.raw_noise):
        g = prm.grad.reshape(-1)
        flat = prm.data.reshape(-1)
        for idx in (0, flat.numel() // 2):
            old = flat[idx].item()
            flat[idx] = old + 1e-6
            up = elbo_sum().item()
            flat[idx] = old - 1e-6
            dn = elbo_sum().item()
            flat[idx] = old
            fd = (up - dn) / 2e-6
            assert abs(fd - g[idx].item()) <= 1e-5 * max(1.0, abs(fd)), (idx, fd, g[idx].item())


def test_fresh_layer_initialises_its_variational_distribution_like_gpytorch():
    """A freshly constructed layer (flag 0) takes m_q <- prior mean, L_q <- chol((K_ZZ + 1e-3 I)^-1) on its first
    call (VariationalStrategy.__call__ -> initialize_variational_dist); a loaded checkpoint (flag 1) is left alone."""
    from dvg_b200.models.gp_models import GPRegressionLayer1
    torch.manual_seed(0)
    layer = GPRegressionLayer1(4, 9).double().train()
    sd0 = {k: v.clone() for k, v in layer.state_dict().items()}
    assert int(sd0[gp_ref.K_VINIT]) == 0
    layer(torch.rand(4, 6, 1, dtype=torch.float64))
    sd1 = layer.state_dict()
    assert int(sd1[gp_ref.K_VINIT]) == 1
    mean, tril = gp_ref2.GPModel(sd0, torch.float64).initial_variational_params()
    torch.testing.assert_close(sd1[gp_ref.K_VMEAN], mean, rtol=1e-9, atol=1e-12)
    torch.testing.assert_close(sd1[gp_ref.K_VCHOL], tril, rtol=1e-7, atol=1e-9)
    layer2 = GPRegressionLayer1(4, 9).double().train()
    layer2.load_state_dict(sd1)
    layer2(torch.rand(4, 6, 1, dtype=torch.float64))
    torch.testing.assert_close(layer2.state_dict()[gp_ref.K_VCHOL], sd1[gp_ref.K_VCHOL])
