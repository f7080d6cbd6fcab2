"""Self-consistency of the (unpinned) GP oracle: fp32 gpytorch-op-order vs fp64 direct,
analytic properties, trigger arithmetic.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import gp_ref, trigger_ref


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
