"""Shared helpers for the parity tests."""
import torch


def relerr(a, b):
    """max-norm relative error  max|a-b| / max|b|."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def make_lstm(sd, gaussian=False, rows=1, variant="fp32"):
    """dvg_b200 drop-in module on cuda:0 loaded with a reference-layout state_dict."""
    from dvg_b200.models.lstm import gaussian_lstm, lstm
    H, G = sd["embed.weight"].shape
    L = len([k for k in sd if k.endswith("weight_ih")])
    if gaussian:
        m = gaussian_lstm(G, sd["mu_net.weight"].shape[0], H, L, rows)
    else:
        m = lstm(G, sd["output.0.weight"].shape[0], H, L, rows)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    m.gemm_variant = variant
    return m


def make_gp(gp_sd, lik_sd):
    from dvg_b200.models.gp_models import GaussianLikelihood, GPRegressionLayer1
    D, M = gp_sd["variational_strategy.variational_distribution.variational_mean"].shape
    gp = GPRegressionLayer1(D, M)
    gp.load_state_dict(gp_sd)
    lik = GaussianLikelihood(D)
    lik.load_state_dict(lik_sd)
    return gp.cuda().eval(), lik.cuda().eval()


TOL = {"fp32": 2e-5, "bf16x3": 1e-4, "bf16": 2e-2}
