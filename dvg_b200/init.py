"""Random initial parameters with the reference's semantics, for harness / bench use.

* LSTM: ``utils.init_weights`` (utils.py:304-311) touches only the Linear layers (N(0, 0.02), zero bias);
  ``nn.LSTMCell`` keeps the torch default U(+-1/sqrt(H)) (train.py:82-83).
* GP: ``GPRegressionLayer1.__init__`` values (models/gp_models.py:11-19) or a "trained-like" set
   whose variational mean is a smooth function of the inducing locations.
"""
from __future__ import annotations

import torch


def init_lstm_state_dict(g_in, g_out, hidden, n_layers, seed=1, gaussian=False):
    gen = torch.Generator().manual_seed(seed)
    sd = {"embed.weight": torch.randn(hidden, g_in, generator=gen) * 0.02, "embed.bias": torch.zeros(hidden)}
    k = 1.0 / hidden ** 0.5
    for l in range(n_layers):
        for name, shape in (("weight_ih", (4 * hidden, hidden)), ("weight_hh", (4 * hidden, hidden)),
                            ("bias_ih", (4 * hidden,)), ("bias_hh", (4 * hidden,))):
            sd[f"lstm.{l}.{name}"] = (torch.rand(*shape, generator=gen) * 2 - 1) * k
    heads = ("mu_net", "logvar_net") if gaussian else ("output.0",)
    for head in heads:
        sd[f"{head}.weight"] = torch.randn(g_out, hidden, generator=gen) * 0.02
        sd[f"{head}.bias"] = torch.zeros(g_out)
    return sd


def init_gp_state_dicts(D=90, M=40, seed=1, trained_like=True):
    g = torch.Generator().manual_seed(seed)
    vs, vd = "variational_strategy.", "variational_strategy.variational_distribution."
    if not trained_like:
        gp = {vs + "inducing_points": torch.rand(D, M, 1, generator=g), vd + "variational_mean": torch.zeros(D, M),
              vd + "chol_variational_covar": torch.eye(M).repeat(D, 1, 1),
              vs + "variational_params_initialized": torch.tensor(1), "mean_module.constant": torch.zeros(D, 1),
              "covar_module.raw_outputscale": torch.zeros(D),
              "covar_module.base_kernel.raw_lengthscale": torch.zeros(D, 1, 1)}
        return gp, {"noise_covar.raw_noise": torch.zeros(D, 1)}
    z = torch.rand(D, M, 1, generator=g) * 2 - 1
    phase = torch.rand(D, 1, generator=g) * 6.283
    gp = {vs + "inducing_points": z, vd + "variational_mean": 0.3 * torch.sin(3.0 * z[..., 0] + phase),
          vd + "chol_variational_covar": torch.tril(0.5 * torch.eye(M).repeat(D, 1, 1)
                                                    + 0.05 * torch.randn(D, M, M, generator=g)),
          vs + "variational_params_initialized": torch.tensor(1),
          "mean_module.constant": 0.1 * torch.randn(D, 1, generator=g),
          "covar_module.raw_outputscale": 0.5 * torch.randn(D, generator=g),
          "covar_module.base_kernel.raw_lengthscale": 0.5 * torch.randn(D, 1, 1, generator=g) - 1.0}
    return gp, {"noise_covar.raw_noise": torch.randn(D, 1, generator=g) - 2.0}
