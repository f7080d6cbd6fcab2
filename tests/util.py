"""Shared helpers for the parity tests."""
import torch


def relerr(a, b):
    """max-norm relative error  max|a-b| / max|b|."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def elemerr(a, b, rtol=1e-4, atol=None):
    """Element-wise error  max_i |a_i - b_i| / (atol + rtol |b_i|)  (<= 1 passes).  ``atol`` defaults to
    rtol * rms(b): an element may be off by rtol of its own magnitude plus rtol of the tensor's typical magnitude
    (hidden states and latents live in (-1, 1); exact zeros and sign changes make a pure relative test meaningless)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    if atol is None:
        atol = rtol * b.pow(2).mean().sqrt().clamp_min(1e-30).item()
    return ((a - b).abs() / (atol + rtol * b.abs())).max().item()


def sd_checksum(sd):
    """SHA-256 over a state_dict (tests/golden/make_golden.py: pins seeded weights that are not stored)."""
    import hashlib
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def big_golden_weights(g):
    """Regenerate the weights of a ``big_*`` golden from its seed and check them against the stored hash."""
    from oracle import lstm_ref
    gi, go, H, L, R = g["dims"]
    gauss = g["kind"] == "gauss_big"
    sd = lstm_ref.random_lstm_state_dict(gi, go, H, L, seed=g["weights_seed"], gaussian=gauss)
    if gauss:
        for k, sdd in (("embed.bias", 7), ("mu_net.bias", 8), ("logvar_net.bias", 9)):
            sd[k].uniform_(-0.2, 0.2, generator=torch.Generator().manual_seed(sdd))
    else:
        sd["embed.bias"].uniform_(-0.1, 0.1, generator=torch.Generator().manual_seed(5))
        sd["output.0.bias"].uniform_(-0.1, 0.1, generator=torch.Generator().manual_seed(6))
    assert sd_checksum(sd) == g["sha256"], "seeded weights differ from the ones the golden was made with"
    gen = torch.Generator().manual_seed(g["x_seed"])
    xs = [torch.tanh(torch.randn(R, gi, generator=gen)) for _ in range(g["steps"])]
    return sd, xs


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


def crafted_trigger_case(G, M, B, S, T, W, seed=0, n_jumps=6):
    """GP whose inducing points cover only (0.5, 1): with L_q ~ 0.5 I the |L_q^T k|^2 term makes the
    predictive variance several times larger inside that range than far from it.  Base latents live near
    -0.8 (low variance); at a few (t, s) pairs after the warm-up the whole rollout jumps into (0.55, 0.95),
    which makes the variance statistic jump up and the trigger fire with a wide margin.  Returns (gp_sd, lik_sd, lat [T,S*B,G], eps [T,S,G,B], jumps)."""
    from oracle import gp_ref
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(G, M, seed=seed + 50, trained_like=True, smooth_mean=True)
    gp_sd[gp_ref.K_INDUCING] = gp_sd[gp_ref.K_INDUCING].abs() * 0.5 + 0.5
    g = torch.Generator().manual_seed(seed)
    lat = -0.8 + 0.1 * torch.tanh(torch.randn(T, S * B, G, generator=g))
    jumps = set()
    while len(jumps) < n_jumps:
        t = int(torch.randint(W, T, (1,), generator=g))
        s = int(torch.randint(0, S, (1,), generator=g))
        if all(abs(t - t2) > 1 or s != s2 for t2, s2 in jumps):
            jumps.add((t, s))
    for t, s in jumps:
        lat[t, s * B:(s + 1) * B] = 0.75 + 0.2 * torch.tanh(torch.randn(B, G, generator=g))
    eps = torch.randn(T, S, G, B, generator=g)
    return gp_sd, lik_sd, lat, eps, jumps


def check_latent_rollout(sd, gp_sd, lik_sd, lat, eps, out, masks, values, B, W, stat_col=3, tol=1e-4):
    """Replay a latent-space trigger rollout on the CPU oracle, one rollout at a time (the reference's
    sequential loop, generate_frames.py:249-300), following the device's decisions where the statistic is
    within tolerance of the threshold and asserting them elsewhere.  Returns (#checked decisions, #fired)."""
    import numpy as np
    from oracle import gp_ref, lstm_ref, trigger_ref
    T = lat.shape[0]
    S = lat.shape[1] // B
    L = len([k for k in sd if k.endswith("weight_ih")])
    H = sd["embed.weight"].shape[0]
    checked = fired_n = 0
    for s in range(S):
        hid = lstm_ref.init_hidden(L, B, H)
        ctx = []
        for t in range(T):
            h = lat[t, s * B:(s + 1) * B]
            pred = gp_ref.predictive(gp_sd, lik_sd, gp_ref.latent_to_gp_input(h), torch.float32, "gpytorch",
                                     full_cov=False)
            v = trigger_ref.trigger_value(pred["variance"].numpy(), stat_col)
            assert abs(values[t, s].item() - float(v)) <= 1e-4 * abs(float(v)), (t, s)
            fired = False
            if t < W:
                ctx.append(v)
                assert not bool(masks[t, s])
            else:
                c = trigger_ref.slide(np.array(ctx, dtype=np.float32), v)
                ctx = list(c)
                thr = trigger_ref.threshold(c)
                if abs(float(v) - float(thr)) > 1e-4 * abs(float(thr)):
                    assert bool(masks[t, s]) == bool(v > thr), (t, s, float(v), float(thr))
                    checked += 1
                fired = bool(masks[t, s])
                fired_n += int(fired)
            if fired:   # rsample of the encoder latent; LSTM state NOT advanced (generate_frames.py:289-292)
                pc = gp_ref.predictive(gp_sd, lik_sd, gp_ref.latent_to_gp_input(h), torch.float64, "direct")
                want = gp_ref.rsample(pc["mean"], pc["covar"], eps[t, s].double()).transpose(0, 1).float()
            else:
                want, hid = lstm_ref.lstm_forward(sd, h, hid)
            assert relerr(out[t, s * B:(s + 1) * B], want) < tol, (t, s, fired)
    return checked, fired_n
