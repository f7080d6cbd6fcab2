"""Frame metrics: oracle pinned to the reference's own finn_ssim / finn_psnr (golden vectors made by
executing the reference definitions), and the CUDA kernel against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "metrics_finn.pt")


def test_oracle_matches_reference_functions():
    for c in torch.load(GOLD, weights_only=False):
        a, b = c["a"].numpy(), c["b"].numpy()
        assert abs(metrics_ref.finn_ssim(a, b).mean() - c["ssim_mean"]) < 1e-12
        assert abs(metrics_ref.finn_psnr(a.astype(np.float64), b.astype(np.float64)) - c["psnr"]) < 1e-10


def test_best_of_n_rule():
    ssim = np.zeros((2, 3, 4))
    ssim[0, 1] = 0.9
    ssim[1, 2] = 0.5
    assert metrics_ref.best_of_n(ssim).tolist() == [1, 2]


@pytest.mark.gpu
def test_cuda_matches_golden():
    from dvg_b200.rollout import eval_seq_finn
    for c in torch.load(GOLD, weights_only=False):
        a, b = c["a"], c["b"]
        H = a.shape[0]
        ssim, psnr = eval_seq_finn(a.reshape(1, 1, 1, H, H).cuda(), b.reshape(1, 1, 1, 1, H, H).cuda())
        assert abs(ssim.item() - c["ssim_mean"]) < 1e-4, (H, ssim.item(), c["ssim_mean"])
        assert abs(psnr.item() - c["psnr"]) < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("C,H", [(1, 64), (3, 64), (3, 128)])
def test_cuda_eval_seq_vs_oracle(C, H):
    from dvg_b200 import shard
    from dvg_b200.rollout import eval_seq_finn
    T, S, B = 3, 4, 2
    g = torch.Generator().manual_seed(C * 100 + H)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, H), indexing="ij")
    base = 0.5 + 0.5 * torch.sin(6 * xx + 3 * yy)
    gt = (base.expand(T, B, C, H, H) + 0.02 * torch.randn(T, B, C, H, H, generator=g)).clamp(0, 1)
    noise = torch.linspace(0.02, 0.3, S).reshape(1, S, 1, 1, 1, 1)
    gen = (gt.unsqueeze(1) + noise * torch.randn(T, S, B, C, H, H, generator=g)).clamp(0, 1)
    ssim, psnr = eval_seq_finn(gt.cuda(), gen.cuda())
    want_ssim, want_psnr = np.zeros((S, B, T)), np.zeros((S, B, T))
    for s in range(S):
        a, b = metrics_ref.finn_eval_seq([gt[t].numpy() for t in range(T)], [gen[t, s].numpy() for t in range(T)])
        want_ssim[s], want_psnr[s] = a, b
    assert np.abs(ssim.cpu().numpy() - want_ssim).max() < 1e-4
    assert np.abs(psnr.cpu().numpy() - want_psnr).max() < 2e-3
    best = shard.select_best(ssim.mean(2), higher_is_better=True).cpu().numpy()
    assert best.tolist() == metrics_ref.best_of_n(np.transpose(want_ssim, (1, 0, 2))).tolist()


def test_skimage_restatement_properties():
    """utils.eval_seq restatement (UNPINNED: skimage absent): identical images -> SSIM 1, PSNR inf; the interior mean does
    not depend on the filter's border mode; known closed form for constant images."""
    g = np.random.default_rng(0)
    a = g.random((32, 32))
    assert abs(metrics_ref.skimage_ssim(a, a) - 1.0) < 1e-12
    # constant images x = p, y = q: S = (2pq + C1) / (p^2 + q^2 + C1) exactly (all variances are 0)
    p, q = 0.3, 0.7
    C1 = (0.01 * 2) ** 2
    s = metrics_ref.skimage_ssim(np.full((20, 20), p), np.full((20, 20), q))
    assert abs(s - (2 * p * q + C1) / (p * p + q * q + C1)) < 1e-12
    # explicit 'valid' window means == uniform_filter + crop
    b = np.clip(a + 0.1 * g.standard_normal((32, 32)), 0, 1)
    from numpy.lib.stride_tricks import sliding_window_view
    wa, wb = sliding_window_view(a, (7, 7)), sliding_window_view(b, (7, 7))
    ux, uy = wa.mean((2, 3)), wb.mean((2, 3))
    cn = 49 / 48
    vx, vy = cn * ((wa ** 2).mean((2, 3)) - ux ** 2), cn * ((wb ** 2).mean((2, 3)) - uy ** 2)
    vxy = cn * ((wa * wb).mean((2, 3)) - ux * uy)
    C2 = (0.03 * 2) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    assert abs(S.mean() - metrics_ref.skimage_ssim(a, b)) < 1e-12
    assert abs(metrics_ref.skimage_psnr(a, b) - 10 * np.log10(1.0 / np.mean((a - b) ** 2))) < 1e-12
    assert abs(metrics_ref.skimage_psnr(a - 0.5, b - 0.5) - 10 * np.log10(4.0 / np.mean((a - b) ** 2))) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("C,H,shift", [(1, 64, 0.0), (3, 64, 0.0), (3, 128, 0.0), (1, 64, -0.5)])
def test_cuda_eval_seq_skimage_vs_oracle(C, H, shift):
    """dvg_eval_seq (the metric make_gifs ranks by) vs the restatement, incl. the negative-ground-truth PSNR range."""
    from dvg_b200 import shard
    from dvg_b200.rollout import eval_seq
    T, S, B = 3, 4, 2
    g = torch.Generator().manual_seed(C * 100 + H)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, H), indexing="ij")
    base = 0.5 + 0.5 * torch.sin(6 * xx + 3 * yy)
    gt = (base.expand(T, B, C, H, H) + 0.02 * torch.randn(T, B, C, H, H, generator=g)).clamp(0, 1) + shift
    noise = torch.linspace(0.02, 0.3, S).reshape(1, S, 1, 1, 1, 1)
    gen = (gt.unsqueeze(1) + noise * torch.randn(T, S, B, C, H, H, generator=g)).clamp(shift, 1 + shift)
    ssim, psnr = eval_seq(gt.cuda(), gen.cuda())
    want_ssim, want_psnr = np.zeros((S, B, T)), np.zeros((S, B, T))
    for s in range(S):
        a, b = metrics_ref.eval_seq([gt[t].numpy() for t in range(T)], [gen[t, s].numpy() for t in range(T)])
        want_ssim[s], want_psnr[s] = a, b
    assert np.abs(ssim.cpu().numpy() - want_ssim).max() < 1e-4
    assert np.abs(psnr.cpu().numpy() - want_psnr).max() < 2e-3
    best = shard.select_best(ssim.mean(2), higher_is_better=True).cpu().numpy()
    assert best.tolist() == metrics_ref.best_of_n(np.transpose(want_ssim, (1, 0, 2))).tolist()
