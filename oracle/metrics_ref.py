"""Oracle: numpy/scipy restatement of the reference's self-contained frame metrics.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  PINNED: ``tests/golden/metrics_finn.pt`` is produced by
executing the reference's own ``finn_psnr`` / ``fspecial_gauss`` / ``finn_ssim`` definitions (extracted from
``/root/reference/utils.py`` at generation time, tests/golden/make_golden_metrics.py).

Restates ``utils.py``:
  :259-261  finn_psnr   10 log10(1 / mse)
  :270-273  fspecial_gauss(size, sigma)
  :275-301  finn_ssim   11x11 Gaussian (sigma 1.5) windows via 'valid' convolution, K1=.01, K2=.03, L=1, float64
  :237-256  finn_eval_seq  per (sequence, frame): channel-mean SSIM (NaN -> -1) and PSNR
and the selection of generate_frames.py:188-189,207 (best = argsort(mean over frames of ssim)[-1]).

(The skimage-based ``eval_seq`` of utils.py:220-234 cannot be restated against anything runnable here: skimage is
absent; only the finn variant is provided.)
"""
from __future__ import annotations

import math

import numpy as np
from scipy import signal


def finn_psnr(x, y):
    mse = ((x - y) ** 2).mean()
    return 10 * np.log(1 / mse) / np.log(10)


def fspecial_gauss(size, sigma):
    x, y = np.mgrid[-size // 2 + 1:size // 2 + 1, -size // 2 + 1:size // 2 + 1]
    g = np.exp(-((x ** 2 + y ** 2) / (2.0 * sigma ** 2)))
    return g / g.sum()


def finn_ssim(img1, img2):
    img1 = np.asarray(img1, dtype=np.float64)
    img2 = np.asarray(img2, dtype=np.float64)
    window = fspecial_gauss(11, 1.5)
    C1, C2 = (0.01 * 1) ** 2, (0.03 * 1) ** 2
    conv = lambda a: signal.fftconvolve(a, window, mode="valid")
    mu1, mu2 = conv(img1), conv(img2)
    s11 = conv(img1 * img1) - mu1 * mu1
    s22 = conv(img2 * img2) - mu2 * mu2
    s12 = conv(img1 * img2) - mu1 * mu2
    return ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s11 + s22 + C2))


def finn_eval_seq(gt, pred):
    """gt, pred: lists over T of arrays [B, C, H, W].  Returns (ssim [B,T], psnr [B,T])."""
    T, bs = len(gt), gt[0].shape[0]
    ssim, psnr = np.zeros((bs, T)), np.zeros((bs, T))
    for i in range(bs):
        for t in range(T):
            C = gt[t][i].shape[0]
            for c in range(C):
                res = finn_ssim(gt[t][i][c], pred[t][i][c]).mean()
                ssim[i, t] += -1 if math.isnan(res) else res
                psnr[i, t] += finn_psnr(np.asarray(gt[t][i][c], dtype=np.float64), np.asarray(pred[t][i][c], dtype=np.float64))
            ssim[i, t] /= C
            psnr[i, t] /= C
    return ssim, psnr


def best_of_n(ssim_BST):
    """generate_frames.py:188-189,207: per sequence, the sample with the highest frame-mean SSIM.
    ssim [B, S, T] -> int64 [B]."""
    return np.argsort(np.mean(ssim_BST, 2), axis=1)[:, -1]
