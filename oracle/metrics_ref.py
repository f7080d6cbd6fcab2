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

``skimage_ssim`` / ``skimage_psnr`` / ``eval_seq`` restate utils.py:13-14,220-234, i.e. the legacy
``skimage.measure.compare_ssim`` / ``compare_psnr`` (skimage <= 0.17) with their defaults for float images, from the
library's documented behaviour.  PARITY UNPINNED for these three: skimage is not installed here,
not vendored and the reference has no fixtures for them.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import ndimage, signal


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


# ---- utils.eval_seq (legacy skimage defaults; UNPINNED, see the module docstring) ---------------------------------
def skimage_ssim(X, Y):
    """skimage.measure.compare_ssim(X, Y) of skimage <= 0.17 with default arguments on float arrays: win_size 7,
    uniform filter, K1=.01, K2=.03, use_sample_covariance=True, data_range = dtype range of float = 2, float64
    arithmetic, mean over the image cropped by (win_size - 1) // 2."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    win, K1, K2, R = 7, 0.01, 0.03, 2.0
    NP = win ** X.ndim
    cov_norm = NP / (NP - 1)
    filt = lambda a: ndimage.uniform_filter(a, size=win)
    ux, uy = filt(X), filt(Y)
    uxx, uyy, uxy = filt(X * X), filt(Y * Y), filt(X * Y)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (K1 * R) ** 2, (K2 * R) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win - 1) // 2
    return S[pad:-pad, pad:-pad].mean()


def skimage_psnr(im_true, im_test):
    """skimage.measure.compare_psnr(im_true, im_test) with data_range=None on float arrays: the dtype range of float
    is (-1, 1); data_range = 1 when min(im_true) >= 0 else 2."""
    im_true = np.asarray(im_true, dtype=np.float64)
    im_test = np.asarray(im_test, dtype=np.float64)
    R = 1.0 if im_true.min() >= 0 else 2.0
    err = np.mean((im_true - im_test) ** 2)
    return 10 * np.log10(R * R / err)


def eval_seq(gt, pred):
    """utils.py:220-234.  gt, pred: lists over T of arrays [B, C, H, W].  Returns (ssim [B,T], psnr [B,T])."""
    T, bs = len(gt), gt[0].shape[0]
    ssim, psnr = np.zeros((bs, T)), np.zeros((bs, T))
    for i in range(bs):
        for t in range(T):
            C = gt[t][i].shape[0]
            for c in range(C):
                ssim[i, t] += skimage_ssim(gt[t][i][c], pred[t][i][c])
                psnr[i, t] += skimage_psnr(gt[t][i][c], pred[t][i][c])
            ssim[i, t] /= C
            psnr[i, t] /= C
    return ssim, psnr
