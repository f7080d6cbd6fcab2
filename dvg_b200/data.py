"""On-device synthetic data: bouncing-digit batches in the layout the rollout loops consume.

``moving_mnist_batch`` is ``MovingMNIST.__getitem__`` (data/moving_mnist.py:38-91) for a whole batch, run by two CUDA
kernels behind ``dvg_moving_mnist``; the result ``[T, B, 1, W, W]`` unbinds into the list of T frames that
``utils.normalize_data`` (utils.py:86-95) hands to train.py / generate_frames.py.  Randomness is an explicit stream of
32-bit integers (default: torch's device generator), consumed in the reference's np.random call order."""
from __future__ import annotations

from typing import Optional

import torch

from . import _capi

DIGIT = 32


def draws_per_seq(n_frames: int, n_digits: int = 2) -> int:
    return n_digits * (5 + 4 * n_frames)


def synthetic_digit_bank(n: int = 64, seed: int = 0, device="cpu") -> torch.Tensor:
    """Stand-in for the 32x32-scaled MNIST digits when the MNIST files are not available (no network): smooth random
    strokes in [0, 1], mostly-zero background like the real digits.  [n, 32, 32] fp32."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.arange(DIGIT, dtype=torch.float32), torch.arange(DIGIT, dtype=torch.float32), indexing="ij")
    bank = torch.zeros(n, DIGIT, DIGIT)
    for i in range(n):
        pts = torch.rand(5, 2, generator=g) * 20 + 6                    # polyline through 5 random points
        for a, b in zip(pts[:-1], pts[1:]):
            for s in torch.linspace(0, 1, 12):
                c = a + (b - a) * s
                bank[i] = torch.maximum(bank[i], torch.exp(-((yy - c[0]) ** 2 + (xx - c[1]) ** 2) / (2 * 1.3 ** 2)))
    bank[bank < 0.05] = 0
    return bank.clamp_(0, 1).to(device)


def moving_mnist_batch(bank: torch.Tensor, n_seq: int, n_frames: int, image_size: int = 64, n_digits: int = 2,
                       deterministic: bool = False, draws: Optional[torch.Tensor] = None,
                       generator: Optional[torch.Generator] = None, out: Optional[torch.Tensor] = None,
                       return_traj: bool = False):
    """frames [n_frames, n_seq, 1, W, W] fp32 on ``bank``'s device (asynchronous on the current stream).

    ``draws`` [n_seq, >= draws_per_seq(n_frames, n_digits)] int32 (raw bits); drawn on the device when omitted."""
    if not bank.is_cuda:
        raise _capi.DvgError("moving_mnist_batch runs on the GPU only (no CPU fallback); got a CPU digit bank")
    assert bank.dtype == torch.float32 and bank.dim() == 3 and bank.shape[1:] == (DIGIT, DIGIT) and bank.is_contiguous()
    lib = _capi.load()
    dev = bank.device
    K = draws_per_seq(n_frames, n_digits)
    if draws is None:
        draws = torch.randint(-2 ** 31, 2 ** 31 - 1, (n_seq, K), device=dev, dtype=torch.int64,
                              generator=generator).to(torch.int32)
    assert draws.is_cuda and draws.dtype == torch.int32 and draws.is_contiguous() and draws.shape[0] == n_seq
    traj = torch.empty(n_seq, n_digits, 1 + 2 * n_frames, dtype=torch.int32, device=dev)
    if out is None:
        out = torch.empty(n_frames, n_seq, 1, image_size, image_size, device=dev)
    assert out.is_contiguous() and out.shape == (n_frames, n_seq, 1, image_size, image_size) and out.dtype == torch.float32
    with torch.cuda.device(dev):
        _capi.check(lib.dvg_moving_mnist(n_seq, n_frames, image_size, n_digits, 1 if deterministic else 0, _capi.ptr(bank),
                                         bank.shape[0], _capi.ptr(draws), draws.shape[1], _capi.ptr(traj), _capi.ptr(out),
                                         _capi.stream_ptr()), "dvg_moving_mnist")
    return (out, traj) if return_traj else out
