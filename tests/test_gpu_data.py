"""dvg_moving_mnist (data/moving_mnist.py:38-91 on the device) -- bit-exact against the reference-made golden vectors
and the oracle, properties at full size, argument errors."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import moving_mnist_ref as ref

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _words(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "moving_mnist_*.npz"))), ids=os.path.basename)
def test_golden(path):
    from dvg_b200.data import moving_mnist_batch
    z = np.load(path)
    T, W, nd, det = int(z["seq_len"]), int(z["image_size"]), int(z["num_digits"]), bool(z["deterministic"])
    B = z["x"].shape[0]
    bank = torch.from_numpy(z["bank"]).cuda()
    frames = moving_mnist_batch(bank, B, T, W, nd, det, draws=_words(z["words"]))
    assert frames.shape == (T, B, 1, W, W)
    assert np.array_equal(frames.cpu().numpy(), z["x"].transpose(1, 0, 4, 2, 3))


@pytest.mark.parametrize("B,T,W,nd,det", [(64, 30, 64, 2, False), (33, 105, 64, 2, False), (5, 9, 36, 4, False),
                                          (17, 25, 128, 2, True), (1, 1, 64, 1, False)])
def test_matches_oracle(B, T, W, nd, det):
    from dvg_b200.data import draws_per_seq, moving_mnist_batch, synthetic_digit_bank
    rng = np.random.RandomState(B * 7 + T)
    bank = synthetic_digit_bank(9, seed=T)
    words = rng.randint(0, 2 ** 32, size=(B, draws_per_seq(T, nd) + 3), dtype=np.uint64).astype(np.uint32)   # longer is fine
    want, want_traj = ref.batch(bank.numpy(), words, T, W, nd, det)
    got, traj = moving_mnist_batch(bank.cuda(), B, T, W, nd, det, draws=_words(words), return_traj=True)
    assert np.array_equal(traj.cpu().numpy(), want_traj)
    assert np.array_equal(got.cpu().numpy(), want)


def test_full_size_properties_and_device_draws():
    from dvg_b200.data import moving_mnist_batch, synthetic_digit_bank
    B, T, W = 1600, 15, 64
    bank = synthetic_digit_bank(32, seed=1).cuda()
    g = torch.Generator(device="cuda").manual_seed(5)
    frames, traj = moving_mnist_batch(bank, B, T, W, generator=g, return_traj=True)
    g.manual_seed(5)
    again = moving_mnist_batch(bank, B, T, W, generator=g)
    assert torch.equal(frames, again)                                   # same generator state -> same batch
    assert frames.min().item() >= 0.0 and frames.max().item() <= 1.0
    pos = traj[:, :, 1:]
    assert pos.min().item() >= 0 and pos.max().item() < W - 32
    assert traj[:, :, 0].min().item() >= 0 and traj[:, :, 0].max().item() < 32
    # per-frame mass never exceeds the mass of its digits
    mass = bank.sum(dim=(1, 2))[traj[:, :, 0].long()].sum(1)            # [B]
    fm = frames.sum(dim=(2, 3, 4))                                      # [T, B]
    assert (fm <= mass[None] * (1 + 1e-5) + 1e-3).all()
    # ... and equals it where the digits do not overlap (needs an arena wider than two digits)
    W2, B2 = 128, 200
    f2, tr2 = moving_mnist_batch(bank, B2, T, W2, generator=g, return_traj=True)
    pos2 = tr2[:, :, 1:]
    sx, sy = pos2[:, :, 0::2], pos2[:, :, 1::2]                         # [B, n, T]
    apart = ((sx[:, 0] - sx[:, 1]).abs() >= 32) | ((sy[:, 0] - sy[:, 1]).abs() >= 32)    # [B, T]
    assert apart.float().mean().item() > 0.3
    mass2 = bank.sum(dim=(1, 2))[tr2[:, :, 0].long()].sum(1)
    fm2 = f2.sum(dim=(2, 3, 4)).t()                                     # [B, T]
    assert torch.allclose(fm2[apart], mass2[:, None].expand(B2, T)[apart], rtol=1e-4, atol=1e-3)
    # digits move: consecutive frames differ for (almost) every sequence
    assert ((frames[1:] - frames[:-1]).abs().sum(dim=(2, 3, 4)) > 0).float().mean().item() > 0.9


def test_argument_errors():
    from dvg_b200 import _capi
    from dvg_b200.data import moving_mnist_batch, synthetic_digit_bank
    bank = synthetic_digit_bank(3)
    with pytest.raises(_capi.DvgError):
        moving_mnist_batch(bank, 2, 5)                                  # CPU bank: no CPU fallback
    bank = bank.cuda()
    with pytest.raises(_capi.DvgError):
        moving_mnist_batch(bank, 2, 5, image_size=32)                   # no room to move (randint(0) in the reference)
    with pytest.raises(_capi.DvgError):
        moving_mnist_batch(bank, 2, 5, draws=torch.zeros(2, 10, dtype=torch.int32, device="cuda"))   # stream too short
