"""The bouncing-digit oracle (oracle/moving_mnist_ref.py) against golden vectors made by executing the reference's own
MovingMNIST.__getitem__ with scripted np.random.randint (tests/golden/make_golden_mnist.py).  Bit-exact."""
import glob
import os

import numpy as np
import pytest

from oracle import moving_mnist_ref as ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "moving_mnist_*.npz")))


def test_fixtures_present():
    assert len(FILES) >= 4


@pytest.mark.parametrize("path", FILES, ids=os.path.basename)
def test_oracle_matches_reference_golden(path):
    z = np.load(path)
    T, W, nd, det = int(z["seq_len"]), int(z["image_size"]), int(z["num_digits"]), bool(z["deterministic"])
    frames, traj = ref.batch(z["bank"], z["words"], T, W, nd, det)
    want = z["x"]                                                      # [B, T, W, W, 1]
    assert frames.shape == (T, want.shape[0], 1, W, W)
    assert np.array_equal(frames, want.transpose(1, 0, 4, 2, 3))
    assert traj[:, :, 1:].min() >= 0 and traj[:, :, 1:].max() < W - 32
    assert z["words"].shape[1] == ref.draws_per_seq(T, nd)


def test_deterministic_mode_mirrors_velocity():
    bank = np.ones((1, 32, 32), dtype=np.float32)
    # idx=0, sx=30, sy=5, dx=+4 (8 - 4), dy=0 (4 - 4)
    words = np.array([0, 30, 5, 8, 4] + [0] * 40, dtype=np.uint32)
    _, traj = ref.sample(bank, words, 6, 64, 1, deterministic=True)
    xs = traj[0, 1::2]
    assert list(xs) == [30, 31, 27, 23, 19, 15]                         # 34 -> clamped to 31, velocity mirrored


def test_host_helpers_without_gpu():
    """draws_per_seq agrees with the oracle's worst case and the C ABI; the synthetic digit bank looks like digits
    (values in [0, 1], mostly background); the generator itself refuses CPU tensors (no CPU fallback)."""
    import torch
    from dvg_b200 import _capi
    from dvg_b200.data import draws_per_seq, moving_mnist_batch, synthetic_digit_bank
    lib = _capi.load()
    for T, n in [(1, 1), (15, 2), (105, 2), (20, 8)]:
        assert draws_per_seq(T, n) == ref.draws_per_seq(T, n) == lib.dvg_moving_mnist_draws(T, n)
    bank = synthetic_digit_bank(5, seed=3)
    assert bank.shape == (5, 32, 32) and bank.dtype == torch.float32
    assert bank.min().item() >= 0.0 and bank.max().item() <= 1.0
    assert (bank == 0).float().mean().item() > 0.5 and bank.sum(dim=(1, 2)).min().item() > 10.0
    assert torch.equal(bank, synthetic_digit_bank(5, seed=3))
    with pytest.raises(_capi.DvgError):
        moving_mnist_batch(bank, 2, 5)
