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
