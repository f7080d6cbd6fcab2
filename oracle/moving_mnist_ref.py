"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU restatement of the reference's bouncing-digit generator,
data/moving_mnist.py:38-91 (``MovingMNIST.__getitem__``), with the lazily drawn np.random integers replaced by an
injected stream: the k-th ``np.random.randint(lo, hi)`` call of a sample is ``lo + draws[k] % (hi - lo)``.

PINNED: ``tests/golden/moving_mnist_*.npz`` were produced by executing the reference's own ``__getitem__`` with
``np.random.randint`` scripted to that rule (tests/golden/make_golden_mnist.py)."""
import numpy as np

DIGIT = 32


class _Stream:
    def __init__(self, words):
        self.w, self.k = words, 0

    def randint(self, lo, hi=None):
        if hi is None:
            lo, hi = 0, lo
        v = lo + int(self.w[self.k]) % (hi - lo)
        self.k += 1
        return v


def draws_per_seq(n_frames, n_digits):
    return n_digits * (5 + 4 * n_frames)


def sample(bank, words, seq_len, image_size=64, num_digits=2, deterministic=False):
    """One ``__getitem__``: returns (x [seq_len, W, W, 1] float32, traj [num_digits, 1 + 2*seq_len] int32)."""
    r = _Stream(words)
    W = image_size
    x = np.zeros((seq_len, W, W, 1), dtype=np.float32)                  # :41-45
    traj = np.zeros((num_digits, 1 + 2 * seq_len), dtype=np.int32)
    for n in range(num_digits):                                         # :46
        idx = r.randint(len(bank))                                      # :47
        digit = bank[idx]
        sx = r.randint(W - DIGIT)                                       # :50-53
        sy = r.randint(W - DIGIT)
        dx = r.randint(-4, 5)
        dy = r.randint(-4, 5)
        traj[n, 0] = idx
        for t in range(seq_len):                                        # :54
            if sy < 0:                                                  # :55-61
                sy = 0
                if deterministic:
                    dy = -dy
                else:
                    dy = r.randint(1, 5)
                    dx = r.randint(-4, 5)
            elif sy >= W - 32:                                          # :62-68
                sy = W - 32 - 1
                if deterministic:
                    dy = -dy
                else:
                    dy = r.randint(-4, 0)
                    dx = r.randint(-4, 5)
            if sx < 0:                                                  # :70-76
                sx = 0
                if deterministic:
                    dx = -dx
                else:
                    dx = r.randint(1, 5)
                    dy = r.randint(-4, 5)
            elif sx >= W - 32:                                          # :77-83
                sx = W - 32 - 1
                if deterministic:
                    dx = -dx
                else:
                    dx = r.randint(-4, 0)
                    dy = r.randint(-4, 5)
            traj[n, 1 + 2 * t], traj[n, 2 + 2 * t] = sx, sy
            x[t, sy:sy + 32, sx:sx + 32, 0] += digit                    # :85
            sy += dy                                                    # :86-87
            sx += dx
    x[x > 1] = 1.0                                                      # :89
    return x, traj


def batch(bank, draws, seq_len, image_size=64, num_digits=2, deterministic=False):
    """B samples + the loader/normalize_data layout (utils.py:86-95): frames [T, B, 1, W, W], traj [B, n, 1+2T]."""
    xs, trs = zip(*[sample(bank, draws[b], seq_len, image_size, num_digits, deterministic) for b in range(len(draws))])
    x = np.stack(xs)                                                    # [B, T, W, W, 1]
    return np.ascontiguousarray(x.transpose(1, 0, 4, 2, 3)), np.stack(trs)
