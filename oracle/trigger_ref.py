"""Oracle: numpy restatement of the GP-variance trigger.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Restates ``generate_frames.py``:
  :230/:275  value = np.linalg.norm(variance.cpu().numpy().transpose(), axis=1)[col]
  :231       context_array = np.concatenate([context_array[1:], [value]])
  :288       threshold = np.mean(ctx) + (2 + 0.01*depth) * np.std(ctx)      (depth == 1 always, :254)
  :289       if value > threshold: resample (LSTM state not advanced), else LSTM step

All arithmetic is float32 (variance is a float32 tensor; under numpy>=2 the
python-float factor 2.01 is a weak scalar so the threshold stays float32).
"""
from __future__ import annotations

import numpy as np

FACTOR = 2 + 0.01 * 1   # generate_frames.py:288 with depth == 1 (:254, never updated)


def trigger_value(variance_DN: np.ndarray, col: int) -> np.float32:
    """variance [D,N] float32 -> ||variance[:, col]||_2 (generate_frames.py:230,275)."""
    v = np.asarray(variance_DN, dtype=np.float32)
    return np.linalg.norm(v.transpose(), axis=1)[col]


def slide(context: np.ndarray, value) -> np.ndarray:
    """generate_frames.py:231."""
    return np.concatenate([context[1:], [value]]).astype(np.float32)


def threshold(context: np.ndarray) -> np.float32:
    """generate_frames.py:288 (population std, ddof=0)."""
    ctx = np.asarray(context, dtype=np.float32)
    return np.float32(np.mean(ctx) + np.float32(FACTOR) * np.std(ctx))


def decide(context: np.ndarray, value) -> bool:
    """generate_frames.py:289 -- strict '>' against the window that already
    contains ``value`` (var_value slides before the threshold is formed, :287-288)."""
    return bool(np.float32(value) > threshold(context))
