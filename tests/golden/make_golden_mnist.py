"""Golden vectors for the bouncing-digit generator, produced by executing the REFERENCE's own
``MovingMNIST.__getitem__`` (/root/reference/data/moving_mnist.py:38-91) with ``np.random.randint`` scripted to the
injected-stream rule ``lo + words[k] % (hi - lo)``.  torchvision (imported at module level by the reference for the
MNIST download) is absent here and irrelevant to ``__getitem__``: a stub module stands in, the instance is created
without ``__init__`` and given an in-memory digit bank.

    python tests/golden/make_golden_mnist.py      # needs /root/reference; writes tests/golden/moving_mnist_*.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/data/moving_mnist.py"

CASES = [
    # name, B, seq_len, image_size, num_digits, deterministic, bank size, seed
    ("nondet_w64", 6, 15, 64, 2, False, 7, 1),
    ("det_w64", 4, 20, 64, 2, True, 5, 2),
    ("nondet_w48_3digits", 3, 40, 48, 3, False, 4, 3),     # small arena: many bounces
    ("nondet_w128_1digit", 2, 12, 128, 1, False, 3, 4),
]


def load_reference():
    tv = types.ModuleType("torchvision")
    tv.datasets = types.ModuleType("torchvision.datasets")
    tv.transforms = types.ModuleType("torchvision.transforms")
    saved = {k: sys.modules.get(k) for k in ("torchvision", "torchvision.datasets", "torchvision.transforms")}
    sys.modules.update({"torchvision": tv, "torchvision.datasets": tv.datasets, "torchvision.transforms": tv.transforms})
    try:
        spec = importlib.util.spec_from_file_location("ref_moving_mnist", REF)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def run_reference(mod, bank, words, seq_len, image_size, num_digits, deterministic):
    ds = mod.MovingMNIST.__new__(mod.MovingMNIST)
    ds.seq_len, ds.num_digits, ds.image_size = seq_len, num_digits, image_size
    ds.step_length, ds.digit_size, ds.deterministic, ds.seed_is_set, ds.channels = 0.1, 32, deterministic, False, 1
    ds.data = [(torch.from_numpy(d)[None], 0) for d in bank]            # (1x32x32 tensor, label) like datasets.MNIST
    ds.N = len(ds.data)
    out = []
    real = np.random.randint
    for b in range(len(words)):
        state = {"k": 0}

        def scripted(lo, hi=None, _w=words[b], _s=state):
            if hi is None:
                lo, hi = 0, lo
            v = lo + int(_w[_s["k"]]) % (hi - lo)
            _s["k"] += 1
            return v
        np.random.randint = scripted
        try:
            out.append(ds[b])
        finally:
            np.random.randint = real
    return np.stack(out)                                                # [B, T, W, W, 1]


def main():
    mod = load_reference()
    for name, B, T, W, nd, det, nbank, seed in CASES:
        rng = np.random.RandomState(seed)
        bank = rng.rand(nbank, 32, 32).astype(np.float32)
        bank[bank < 0.6] = 0.0                                          # sparse strokes, values up to 1 (sums clip)
        words = rng.randint(0, 2 ** 32, size=(B, nd * (5 + 4 * T)), dtype=np.uint64).astype(np.uint32)
        x = run_reference(mod, bank, words, T, W, nd, det)
        np.savez_compressed(os.path.join(HERE, f"moving_mnist_{name}.npz"), bank=bank, words=words, x=x,
                            seq_len=T, image_size=W, num_digits=nd, deterministic=det)
        print(name, x.shape, float(x.mean()), "clipped px:", int((x == 1.0).sum()))


if __name__ == "__main__":
    main()
