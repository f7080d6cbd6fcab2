"""Pin oracle/lstm_ref.py against golden vectors produced by the real reference
(models/lstm.py, via tests/golden/make_golden.py).  CPU only."""
import glob
import os

import pytest
import torch

from oracle import lstm_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "lstm_*.pt"))), ids=os.path.basename)
def test_lstm_oracle_matches_reference(path):
    g = torch.load(path, weights_only=False)
    gi, go, H, L, B = g["dims"]
    hidden = lstm_ref.init_hidden(L, B, H)
    for t, x in enumerate(g["x"]):
        y, hidden = lstm_ref.lstm_forward(g["state_dict"], x, hidden)
        torch.testing.assert_close(y, g["y"][t], rtol=1e-6, atol=1e-7)
        for l in range(L):
            torch.testing.assert_close(hidden[l][0], g["hidden"][t][l][0], rtol=1e-6, atol=1e-7)
            torch.testing.assert_close(hidden[l][1], g["hidden"][t][l][1], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "gauss_*.pt"))), ids=os.path.basename)
def test_gaussian_lstm_oracle_matches_reference(path):
    g = torch.load(path, weights_only=False)
    gi, Z, H, L, B = g["dims"]
    hidden = lstm_ref.init_hidden(L, B, H)
    for t, x in enumerate(g["x"]):
        z, mu, logvar, hidden = lstm_ref.gaussian_lstm_forward(g["state_dict"], x, hidden, g["eps"][t])
        rz, rmu, rlv = g["out"][t]
        torch.testing.assert_close(mu, rmu, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(logvar, rlv, rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(z, rz, rtol=1e-6, atol=1e-7)


def test_oracle_matches_reference_at_full_size():
    """G90 / H256 / L2, 300 rows, 12 free-running steps of the REAL reference classes (big_* goldens)."""
    from util import big_golden_weights
    g = torch.load(os.path.join(GOLD, "big_lstm_g90_h256_r300.pt"), weights_only=False)
    sd, xs = big_golden_weights(g)
    gi, go, H, L, R = g["dims"]
    hidden = lstm_ref.init_hidden(L, R, H)
    for t, x in enumerate(xs):
        y, hidden = lstm_ref.lstm_forward(sd, x, hidden)
        if t in g["keep"]:
            torch.testing.assert_close(y, g["y"][t], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(hidden[L - 1][0], g["h_top"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(hidden[0][1], g["c0"], rtol=1e-5, atol=1e-6)
    g = torch.load(os.path.join(GOLD, "big_gauss_g90_z10_h256_r300.pt"), weights_only=False)
    sd, xs = big_golden_weights(g)
    gi, Z, H, L, R = g["dims"]
    hidden = lstm_ref.init_hidden(L, R, H)
    for t, x in enumerate(xs):
        z, mu, logvar, hidden = lstm_ref.gaussian_lstm_forward(sd, x, hidden, g["eps"][t])
        rz, rmu, rlv = g["out"][t]
        torch.testing.assert_close(mu, rmu, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(logvar, rlv, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(z, rz, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(hidden[L - 1][0], g["h_top"], rtol=1e-5, atol=1e-6)


def test_fp64_oracle_brackets_fp32():
    sd = lstm_ref.random_lstm_state_dict(90, 90, 256, 2, seed=3)
    x = torch.tanh(torch.randn(50, 90, generator=torch.Generator().manual_seed(0)))
    h32 = lstm_ref.init_hidden(2, 50, 256)
    h64 = lstm_ref.init_hidden(2, 50, 256, torch.float64)
    sd64 = lstm_ref.to_dtype(sd, torch.float64)
    for _ in range(20):
        y32, h32 = lstm_ref.lstm_forward(sd, x, h32)
        y64, h64 = lstm_ref.lstm_forward(sd64, x.double(), h64)
    assert (y32.double() - y64).abs().max() < 1e-5
    assert (h32[1][0].double() - h64[1][0]).abs().max() < 1e-5
