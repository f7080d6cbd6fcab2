"""GPU parity of the LSTM hot path (through the C ABI) against the golden vectors of the real reference
and against the CPU oracle.  Tolerances: fp32 / bf16x3 variants <= 1e-4 relative (north_star fp32 bar),
bf16 variant <= 2e-2."""
import contextlib
import glob
import os

import pytest
import torch

from oracle import lstm_ref
from util import TOL, big_golden_weights, elemerr, make_lstm, relerr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _variants(H):
    return ["fp32", "bf16x3", "bf16"] if H % 64 == 0 else ["fp32"]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "lstm_*.pt"))), ids=os.path.basename)
def test_golden_lstm(path):
    g = torch.load(path, weights_only=False)
    gi, go, H, L, B = g["dims"]
    for variant in _variants(H):
        m = make_lstm(g["state_dict"], rows=B, variant=variant)
        with torch.no_grad():
            m.hidden = m.init_hidden()
            for t, x in enumerate(g["x"]):
                y = m(x.cuda())
                assert relerr(y, g["y"][t]) < TOL[variant], (variant, t)
                for l in range(L):
                    assert relerr(m.hidden[l][0], g["hidden"][t][l][0]) < TOL[variant], (variant, t, l, "h")
                    assert relerr(m.hidden[l][1], g["hidden"][t][l][1]) < TOL[variant], (variant, t, l, "c")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "gauss_*.pt"))), ids=os.path.basename)
def test_golden_gaussian_lstm(path):
    g = torch.load(path, weights_only=False)
    gi, Z, H, L, B = g["dims"]
    for variant in _variants(H):
        m = make_lstm(g["state_dict"], gaussian=True, rows=B, variant=variant)
        with torch.no_grad():
            m.hidden = m.init_hidden()
            for t, x in enumerate(g["x"]):
                z, mu, logvar = m(x.cuda(), eps=g["eps"][t].cuda())
                rz, rmu, rlv = g["out"][t]
                assert relerr(mu, rmu) < TOL[variant], (variant, t)
                assert relerr(logvar, rlv) < TOL[variant], (variant, t)
                assert relerr(z, rz) < TOL[variant], (variant, t)


@pytest.mark.parametrize("variant", ["fp32", "bf16x3", "bf16"])
def test_reference_golden_on_the_fused_step_kernel(variant):
    """The real reference class at G90 / H256 / L2, 300 rows (3 row tiles -> lstm_step_kernel for the tensor-core
    variants), 12 free-running steps: y at steps 0 / 5 / 11 and the final state, max-norm AND element-wise."""
    g = torch.load(os.path.join(GOLD, "big_lstm_g90_h256_r300.pt"), weights_only=False)
    sd, xs = big_golden_weights(g)
    gi, go, H, L, R = g["dims"]
    m = make_lstm(sd, rows=R, variant=variant)
    with torch.no_grad():
        m.hidden = m.init_hidden()
        for t, x in enumerate(xs):
            y = m(x.cuda())
            if t in g["keep"]:
                assert relerr(y, g["y"][t]) < TOL[variant], (t, relerr(y, g["y"][t]))
                assert elemerr(y, g["y"][t], rtol=TOL[variant]) < 1.0, (t, elemerr(y, g["y"][t], rtol=TOL[variant]))
    for got, want in ((m.hidden[L - 1][0], g["h_top"]), (m.hidden[0][1], g["c0"])):
        assert relerr(got, want) < TOL[variant], relerr(got, want)
        assert elemerr(got, want, rtol=TOL[variant]) < 1.0, elemerr(got, want, rtol=TOL[variant])


@pytest.mark.parametrize("variant", ["fp32", "bf16x3", "bf16"])
def test_reference_golden_gaussian_on_the_fused_step_kernel(variant):
    """gaussian_lstm (models/lstm.py:140-175) of the real reference at H256, 300 rows: the PH_GAUSS head of
    lstm_step_kernel (mu / logvar / z with the injected eps)."""
    g = torch.load(os.path.join(GOLD, "big_gauss_g90_z10_h256_r300.pt"), weights_only=False)
    sd, xs = big_golden_weights(g)
    gi, Z, H, L, R = g["dims"]
    m = make_lstm(sd, gaussian=True, rows=R, variant=variant)
    with torch.no_grad():
        m.hidden = m.init_hidden()
        for t, x in enumerate(xs):
            z, mu, logvar = m(x.cuda(), eps=g["eps"][t].cuda())
            rz, rmu, rlv = g["out"][t]
            for name, got, want in (("mu", mu, rmu), ("logvar", logvar, rlv), ("z", z, rz)):
                assert relerr(got, want) < TOL[variant], (t, name, relerr(got, want))
                assert elemerr(got, want, rtol=TOL[variant]) < 1.0, (t, name)
    assert relerr(m.hidden[L - 1][0], g["h_top"]) < TOL[variant]
    assert relerr(m.hidden[0][1], g["c0"]) < TOL[variant]


@pytest.mark.parametrize("variant", ["fp32", "bf16x3"])
def test_gaussian_lstm_5000_rows_vs_oracle(variant):
    """gaussian_lstm at the bench's row count (40 row tiles, every pair of lstm_step_kernel busy)."""
    rows, Z = 5000, 10
    sd = lstm_ref.random_lstm_state_dict(90, Z, 256, 2, seed=31, gaussian=True)
    m = make_lstm(sd, gaussian=True, rows=rows, variant=variant)
    gen = torch.Generator().manual_seed(5)
    hid = lstm_ref.init_hidden(2, rows, 256)
    with torch.no_grad():
        m.hidden = m.init_hidden()
        for t in range(3):
            x = torch.tanh(torch.randn(rows, 90, generator=gen))
            eps = torch.randn(rows, Z, generator=gen)
            z_ref, mu_ref, lv_ref, hid = lstm_ref.gaussian_lstm_forward(sd, x, hid, eps)
            z, mu, logvar = m(x.cuda(), eps=eps.cuda())
            for name, got, want in (("mu", mu, mu_ref), ("logvar", logvar, lv_ref), ("z", z, z_ref)):
                assert relerr(got, want) < TOL[variant], (t, name, relerr(got, want))
                assert elemerr(got, want, rtol=TOL[variant]) < 1.0, (t, name)


@pytest.mark.parametrize("variant", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("rows", [16, 50, 100, 300, 5000])
def test_full_size_vs_oracle(variant, rows):
    """G90/H256/L2 (the reference's sizes): per-step with re-synchronised state, then free-running."""
    sd = lstm_ref.random_lstm_state_dict(90, 90, 256, 2, seed=7)
    m = make_lstm(sd, rows=rows, variant=variant)
    gen = torch.Generator().manual_seed(rows)
    steps = 12 if rows <= 300 else 4
    xs = [torch.tanh(torch.randn(rows, 90, generator=gen)) for _ in range(steps)]
    hid = lstm_ref.init_hidden(2, rows, 256)
    with torch.no_grad():
        m.hidden = m.init_hidden()
        for t, x in enumerate(xs):               # free-running on both sides
            y_ref, hid = lstm_ref.lstm_forward(sd, x, hid)
            y = m(x.cuda())
            tol = TOL[variant]
            assert relerr(y, y_ref) < tol, (t, relerr(y, y_ref))
            for l in range(2):
                assert relerr(m.hidden[l][0], hid[l][0]) < tol, (t, l)
                assert relerr(m.hidden[l][1], hid[l][1]) < tol, (t, l)
        # re-synchronised single step from the oracle's state (foreign hidden tensors -> import path)
        m.hidden = [(h.cuda(), c.cuda()) for h, c in hid]
        y_ref, hid2 = lstm_ref.lstm_forward(sd, xs[0], hid)
        y = m(xs[0].cuda())
        assert relerr(y, y_ref) < TOL[variant]
        assert relerr(m.hidden[1][0], hid2[1][0]) < TOL[variant]


@pytest.mark.parametrize("variant", ["fp32", "bf16x3"])
@pytest.mark.parametrize("rows", [50, 300])
def test_long_free_running(variant, rows):
    """104 recurrent steps (generate_frames.py horizon) against the fp64 oracle: 50 rows (the 16-CTA cluster
    kernel) and 300 rows (three row tiles: the persistent step kernel with the embed Linear folded into layer 0)."""
    sd = lstm_ref.random_lstm_state_dict(90, 90, 256, 2, seed=11)
    sd64 = lstm_ref.to_dtype(sd, torch.float64)
    m = make_lstm(sd, rows=rows, variant=variant)
    gen = torch.Generator().manual_seed(3)
    hid = lstm_ref.init_hidden(2, rows, 256, torch.float64)
    x = torch.tanh(torch.randn(rows, 90, generator=gen))
    worst = worst_el = 0.0
    with torch.no_grad():
        m.hidden = m.init_hidden()
        xg = x.cuda()
        xr = x.double()
        for t in range(104):
            y_ref, hid = lstm_ref.lstm_forward(sd64, xr, hid)
            y = m(xg)
            worst = max(worst, relerr(y, y_ref))
            worst_el = max(worst_el, elemerr(y, y_ref, rtol=1e-4))
            xg, xr = y, y_ref           # autoregressive in latent space
    assert worst < 1e-4, worst
    assert worst_el < 1.0, worst_el
    for l in range(2):
        assert relerr(m.hidden[l][0], hid[l][0]) < 1e-4 and relerr(m.hidden[l][1], hid[l][1]) < 1e-4


@pytest.mark.parametrize("variant", ["fp32", "bf16x3"])
def test_hold_mask(variant):
    """Rows whose hold flag is set keep (h, c) -- generate_frames.py:289-295."""
    sd = lstm_ref.random_lstm_state_dict(90, 90, 256, 2, seed=5)
    S, N = 6, 50
    rows = S * N
    m = make_lstm(sd, rows=rows, variant=variant)
    gen = torch.Generator().manual_seed(0)
    x0 = torch.tanh(torch.randn(rows, 90, generator=gen)).cuda()
    x1 = torch.tanh(torch.randn(rows, 90, generator=gen)).cuda()
    hold = torch.tensor([0, 1, 0, 0, 1, 1], dtype=torch.uint8, device="cuda")
    with torch.no_grad():
        m.hidden = m.init_hidden()
        m(x0)
        before = [(h.clone(), c.clone()) for h, c in m.hidden]
        y_free = None
        m2 = make_lstm(sd, rows=rows, variant=variant)
        m2.hidden = [(h.clone(), c.clone()) for h, c in before]
        y_free = m2(x1)
        y = m(x1, hold=hold, rows_per_flag=N)
        after = m.hidden
        rowmask = hold.bool().repeat_interleave(N)
        for l in range(2):
            assert torch.equal(after[l][0][rowmask], before[l][0][rowmask])
            assert torch.equal(after[l][1][rowmask], before[l][1][rowmask])
            assert torch.equal(after[l][0][~rowmask], m2.hidden[l][0][~rowmask])
            assert torch.equal(after[l][1][~rowmask], m2.hidden[l][1][~rowmask])
        assert torch.equal(y[~rowmask], y_free[~rowmask])
        # one more step must consume the held state consistently (packed copy == fp32 copy)
        y2 = m(x0)
        m2.hidden = [(h.clone(), c.clone()) for h, c in after]
        assert relerr(y2, m2(x0)) < 1e-6


def test_inplace_state_edit_is_seen():
    sd = lstm_ref.random_lstm_state_dict(90, 90, 256, 2, seed=5)
    m = make_lstm(sd, rows=20, variant="bf16x3")
    x = torch.tanh(torch.randn(20, 90)).cuda()
    with torch.no_grad():
        m.hidden = m.init_hidden()
        m(x)
        m.hidden[0][0].mul_(0.5)                  # caller edits our view in place
        hid = [(h.cpu().clone(), c.cpu().clone()) for h, c in m.hidden]
        y = m(x)
        y_ref, _ = lstm_ref.lstm_forward(sd, x.cpu(), hid)
        assert relerr(y, y_ref) < 1e-4


def test_cpu_input_is_rejected():
    from dvg_b200._capi import DvgError
    sd = lstm_ref.random_lstm_state_dict(90, 90, 64, 1, seed=5)
    m = make_lstm(sd, rows=4)
    with torch.no_grad(), pytest.raises(DvgError):
        m(torch.zeros(4, 90))


def test_autograd_delegate_matches_fast_path():
    sd = lstm_ref.random_lstm_state_dict(90, 90, 256, 2, seed=9)
    m = make_lstm(sd, rows=8, variant="fp32")
    x = torch.tanh(torch.randn(8, 90)).cuda()
    m.train()
    m.hidden = m.init_hidden()
    y_grad = m(x)                     # train() mode with grad enabled -> torch ops on the GPU
    assert y_grad.requires_grad
    m.eval()
    m.hidden = m.init_hidden()
    y_eval = m(x)                     # eval() mode takes the kernels even without no_grad (the reference detaches)
    assert not y_eval.requires_grad and relerr(y_eval, y_grad) < 2e-5
    with torch.no_grad():
        m.hidden = m.init_hidden()
        y = m(x)
    assert relerr(y, y_grad) < 2e-5


# --- BASELINE.json configs as parity cases (config 1, 3, 4 row counts; config 5 hidden sizes) -------------------
@pytest.mark.parametrize("rows,H,L", [(16, 256, 2),        # cfg 1: SM-MNIST, batch 16, one rollout
                                      (1600, 256, 2),      # cfg 3: BAIR, 32 rollouts x 50 per GPU
                                      (6400, 256, 2),      # cfg 4: UCF, batch 64 x 100 samples
                                      (257, 512, 1),       # cfg 5 sweep: hidden 512 (odd row count: 3 row tiles)
                                      (130, 1024, 2),      # cfg 5 sweep: hidden 1024
                                      (100, 256, 2),       # one row tile, > 64 rows: persistent kernel, peer CTA all padding
                                      (16, 512, 2),        # cfg 5 sweep: 16 samples at hidden 512 / 1024 (persistent kernel,
                                      (64, 1024, 2)])      #   the cluster kernel is H = 256 only)
def test_config_shapes(rows, H, L):
    sd = lstm_ref.random_lstm_state_dict(90, 90, H, L, seed=H + rows)
    gen = torch.Generator().manual_seed(1)
    xs = [torch.tanh(torch.randn(rows, 90, generator=gen)) for _ in range(3)]
    for variant in ("bf16x3", "fp32"):
        m = make_lstm(sd, rows=rows, variant=variant)
        hid = lstm_ref.init_hidden(L, rows, H)
        with torch.no_grad():
            m.hidden = m.init_hidden()
            for t, x in enumerate(xs):
                y_ref, hid = lstm_ref.lstm_forward(sd, x, hid)
                y = m(x.cuda())
                tol = TOL[variant]
                assert relerr(y, y_ref) < tol, (variant, t, relerr(y, y_ref))
                assert relerr(m.hidden[L - 1][0], hid[L - 1][0]) < tol, (variant, t)
                assert relerr(m.hidden[0][1], hid[0][1]) < tol, (variant, t)


def test_unfused_tensor_core_path_matches_fused(monkeypatch):
    """The one-launch-per-GEMM tensor-core path (used for a single row tile) and the fused persistent step
    kernel must agree: run 300 rows (fused) against the same rows split into 3 calls of 100 (single tile)."""
    sd = lstm_ref.random_lstm_state_dict(90, 90, 256, 2, seed=21)
    x = torch.tanh(torch.randn(300, 90, generator=torch.Generator().manual_seed(2))).cuda()
    with torch.no_grad():
        big = make_lstm(sd, rows=300, variant="bf16x3")
        big.hidden = big.init_hidden()
        y_big = big(x)
        y_big = big(y_big)
        outs = []
        for i in range(3):
            small = make_lstm(sd, rows=100, variant="bf16x3")
            small.hidden = small.init_hidden()
            y = small(x[i * 100:(i + 1) * 100])
            outs.append(small(y))
        y_small = torch.cat(outs)
    assert relerr(y_big, y_small) < 2e-5


@pytest.mark.parametrize("kind,variant", [("lstm", "bf16x3"), ("lstm", "bf16"), ("gaussian", "bf16x3")])
def test_dropin_chained_steps_equal_stream_ordered_steps(kind, variant):
    """``module.chained()`` (dvg_lstm_chain_begin/_end under the drop-in classes): 5000 rows so that the grid covers the
    machine and the launches really overlap; outputs and final state bit-identical to the plain loop, for both
    tensor-core variants and for gaussian_lstm (the PH_GAUSS head with its injected eps)."""
    rows, T, Z = 5000, 14, 10
    gauss = kind == "gaussian"
    sd = lstm_ref.random_lstm_state_dict(90, Z if gauss else 90, 256, 2, seed=21, gaussian=gauss)
    m = make_lstm(sd, gaussian=gauss, rows=rows, variant=variant)
    g = torch.Generator().manual_seed(2)
    xs = torch.tanh(torch.randn(T, rows, 90, generator=g)).cuda()
    eps = torch.randn(T, rows, Z, generator=g).cuda()

    def step(t):
        if gauss:
            return torch.cat(m(xs[t], eps=eps[t]), 1)
        return m(xs[t])

    res = []
    with torch.no_grad():
        for chained in (False, True, True):
            m.hidden = m.init_hidden()
            with (m.chained() if chained else contextlib.nullcontext()):
                ys = [step(t) for t in range(T)]
            torch.cuda.synchronize()
            res.append((torch.stack(ys), [(h.clone(), c.clone()) for h, c in m.hidden]))
    for ys, hid in res[1:]:
        assert torch.equal(ys, res[0][0])
        for (h1, c1), (h0, c0) in zip(hid, res[0][1]):
            assert torch.equal(h1, h0) and torch.equal(c1, c0)
