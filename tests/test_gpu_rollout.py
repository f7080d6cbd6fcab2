"""N-diverse-futures bookkeeping: the sample-batched engine / drivers against the sequential oracle loops
(oracle/rollout_ref.py restating generate_frames.py:138-178, :249-300, train.py:262-289), on the same inputs,
weights and injected noise.  Encoder/decoder are small latent-space stand-ins so the test isolates the hot
path (the conv nets stay on the stock PyTorch path and are not under test)."""
import contextlib

import numpy as np
import pytest
import torch

from oracle import gp_ref, lstm_ref, rollout_ref
from util import make_gp, make_lstm, relerr

pytestmark = pytest.mark.gpu

G, H, L, M = 90, 256, 2, 40


class ToyCodec:
    """x is a 'frame' [N, G]; encoder = tanh(x A) (+ skip = x), decoder = tanh(v Bm + 0.1 skip)."""

    def __init__(self, device, dtype):
        g = torch.Generator().manual_seed(42)
        self.A = (torch.randn(G, G, generator=g) / G ** 0.5).to(device, dtype)
        self.Bm = (torch.randn(G, G, generator=g) / G ** 0.5).to(device, dtype)

    def encoder(self, x):
        return torch.tanh(x.to(self.A.dtype) @ self.A), [x]

    def decoder(self, inp):
        vec, skip = inp
        return torch.tanh(vec.to(self.A.dtype) @ self.Bm + 0.1 * skip[0].to(self.A.dtype))


def _models(seed=3):
    sd = lstm_ref.random_lstm_state_dict(G, G, H, L, seed=seed)
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(G, M, seed=seed, trained_like=True, smooth_mean=True)
    return sd, gp_sd, lik_sd


@pytest.mark.parametrize("variant", ["fp32", "bf16x3"])
def test_diverse_rollout_matches_sequential_oracle(variant):
    from dvg_b200.rollout import diverse_rollout
    sd, gp_sd, lik_sd = _models()
    B, S, n_past, n_eval = 6, 4, 3, 12
    g = torch.Generator().manual_seed(0)
    x = [torch.rand(B, G, generator=g) for _ in range(n_eval)]
    eps = {(s, i): torch.randn(G, B, generator=g) for s in range(S) for i in range(n_eval)}
    cpu = ToyCodec("cpu", torch.float32)
    om = rollout_ref.OracleModels(sd, gp_sd, lik_sd, cpu.encoder, cpu.decoder, gp_mode="direct")
    ref = rollout_ref.diverse_rollout(om, x, n_past, n_eval, S, eps, resample_every=5)
    fp = make_lstm(sd, rows=B, variant=variant)
    gp, lik = make_gp(gp_sd, lik_sd)
    gpu = ToyCodec("cuda", torch.float32)
    got = diverse_rollout(fp, gp, lik, gpu.encoder, gpu.decoder, [t.cuda() for t in x], n_past, n_eval, S,
                          eps=eps, resample_every=5, variant=variant)
    assert len(got) == n_eval
    for t in range(n_eval):
        for s in range(S):
            assert relerr(got[t][s], ref[s][t]) < 1e-4, (t, s)
    # samples are identical until the first resample step and differ afterwards
    assert torch.equal(got[4][0], got[4][1])
    assert not torch.equal(got[6][0], got[6][1])


def test_trigger_rollout_matches_sequential_oracle():
    from dvg_b200.rollout import trigger_rollout
    sd, gp_sd, lik_sd = _models(seed=4)
    B, S, warm, n_steps = 8, 5, 6, 40
    g = torch.Generator().manual_seed(1)
    x0 = torch.rand(B, G, generator=g)
    eps = {(s, i): torch.randn(G, B, generator=g) for s in range(S) for i in range(n_steps)}
    cols = [0, 1, 2, 3, 4]
    fp = make_lstm(sd, rows=B, variant="bf16x3")
    gp, lik = make_gp(gp_sd, lik_sd)
    gpu = ToyCodec("cuda", torch.float32)
    got = trigger_rollout(fp, gp, lik, gpu.encoder, gpu.decoder, x0.cuda(), S, eps=eps, warmup=warm, n_steps=n_steps,
                          stat_col=3, stat_cols_warmup=cols, skip_until=5)
    trig = got["triggers"].cpu().numpy().astype(bool)
    vals = got["values"].cpu().numpy()
    cpu = ToyCodec("cpu", torch.float64)
    sd64 = lstm_ref.to_dtype(sd, torch.float64)
    n_fired = 0
    for s in range(S):
        om = rollout_ref.OracleModels(sd64, gp_sd, lik_sd, cpu.encoder, cpu.decoder, dtype=torch.float64,
                                      gp_mode="direct")
        ref = rollout_ref.trigger_rollout(om, x0.double(), {i: eps[(s, i)] for i in range(n_steps)}, warmup=warm,
                                          n_steps=n_steps, stat_col_warmup=cols[s], stat_col=3, skip_until=5)
        for i in range(n_steps):
            # free-running comparison is only meaningful while the decision histories agree
            v, thr = float(ref["values"][i]), float(ref["thresholds"][i])
            assert abs(vals[i, s] - v) <= 2e-4 * abs(v), (s, i)
            if i >= warm and abs(v - thr) <= 1e-3 * abs(thr):
                break                                   # inside the tolerance band: histories may fork
            assert bool(trig[i, s]) == bool(ref["triggers"][i]), (s, i)
            n_fired += int(trig[i, s])
            assert relerr(got["latents"][i][s], ref["latents"][i]) < 1e-4, (s, i)
            assert relerr(got["gen_seq"][i][s], ref["gen_seq"][i]) < 1e-4, (s, i)
    # (the closed toy loop settles quickly, so triggers are rare here; the trigger -> hold -> rsample
    #  semantics are pinned by test_latent_trigger_rollout_crafted below)


@pytest.mark.parametrize("variant", ["fp32", "bf16x3"])
def test_latent_trigger_rollout_crafted(variant):
    """Engine-level GPtrigger_gen semantics with guaranteed triggers: value / window / threshold / decision,
    LSTM state held on a triggered step, GP sample substituted for the rollouts that fired."""
    from dvg_b200.rollout import RolloutConfig, RolloutEngine
    from util import check_latent_rollout, crafted_trigger_case
    B, S, T, W = 10, 7, 24, 6
    sd = lstm_ref.random_lstm_state_dict(G, G, H, L, seed=6)
    gp_sd, lik_sd, lat, eps, jumps = crafted_trigger_case(G, M, B, S, T, W, seed=1)
    fp = make_lstm(sd, rows=B, variant=variant)
    gp, lik = make_gp(gp_sd, lik_sd)
    eng = RolloutEngine(fp, gp, lik, RolloutConfig(n_points=B, n_rollouts=S, window=W, variant=variant))
    out = torch.empty(T, S * B, G, device="cuda")
    masks = torch.zeros(T, S, dtype=torch.uint8, device="cuda")
    values = torch.zeros(T, S, device="cuda")
    with torch.no_grad():
        eng.latent_rollout(lat.cuda(), eps.cuda(), out, masks=masks, values=values)
    torch.cuda.synchronize()
    checked, fired = check_latent_rollout(sd, gp_sd, lik_sd, lat, eps, out.cpu(), masks.cpu(), values.cpu(), B, W)
    assert checked > 100
    # a jump fires unless an earlier jump of the same rollout still sits in the 6-step window
    assert fired >= 4, (fired, sorted(jumps), masks.cpu().nonzero().tolist())


def test_kth_shaped_trigger_rollout_vs_oracle():
    """The bench's own shape per rollout (B = 50 points: the in-kernel resample solves 50 x 50 problems; 8 rollouts =
    400 rows = 4 row tiles -> lstm_step_kernel, window 12 as in generate_frames.py:266) with crafted fires, through
    dvg_rollout_step, against the sequential oracle."""
    from dvg_b200.rollout import RolloutConfig, RolloutEngine
    from util import check_latent_rollout, crafted_trigger_case
    B, S, T, W = 50, 8, 30, 12
    sd = lstm_ref.random_lstm_state_dict(G, G, H, L, seed=16)
    gp_sd, lik_sd, lat, eps, jumps = crafted_trigger_case(G, M, B, S, T, W, seed=4, n_jumps=6)
    fp = make_lstm(sd, rows=B, variant="bf16x3")
    gp, lik = make_gp(gp_sd, lik_sd)
    eng = RolloutEngine(fp, gp, lik, RolloutConfig(n_points=B, n_rollouts=S, window=W, variant="bf16x3"))
    out = torch.empty(T, S * B, G, device="cuda")
    masks = torch.zeros(T, S, dtype=torch.uint8, device="cuda")
    values = torch.zeros(T, S, device="cuda")
    with torch.no_grad():
        eng.latent_rollout(lat.cuda(), eps.cuda(), out, masks=masks, values=values)
    torch.cuda.synchronize()
    checked, fired = check_latent_rollout(sd, gp_sd, lik_sd, lat, eps, out.cpu(), masks.cpu(), values.cpu(), B, W)
    assert checked > 100 and fired >= 1, (checked, fired, sorted(jumps), masks.cpu().nonzero().tolist())


@pytest.mark.parametrize("B,S", [(16, 1), (30, 2), (64, 1), (5, 1)])
def test_small_batch_trigger_rollout_vs_oracle(B, S):
    """<= 64 rows (BASELINE configs[0]: batch 16, one rollout; configs[3]: batch 64): the 16-CTA cluster kernel with the
    trigger fused (lstm_small.cu), crafted fires, against the sequential oracle."""
    from dvg_b200.rollout import RolloutConfig, RolloutEngine
    from util import check_latent_rollout, crafted_trigger_case
    T, W = 26, 6
    sd = lstm_ref.random_lstm_state_dict(G, G, H, L, seed=26)
    gp_sd, lik_sd, lat, eps, jumps = crafted_trigger_case(G, M, B, S, T, W, seed=7, n_jumps=3)
    fp = make_lstm(sd, rows=B, variant="bf16x3")
    gp, lik = make_gp(gp_sd, lik_sd)
    eng = RolloutEngine(fp, gp, lik, RolloutConfig(n_points=B, n_rollouts=S, window=W, variant="bf16x3", stat_col=min(3, B - 1)))
    out = torch.empty(T, S * B, G, device="cuda")
    masks = torch.zeros(T, S, dtype=torch.uint8, device="cuda")
    values = torch.zeros(T, S, device="cuda")
    with torch.no_grad():
        eng.latent_rollout(lat.cuda(), eps.cuda(), out, masks=masks, values=values)
    torch.cuda.synchronize()
    checked, fired = check_latent_rollout(sd, gp_sd, lik_sd, lat, eps, out.cpu(), masks.cpu(), values.cpu(), B, W,
                                          stat_col=min(3, B - 1))
    assert checked >= 10 and fired >= 1, (checked, fired, sorted(jumps), masks.cpu().nonzero().tolist())


def _sm_count():
    return torch.cuda.get_device_properties(0).multi_processor_count


@pytest.mark.parametrize("kind", ["trigger", "plain"])
def test_chained_steps_equal_stream_ordered_steps(kind):
    """dvg_lstm_chain_begin/_end (include/dvg_b200.h): consecutive step launches overlap on the GPU, synchronised per tile
    through the previous launch's counters instead of the grid boundary.  Needs a grid that covers the machine, so the
    bench's own shape: 5000 rows (B = 50 x S = 100), with crafted fires -- a fired step makes the next launch wait for the
    restored state rows.  The chained sequence (eager and replayed from a CUDA graph) must equal the stream-ordered one
    bit for bit: outputs of every step, masks, and the final state; two rollouts are also replayed on the CPU oracle."""
    from dvg_b200.rollout import RolloutConfig, RolloutEngine
    from util import check_latent_rollout, crafted_trigger_case
    B, S, T, W = 50, 100, 24, 12
    sd = lstm_ref.random_lstm_state_dict(G, G, H, L, seed=31)
    gp_sd, lik_sd, lat, eps, jumps = crafted_trigger_case(G, M, B, S, T, W, seed=9, n_jumps=10)
    fp = make_lstm(sd, rows=B, variant="bf16x3")
    gp, lik = make_gp(gp_sd, lik_sd)
    eng = RolloutEngine(fp, gp, lik, RolloutConfig(n_points=B, n_rollouts=S, window=W, variant="bf16x3"))
    lat_d, eps_d = lat.cuda(), eps.cuda()

    def run(chained, out, masks, vals=None):
        eng.reset()
        ctx = eng.chained() if chained else contextlib.nullcontext()
        with ctx:
            for t in range(T):
                if kind == "plain":
                    eng.step_manual_mode(lat_d[t], None, out[t], resample=False)
                else:
                    eng._mask_buf = masks[t]
                    eng._value_buf = vals[t] if vals is not None else eng.value
                    eng.step_trigger_mode(lat_d[t], eps_d[t], out[t], warmup=t < W)
        eng._mask_buf, eng._value_buf = eng.mask, eng.value
        return [(h.clone(), c.clone()) for h, c in eng.hidden()]

    outs, masks, states = [], [], []
    vals = torch.zeros(T, S, device="cuda")
    with torch.no_grad():
        for chained in (False, True, True):
            o = torch.empty(T, S * B, G, device="cuda")
            m = torch.zeros(T, S, dtype=torch.uint8, device="cuda")
            states.append(run(chained, o, m, vals))
            torch.cuda.synchronize()
            outs.append(o)
            masks.append(m)
        # the same chain replayed from a CUDA graph (the launches then run back to back without host gaps)
        og = torch.empty(T, S * B, G, device="cuda")
        mg = torch.zeros(T, S, dtype=torch.uint8, device="cuda")
        run(True, og, mg)                      # warm outside capture
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            with torch.cuda.graph(g):
                sg = run(True, og, mg)
        for _ in range(3):
            og.zero_()
            g.replay()
        torch.cuda.synchronize()
    for o, m, stt in zip(outs[1:] + [og], masks[1:] + [mg], states[1:] + [sg]):
        assert torch.equal(m, masks[0])
        assert torch.equal(o, outs[0])
        for (h1, c1), (h0, c0) in zip(stt, states[0]):
            assert torch.equal(h1, h0) and torch.equal(c1, c0)
    if kind == "trigger":
        fired = int(masks[0].sum())
        assert fired >= 5, (fired, sorted(jumps))
        # oracle replay of two rollouts that fired and one that did not
        rolls = sorted({s_ for _, s_ in jumps})[:2] + [next(s_ for s_ in range(S) if all(s_ != j for _, j in jumps))]
        for s_ in rolls:
            rows = slice(s_ * B, (s_ + 1) * B)
            checked, f = check_latent_rollout(sd, gp_sd, lik_sd, lat[:, rows], eps[:, s_:s_ + 1], outs[1][:, rows].cpu(),
                                              masks[1][:, s_:s_ + 1].cpu(), vals[:, s_:s_ + 1].cpu(), B, W)
            assert checked >= 5


def test_cuda_graph_latent_rollout_equals_eager():
    from dvg_b200.rollout import RolloutConfig, RolloutEngine
    from util import crafted_trigger_case
    B, S, T, W = 10, 12, 16, 6      # a single outlier can only fire for window >= 6 (1/W + 2.01 sqrt(W-1)/W < 1)
    sd = lstm_ref.random_lstm_state_dict(G, G, H, L, seed=5)
    gp_sd, lik_sd, lat, eps, jumps = crafted_trigger_case(G, M, B, S, T, W, seed=2)
    fp = make_lstm(sd, rows=B, variant="bf16x3")
    gp, lik = make_gp(gp_sd, lik_sd)
    eng = RolloutEngine(fp, gp, lik, RolloutConfig(n_points=B, n_rollouts=S, window=W))
    lat, eps = lat.cuda(), eps.cuda()
    out_e = torch.empty(T, S * B, G, device="cuda")
    m_e = torch.zeros(T, S, dtype=torch.uint8, device="cuda")
    with torch.no_grad():
        eng.reset()
        eng.latent_rollout(lat, eps, out_e, masks=m_e)
        out_g = torch.empty_like(out_e)
        m_g = torch.zeros_like(m_e)
        graph = eng.capture_latent_rollout(lat, eps, out_g, masks=m_g)
        for _ in range(3):
            graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(m_e, m_g)
    assert torch.equal(out_e, out_g)
    assert 0 < int(m_e.sum()) < m_e.numel()


def test_streaming_pipeline_matches_direct_rollout():
    """LatentRolloutPipeline (pinned host in, two graph-backed buffer sets, results D2H) against the direct engine call,
    with full outputs and with winners-only results (post hook captured in the graphs)."""
    from dvg_b200 import shard
    from dvg_b200.rollout import LatentRolloutPipeline, RolloutConfig, RolloutEngine, score_rollouts
    from util import crafted_trigger_case
    B, S, T, W = 10, 12, 16, 6
    sd = lstm_ref.random_lstm_state_dict(G, G, H, L, seed=5)
    gp_sd, lik_sd, lat, eps, jumps = crafted_trigger_case(G, M, B, S, T, W, seed=2)
    fp = make_lstm(sd, rows=B, variant="bf16x3")
    gp, lik = make_gp(gp_sd, lik_sd)
    eng = RolloutEngine(fp, gp, lik, RolloutConfig(n_points=B, n_rollouts=S, window=W))
    out_d = torch.empty(T, S * B, G, device="cuda")
    m_d = torch.zeros(T, S, dtype=torch.uint8, device="cuda")
    with torch.no_grad():
        eng.reset()
        eng.latent_rollout(lat.cuda(), eps.cuda(), out_d, masks=m_d)
    target = lat[:, :B].cuda().contiguous()

    def post(o):
        sc = score_rollouts(o, target, S, B)
        best = shard.select_best(sc, higher_is_better=False)
        return sc, best, shard.gather_winners(o.view(T, S, B, G).permute(1, 2, 0, 3), best, S)

    sc_d, best_d, win_d = post(out_d)
    lat_h, eps_h = lat.pin_memory(), eps.pin_memory()
    with torch.no_grad():
        full = LatentRolloutPipeline(eng, T)
        tickets = [full.submit(lat_h, eps_h) for _ in range(3)]          # exercises both buffer sets
        for tk in tickets:
            o, m, extra = full.result(tk)
            assert extra == () and torch.equal(o, out_d.cpu()) and torch.equal(m, m_d.cpu())
        win = LatentRolloutPipeline(eng, T, post=post, full_output=False)
        tickets = [win.submit(lat_h, eps_h) for _ in range(3)]
        for tk in tickets:
            o, m, extra = win.result(tk)
            assert o is None and torch.equal(m, m_d.cpu())
            assert torch.equal(extra[0], sc_d.cpu()) and torch.equal(extra[1], best_d.cpu())
            assert torch.equal(extra[2], win_d.cpu())
        assert win.d2h_bytes() == m_d.numel() + sc_d.numel() * 4 + best_d.numel() * 8 + win_d.numel() * 4
    assert 0 < int(m_d.sum())


def test_many_rollouts_fire_in_the_same_step():
    """Stress of the end-of-kernel restore + resample: most rollouts jump (and fire) at the same time step, more
    (rollout, dim) problems than CTAs; checked against the sequential oracle like the crafted case."""
    from dvg_b200.rollout import RolloutConfig, RolloutEngine
    from util import check_latent_rollout, crafted_trigger_case
    B, S, T, W = 10, 40, 12, 6
    sd = lstm_ref.random_lstm_state_dict(G, G, H, L, seed=9)
    gp_sd, lik_sd, lat, eps, _ = crafted_trigger_case(G, M, B, S, T, W, seed=3, n_jumps=1)
    g = torch.Generator().manual_seed(11)
    lat = -0.8 + 0.1 * torch.tanh(torch.randn(T, S * B, G, generator=g))
    for s in range(S):
        if s % 4 != 3:                                    # 30 of the 40 rollouts jump together at t = 8
            lat[8, s * B:(s + 1) * B] = 0.75 + 0.2 * torch.tanh(torch.randn(B, G, generator=g))
    fp = make_lstm(sd, rows=B, variant="bf16x3")
    gp, lik = make_gp(gp_sd, lik_sd)
    eng = RolloutEngine(fp, gp, lik, RolloutConfig(n_points=B, n_rollouts=S, window=W, variant="bf16x3"))
    out = torch.empty(T, S * B, G, device="cuda")
    masks = torch.zeros(T, S, dtype=torch.uint8, device="cuda")
    values = torch.zeros(T, S, device="cuda")
    with torch.no_grad():
        eng.latent_rollout(lat.cuda(), eps.cuda(), out, masks=masks, values=values)
    torch.cuda.synchronize()
    checked, fired = check_latent_rollout(sd, gp_sd, lik_sd, lat, eps, out.cpu(), masks.cpu(), values.cpu(), B, W)
    assert int(masks[8].sum()) >= 25, masks[8].tolist()
    assert fired >= 25 and checked > 100


@pytest.mark.parametrize("env", [{"DVG_STEP_SCHED": "2"}, {"DVG_STEP_SCHED": "1"}, {"DVG_STEP_PDL": "0"}])
def test_step_kernel_developer_switches(env):
    """The opt-in item schedules and the PDL switch of the step kernel must not change results: re-run the
    config-shape LSTM parity tests (R = 5000 is the shape the two-layer pattern applies to) in a subprocess."""
    import os
    import subprocess
    import sys
    e = dict(os.environ)
    e.update(env)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_lstm.py", "-q", "-x", "-k", "config_shapes or full_size"],
                       cwd=root, env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_plot_rollout_matches_sequential_oracle():
    """train.py:256-310 plot(): 5 futures, one resample at step 10, best-of-5 by summed squared error."""
    from dvg_b200.rollout import plot_rollout
    sd, gp_sd, lik_sd = _models(seed=7)
    B, S, n_past, n_eval = 6, 5, 3, 14
    g = torch.Generator().manual_seed(2)
    x = [torch.rand(B, G, generator=g) for _ in range(n_eval)]
    eps = {(s, 10): torch.randn(G, B, generator=g) for s in range(S)}
    fp = make_lstm(sd, rows=B, variant="bf16x3")
    gp, lik = make_gp(gp_sd, lik_sd)
    gpu = ToyCodec("cuda", torch.float32)
    got = plot_rollout(fp, gp, lik, gpu.encoder, gpu.decoder, [t.cuda() for t in x], n_past, n_eval, S, eps=eps)
    cpu = ToyCodec("cpu", torch.float64)
    om = rollout_ref.OracleModels(lstm_ref.to_dtype(sd, torch.float64), gp_sd, lik_sd, cpu.encoder, cpu.decoder,
                                  dtype=torch.float64, gp_mode="direct")
    ref = rollout_ref.diverse_rollout(om, [t.double() for t in x], n_past, n_eval, S, eps, resample_every=None,
                                      resample_at=[10])
    sse = torch.zeros(S, B, dtype=torch.float64)
    for s in range(S):
        for t in range(n_eval):
            assert relerr(got["samples"][t][s], ref[s][t]) < 1e-4, (s, t)
            sse[s] += (x[t].double() - ref[s][t]).pow(2).reshape(B, -1).sum(1)
    assert relerr(got["sse"], sse) < 1e-3
    assert got["best"].cpu().tolist() == sse.argmin(0).tolist()
    # the samples differ only from the resample step on
    assert torch.equal(got["samples"][9][0], got["samples"][9][1]) and not torch.equal(got["samples"][10][0], got["samples"][10][1])


def test_score_rollouts_matches_torch():
    from dvg_b200.rollout import score_rollouts
    T, S, B = 7, 5, 6
    g = torch.Generator().manual_seed(3)
    out = torch.randn(T, S * B, G, generator=g).cuda()
    target = torch.randn(T, B, G, generator=g).cuda()
    got = score_rollouts(out, target, S, B)
    want = (out.view(T, S, B, G) - target.view(T, 1, B, G)).double().pow(2).mean(dim=(0, 3))
    assert relerr(got, want) < 1e-5


def test_make_gifs_pixel_space_with_reference_convnets():
    """Whole make_gifs computation in pixel space with the dcgan_64 encoder/decoder (stock PyTorch path, eval mode)
    against the sequential CPU oracle using the same nets, weights and injected noise; then SSIM-based best-of-N."""
    import numpy as np
    from dvg_b200.convnets import make_codec
    from dvg_b200.rollout import make_gifs
    from oracle import metrics_ref
    torch.manual_seed(0)
    g_dim, B, S, n_past, n_eval = 90, 3, 3, 2, 7
    enc, dec = make_codec("dcgan_64", g_dim, 1)
    for m in (enc, dec):
        for mod in m.modules():
            if isinstance(mod, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)):
                mod.weight.data.normal_(0.0, 0.02); mod.bias.data.zero_()
            elif isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.data.normal_(1.0, 0.02); mod.bias.data.zero_()
                mod.running_mean.normal_(0, 0.05); mod.running_var.uniform_(0.5, 1.5)
        m.eval()
    sd = lstm_ref.random_lstm_state_dict(g_dim, g_dim, H, L, seed=12)
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(g_dim, M, seed=12, trained_like=True, smooth_mean=True)
    g = torch.Generator().manual_seed(5)
    x = [torch.rand(B, 1, 64, 64, generator=g) for _ in range(n_eval)]
    eps = {(s, i): torch.randn(g_dim, B, generator=g) for s in range(S) for i in range(n_eval)}
    with torch.no_grad():
        om = rollout_ref.OracleModels(sd, gp_sd, lik_sd, enc, dec, gp_mode="direct")
        ref = rollout_ref.diverse_rollout(om, x, n_past, n_eval, S, eps, resample_every=3)
        ref_post = rollout_ref.posterior_rollout(om, x, n_past, n_eval)
    fp = make_lstm(sd, rows=B, variant="bf16x3")
    gp, lik = make_gp(gp_sd, lik_sd)
    enc_g, dec_g = enc.cuda(), dec.cuda()
    out = make_gifs(fp, gp, lik, enc_g, dec_g, [t.cuda() for t in x], n_past, n_eval, S, eps=eps, resample_every=3)
    for t in range(n_eval):
        assert relerr(out["posterior"][t], ref_post[t]) < 2e-3, t
        for s in range(S):
            assert relerr(out["samples"][t][s], ref[s][t]) < 2e-3, (t, s)
    # metrics + selection against the oracle metrics evaluated on the oracle frames: utils.eval_seq (default, what the
    # script ranks by) and the finn variant
    for metric, fn in (("skimage", metrics_ref.eval_seq), ("finn", metrics_ref.finn_eval_seq)):
        if metric == "finn":
            out = make_gifs(fp, gp, lik, enc_g, dec_g, [t.cuda() for t in x], n_past, n_eval, S, eps=eps, resample_every=3,
                            metric="finn")
        want = np.zeros((B, S, n_eval - n_past))
        for s in range(S):
            a, _ = fn([x[t].numpy() for t in range(n_past, n_eval)], [ref[s][t].numpy() for t in range(n_past, n_eval)])
            want[:, s] = a
        assert np.abs(out["ssim"].cpu().numpy() - want).max() < 2e-3, metric
    enc.cpu(); dec.cpu()


def test_generation_and_var_value_helpers():
    """The script-level helpers of generate_frames.py:220-232 (``generation``, ``var_value``) against the oracle."""
    from dvg_b200.convnets import make_codec
    from dvg_b200.rollout import generation, var_value
    from oracle import trigger_ref
    torch.manual_seed(0)
    g_dim, B = 90, 6
    enc, dec = make_codec("dcgan_64", g_dim, 1)
    enc.eval(); dec.eval()
    sd = lstm_ref.random_lstm_state_dict(g_dim, g_dim, H, L, seed=15)
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(g_dim, M, seed=15, trained_like=True, smooth_mean=True)
    g = torch.Generator().manual_seed(8)
    x0, x1 = torch.rand(B, 1, 64, 64, generator=g), torch.rand(B, 1, 64, 64, generator=g)
    ctx = torch.rand(12, generator=g)
    with torch.no_grad():
        h0, skip = enc(x0)
        pred, _ = lstm_ref.lstm_forward(sd, h0, lstm_ref.init_hidden(L, B, H))
        want_frame = dec([pred, skip])
        h1 = enc(x1)[0]
        ref = gp_ref.predictive(gp_sd, lik_sd, gp_ref.latent_to_gp_input(h1), torch.float64, "direct", full_cov=False)
        want_value = trigger_ref.trigger_value(ref["variance"].float().numpy(), 3)
        want_ctx = trigger_ref.slide(ctx.numpy(), want_value)
    fp = make_lstm(sd, rows=B, variant="bf16x3")
    gp, lik = make_gp(gp_sd, lik_sd)
    enc_g, dec_g = enc.cuda(), dec.cuda()
    fp.hidden = fp.init_hidden()
    got = generation(fp, enc_g, dec_g, x0.cuda(), [s.cuda() for s in skip])
    assert relerr(got, want_frame) < 2e-3
    value, new_ctx = var_value(gp, lik, enc_g, x1.cuda(), ctx)
    assert value.is_cuda and new_ctx.is_cuda and new_ctx.shape == (12,)
    assert abs(value.item() - float(want_value)) < 1e-4 * float(want_value)
    assert np.allclose(new_ctx.cpu().numpy(), want_ctx, rtol=1e-4)
    enc.cpu(); dec.cpu()


@pytest.mark.parametrize("model,nc", [("dcgan_64", 1), ("vgg_64", 3)])
def test_batched_codec_pixel_rollout_matches_oracle(model, nc):
    """Sample-batched conv execution (folded BN, channels-last, shared-skip decoder, row chunks)
    and the CUDA-graphed ``PixelRollout`` against the sequential CPU oracle on the plain nets."""
    from dvg_b200.codec import BatchedCodec
    from dvg_b200.convnets import make_codec
    from dvg_b200.rollout import PixelRollout, diverse_rollout, resample_steps
    torch.manual_seed(0)
    g_dim, B, S, n_past, n_eval = 90, 3, 4, 3, 8
    enc, dec = make_codec(model, g_dim, nc)
    for m in (enc, dec):
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.weight.data.normal_(1.0, 0.02); mod.bias.data.zero_()
                mod.running_mean.normal_(0, 0.05); mod.running_var.uniform_(0.5, 1.5)
        m.eval()
    sd = lstm_ref.random_lstm_state_dict(g_dim, g_dim, H, L, seed=14)
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(g_dim, M, seed=14, trained_like=True, smooth_mean=True)
    g = torch.Generator().manual_seed(6)
    x = [torch.rand(B, nc, 64, 64, generator=g) for _ in range(n_eval)]
    eps = {(s, i): torch.randn(g_dim, B, generator=g) for s in range(S) for i in range(n_eval)}
    with torch.no_grad():
        om = rollout_ref.OracleModels(sd, gp_sd, lik_sd, enc, dec, gp_mode="direct")
        ref = rollout_ref.diverse_rollout(om, x, n_past, n_eval, S, eps, resample_every=3)
    fp = make_lstm(sd, rows=B, variant="bf16x3")
    gp, lik = make_gp(gp_sd, lik_sd)
    enc_g, dec_g = enc.cuda(), dec.cuda()
    xg = [t.cuda() for t in x]
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        codec = BatchedCodec(enc_g, dec_g, n_points=B, chunk_rows=2 * B)      # 2 chunks of the 12 rows
        got = diverse_rollout(fp, gp, lik, enc_g, dec_g, xg, n_past, n_eval, S, eps=eps, resample_every=3, codec=codec)
        for t in range(n_eval):
            for s in range(S):
                assert relerr(got[t][s], ref[s][t]) < 2e-3, (t, s)
        hits = resample_steps(n_past, n_eval, 3)
        assert hits == [3, 6]
        eps_dev = torch.stack([torch.stack([eps[(s, i)] for s in range(S)]) for i in hits]).cuda()
        for graph in (False, True):
            pr = PixelRollout(fp, gp, lik, enc_g, dec_g, (nc, 64, 64), B, S, n_past, n_eval, resample_every=3, graph=graph)
            for rep in range(2):                                 # a replay must not depend on leftover state
                frames = pr.run(xg, eps_dev).clone()
                for t in range(n_past, n_eval):
                    for s in range(S):
                        assert relerr(frames[t - n_past].view(S, B, nc, 64, 64)[s], ref[s][t]) < 2e-3, (graph, rep, t, s)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
        enc.cpu(); dec.cpu()


@pytest.mark.parametrize("T,S,B,Gd", [(39, 100, 50, 90), (5, 3, 7, 90), (4, 9, 5, 10), (3, 2, 3, 128), (6, 5, 5, 66)])
def test_score_rollouts_shapes(T, S, B, Gd):
    """Streaming scoring kernel: row counts that are not multiples of 8, latent sizes not multiples of 4."""
    from dvg_b200.rollout import score_rollouts
    g = torch.Generator().manual_seed(T * 100 + S)
    out = torch.randn(T, S * B, Gd, generator=g).cuda()
    target = torch.randn(T, B, Gd, generator=g).cuda()
    got = score_rollouts(out, target, S, B)
    want = (out.view(T, S, B, Gd) - target.view(T, 1, B, Gd)).double().pow(2).mean(dim=(0, 3))
    assert relerr(got, want) < 1e-5
