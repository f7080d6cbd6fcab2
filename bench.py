#!/usr/bin/env python
"""Benchmark of the DVG stochastic-rollout hot path (BASELINE.json metric: generated frames/sec for N
diverse futures, plus roofline fraction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--variant bf16x3|bf16|fp32]

A "step" is one complete rollout of the hot path over one batch of synthetic encoder latents:
workload ``kth_s100`` = BASELINE.json configs[1] (KTH-shaped: g_dim 90, rnn_size 256, 2 LSTM layers, GP with
40 inducing points, B=50 sequences x S=100 diverse futures = 5000 rows, 10 past + 30 future frames ->
39 time steps, trigger window 12).  Every time step runs: GP variance trigger -> fused LSTM step (state
held for triggered rollouts) -> masked GP rsample.  The encoder/decoder convolutions stay on the stock
PyTorch path (north_star) and are NOT part of the timed hot path; the latents they would produce are
synthetic tensors resident in HBM.  frames = S * B * n_future per rollout.

JSON keys beyond the base contract:
``roofline``          dominant kernel = ``lstm_step_kernel`` (ONE persistent tcgen05 launch per time step: x-pack, all
                      LSTM layers, head, GP trigger, restore/resample of fired rollouts); launch duration = CUDA-graph
                      replay of the T step launches of one rollout / T, CUDA events on the launching stream.  The launches of
                      a rollout are CHAINED (``config.step_launches``): all inputs of the latent-space rollout exist up front,
                      so step t+1 starts on the SMs step t no longer needs; ``lstm_step_ms_stream_ordered`` is the same
                      sequence without the chain (what a caller with the conv nets between the steps gets)
``cpu_baseline``      the reference's CPU path on the host cores, bounded sample (LSTM stage: the reference's own
                      ``models/lstm.py`` when present under baseline/_ref or /root/reference, GP stage: oracle port)
``e2e``               the same metric through the public Python API with pinned HOST buffers, H2D + D2H in the timed region
``stock_torch_b200``  BASELINE.md section 3's practical bar: the reference ``lstm`` module with stock PyTorch ops (cuBLAS sgemm +
                      ATen LSTM-cell pointwise) ON THE SAME B200, same rows / steps, allow_tf32 off and on, eager and as a CUDA
                      graph, next to our plain LSTM step (no trigger) timed the same way
``pixel_e2e``         SURVEY 8d (i): generated frames/s of the same workload in PIXEL space (stock cuDNN encoder / decoder +
                      the hot path, ``PixelRollout``), with the hot path's share of that time
``gpu_launches``, ``clocks``.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: g_dim, rnn_size, layers, inducing, B, S, n_past, n_future, window
    "kth_s100": dict(G=90, H=256, L=2, M=40, B=50, S=100, n_past=10, n_future=30, window=12),
    "smmnist_b16": dict(G=90, H=256, L=2, M=40, B=16, S=1, n_past=5, n_future=10, window=5),
    "bair_s32": dict(G=90, H=256, L=2, M=40, B=50, S=32, n_past=2, n_future=28, window=12),
    "ucf_s100": dict(G=90, H=256, L=2, M=40, B=64, S=100, n_past=5, n_future=25, window=12),
}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
# capture (profiles/r02_lstm_step.md, profiles/r02_lstm_step_ncu_raw.csv); keyed by (workload, variant).
NCU_TRAFFIC_BYTES = {("kth_s100", "bf16x3"): 34.0e6}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def build_models(w, device, variant, seed=1):
    from dvg_b200.models.gp_models import GaussianLikelihood, GPRegressionLayer1
    from dvg_b200.models.lstm import lstm
    from dvg_b200.init import init_gp_state_dicts, init_lstm_state_dict
    fp = lstm(w["G"], w["G"], w["H"], w["L"], w["B"])
    fp.load_state_dict(init_lstm_state_dict(w["G"], w["G"], w["H"], w["L"], seed))
    gp = GPRegressionLayer1(w["G"], w["M"])
    lik = GaussianLikelihood(w["G"])
    gsd, lsd = init_gp_state_dicts(w["G"], w["M"], seed)
    gp.load_state_dict(gsd)
    lik.load_state_dict(lsd)
    fp, gp, lik = fp.to(device).eval(), gp.to(device).eval(), lik.to(device).eval()
    fp.gemm_variant = variant
    return fp, gp, lik


def synth_latents(w, T, R, device, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    scale = torch.linspace(0.4, 1.2, T).reshape(T, 1, 1)
    lat = torch.tanh(torch.randn(T, R, w["G"], generator=g) * scale)      # encoder outputs end in tanh
    # bf16-representable values (still fp32 tensors): the end-to-end arm can then ship them host -> device at half
    # width without changing a single bit of what the kernels compute
    lat = lat.bfloat16().float()
    eps = torch.randn(T, R // w["B"], w["G"], w["B"], generator=g)
    return lat, eps


def launches_per_rollout(w, T):
    # per time step ONE persistent kernel (GP trigger + whole LSTM step + restore/resample of fired rollouts), plus one
    # scoring kernel per rollout
    return T + 1


def flops_bytes(w, R):
    G, H, L = w["G"], w["H"], w["L"]
    f_row = 2 * (G * H + L * (2 * H) * (4 * H) + H * G)
    b_row = 4 * (2 * (2 * L * H) + 2 * G)
    return f_row, b_row, 2 * R * (2 * H) * (4 * H)   # last: one LSTM-layer GEMM launch


# ---------------------------------------------------------------------------------------------------------
def reference_lstm_module():
    """The reference's own ``models/lstm.py`` (unmodified file), from baseline/_ref (written by
    ``__graft_entry__.build()`` where /root/reference exists; git-ignored, travels to the GPU box) or /root/reference.
    Returns (module, where) or (None, why)."""
    import importlib.util
    for root in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        path = os.path.join(root, "models", "lstm.py")
        if os.path.exists(path):
            spec = importlib.util.spec_from_file_location("_dvg_reference_models_lstm", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod, path
    return None, "models/lstm.py not found under baseline/_ref or /root/reference"


def graph_time_us(fn, n_steps, reps=10):
    """CUDA-graph replay of ``fn`` (which enqueues n_steps steps) timed with CUDA events: us per step (best, mean)."""
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    with torch.cuda.graph(g), torch.no_grad():
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best, tot = 1e30, 0.0
    for _ in range(reps):
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        best, tot = min(best, ms), tot + ms
    return best * 1e3 / n_steps, tot / reps * 1e3 / n_steps


def stock_torch_b200(w, R, T, device, eng, lat):
    """BASELINE.md section 3 "also measure": the reference ``lstm`` (models/lstm.py:42-72) with stock PyTorch ops on
    this GPU -- embed addmm, 2 x (2 sgemm + ATen fused LSTM-cell pointwise), output addmm + tanh -- at the bench's
    rows and time steps, against our plain LSTM step (no trigger, same rows), both timed as CUDA-graph replays of T
    steps; the stock module also eagerly (what a user gets with no effort: launch overhead included)."""
    from dvg_b200.init import init_lstm_state_dict
    res = {"rows": R, "time_steps": T, "what": "LSTM step only (the reference's GP stage needs gpytorch, absent here)"}
    mod, where = reference_lstm_module()
    sd = init_lstm_state_dict(w["G"], w["G"], w["H"], w["L"], 1)
    if mod is not None:
        m = mod.lstm(w["G"], w["G"], w["H"], w["L"], R)          # the constructor calls .cuda() (models/lstm.py:61-62)
        res["source"] = "reference models/lstm.py, unmodified (%s)" % where
    else:
        import torch.nn as nn

        class _Restated(nn.Module):                               # same modules and call order as models/lstm.py:42-72
            def __init__(s_, G, H, L):
                super().__init__()
                s_.embed = nn.Linear(G, H)
                s_.lstm = nn.ModuleList([nn.LSTMCell(H, H) for _ in range(L)])
                s_.output = nn.Sequential(nn.Linear(H, G), nn.Tanh())
                s_.n_layers, s_.hidden_size = L, H

            def init_hidden(s_):
                return [(torch.zeros(R, s_.hidden_size, device=device), torch.zeros(R, s_.hidden_size, device=device))
                        for _ in range(s_.n_layers)]

            def forward(s_, x):
                h_in = s_.embed(x.view(-1, x.shape[-1]))
                for i in range(s_.n_layers):
                    s_.hidden[i] = s_.lstm[i](h_in, s_.hidden[i])
                    h_in = s_.hidden[i][0]
                return s_.output(h_in)
        m = _Restated(w["G"], w["H"], w["L"])
        res["source"] = "restated torch modules (%s)" % where
    m.load_state_dict(sd)
    m = m.to(device).eval()
    x = lat[:T]

    def stock_steps(init=True):
        if init:
            m.hidden = m.init_hidden()          # (zeros on the host + .cuda(): not capturable, done outside the graph)
        for t in range(T):
            m(x[t])

    def stock_steps_graphed():
        m.hidden = h0                           # static initial state: every replay starts from the same tensors
        stock_steps(init=False)

    old_tf32 = torch.backends.cuda.matmul.allow_tf32
    try:
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                for _ in range(2):
                    stock_steps()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    stock_steps()
                e1.record()
                torch.cuda.synchronize()
            res["eager_us_per_step_" + name] = e0.elapsed_time(e1) * 1e3 / (3 * T)
            h0 = m.init_hidden()
            res["graph_us_per_step_" + name] = graph_time_us(stock_steps_graphed, T)[0]
        try:
            from torch.profiler import ProfilerActivity, profile
            torch.backends.cuda.matmul.allow_tf32 = False
            with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
                stock_steps()
                torch.cuda.synchronize()
            n_k = sum(e.count for e in prof.key_averages() if getattr(e, "device_type", None) is not None
                      and "cuda" in str(e.device_type).lower())
            res["launches_per_step"] = round(n_k / T, 2)
        except Exception as e:                                    # noqa: BLE001
            res["launches_per_step"] = "profiler unavailable: %s" % type(e).__name__
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old_tf32
    out = torch.empty(T, R, w["G"], device=device)

    def ours_steps(chained=True):
        eng.reset()
        with (eng.chained() if chained else contextlib.nullcontext()):
            for t in range(T):
                eng.step_manual_mode(x[t], None, out[t], resample=False)

    # both arms replay back-to-back step launches with all inputs resident; ours may then chain the launches
    # (include/dvg_b200.h: dvg_lstm_chain_begin) -- the stream-ordered figure is what a caller with the conv nets between
    # the steps gets
    res["ours_us_per_step"] = graph_time_us(ours_steps, T)[0]
    res["ours_us_per_step_stream_ordered"] = graph_time_us(lambda: ours_steps(False), T)[0]
    res["ours_launches_per_step"] = 1
    # numerics of the two arms on the same inputs (stock fp32 as the reference)
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        m.hidden = m.init_hidden()
        ys = torch.stack([m(x[t]) for t in range(T)])
        ours_steps()
    torch.backends.cuda.matmul.allow_tf32 = old_tf32
    torch.cuda.synchronize()
    res["ours_vs_stock_fp32_max_rel_err"] = ((out - ys).abs().max() / ys.abs().max()).item()
    res["speedup_vs_stock_fp32_graph"] = res["graph_us_per_step_fp32"] / res["ours_us_per_step"]
    res["speedup_vs_stock_tf32_graph"] = res["graph_us_per_step_tf32"] / res["ours_us_per_step"]
    res["speedup_vs_stock_fp32_eager"] = res["eager_us_per_step_fp32"] / res["ours_us_per_step"]
    eng.reset()
    return res


PIXEL = {   # workload -> (conv nets, channels, width): the reference's encoder / decoder for that data set
    "smmnist_b16": ("dcgan_64", 1, 64), "kth_s100": ("vgg_64", 1, 64), "bair_s32": ("vgg_64", 3, 64),
}


def pixel_e2e(w, name, device, variant, hot_ms_per_rollout):
    """the same workload in pixel space -- context frames in, S*B*n_future generated frames out, the
    reference's conv nets on the stock cuDNN path (sample-batched through dvg_b200.codec.BatchedCodec, bf16, one CUDA
    graph) around the hot path (make_gifs pass B, generate_frames.py:138-178: resample every 15th step)."""
    if name not in PIXEL:
        return {"skipped": "no conv-net table entry for %s" % name}
    from dvg_b200.convnets import make_codec
    from dvg_b200.rollout import PixelRollout
    model, nc, width = PIXEL[name]
    B, S = w["B"], w["S"]
    n_past, n_eval = w["n_past"], w["n_past"] + w["n_future"]
    fp, gp, lik = build_models(w, device, variant)
    torch.manual_seed(1)
    enc, dec = make_codec(model, w["G"], nc)
    enc, dec = enc.to(device).eval(), dec.to(device).eval()
    g = torch.Generator().manual_seed(2)
    x = torch.rand(n_past, B, nc, width, width, generator=g).to(device)
    pr = PixelRollout(fp, gp, lik, enc, dec, (nc, width, width), B, S, n_past, n_eval, variant=variant,
                      codec_dtype=torch.bfloat16, graph=True)
    pr.x.copy_(x[:pr.n_ctx])
    pr.eps.normal_()
    pr.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    e0.record()
    for _ in range(reps):
        pr.run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    frames = S * B * w["n_future"]
    res = {"value": frames / (ms * 1e-3), "unit": "frames/s", "ms_per_rollout": ms, "frames_per_rollout": frames,
           "conv_nets": model + " (reference architecture, random init, eval; stock cuDNN via BatchedCodec bf16, channels-last, "
                        "BatchNorm folded, skip half of the decoder convs shared across the S futures)",
           "frame": "%dx%dx%d" % (nc, width, width), "cuda_graph": True,
           "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2),
           "hot_path_share": hot_ms_per_rollout / ms,
           "note": "the convolutions are out of the hot path's scope (north_star: they stay on the PyTorch path) and take "
                   "all but hot_path_share of this time"}
    del pr, fp, gp, lik, enc, dec
    torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from dvg_b200 import _capi
    from dvg_b200 import shard
    from dvg_b200.rollout import LatentRolloutPipeline, RolloutConfig, RolloutEngine, score_rollouts
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa = shard.bind_to_gpu_numa_node(local) if world > 1 else None     # before any pinned allocation
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    w = WORKLOADS[args.workload]
    T = w["n_past"] + w["n_future"] - 1
    S, B = w["S"], w["B"]
    R = S * B
    fp, gp, lik = build_models(w, device, args.variant)
    eng = RolloutEngine(fp, gp, lik, RolloutConfig(n_points=B, n_rollouts=S, window=w["window"], variant=args.variant))
    # every rank draws the SAME synthetic latents (weak scaling with identical per-GPU work): with rank-dependent draws
    # the slowest rank's number of fired rollouts, not the kernels, set the multi-GPU time
    lat_h, eps_h = synth_latents(w, T, R, device, seed=100)
    lat_h, eps_h = lat_h.pin_memory(), eps_h.pin_memory()
    lat, eps = lat_h.to(device), eps_h.to(device)
    lat_h16 = lat_h.bfloat16().pin_memory()       # lossless: the synthetic latents are bf16-representable
    assert torch.equal(lat_h16.float(), lat_h)
    out = torch.empty(T, R, w["G"], device=device)
    masks = torch.zeros(T, S, dtype=torch.uint8, device=device)
    # best-of-N selection stand-in for the SSIM selection (generate_frames.py:185-190): per-rollout latent
    # MSE against the first rollout's context latents, gathered across ranks (the only collective).  The scoring pass
    # (and, on one GPU, the arg-best) is captured in the same CUDA graph as the rollout.
    target = lat[:, :B].clone()
    held = {}

    # The cross-rank selection (ONE all-gather of the [S, B] score matrix + arg-best) is captured INSIDE the rollout
    # graph, in stream order after the last step.  Round 1 ran it on a side stream under the next rollout: an NCCL
    # kernel waiting for a slower peer then held an SM that the persistent step kernel (one CTA per SM, all pairs
    # co-resident) needs, and the device-timed scaling lost 8 % at 8 GPUs.  In stream it costs one small collective
    # (~tens of us) per 1.9 ms rollout and never overlaps a step kernel.
    nccl_in_graph = world > 1 and not os.environ.get("DVG_BENCH_SIDE_STREAM")

    def score_in_graph():
        held["sc"] = score_rollouts(out, target, S, B)       # [S_local, B], one fused pass over `out`
        if world == 1:
            held["best"] = shard.select_best(held["sc"], higher_is_better=False)
        elif nccl_in_graph:
            held["best"] = shard.select_best(shard.gather_scores(held["sc"], world * S), higher_is_better=False)

    graph = eng.capture_latent_rollout(lat, eps, out, masks=masks, post=score_in_graph)

    def select_best():
        sc = score_rollouts(out, target, S, B)
        allsc = shard.gather_scores(sc, world * S)
        return shard.select_best(allsc, higher_is_better=False)

    # N > 1: the score all-gather + arg-best of rollout i run on a side stream, overlapped with rollout i + 1 (the
    # exchange never stalls the compute stream; only a snapshot of the [S, B] score matrix is taken in stream order)
    s_sel = torch.cuda.Stream() if world > 1 else None
    sel = {}

    def one_step():
        graph.replay()
        if world == 1 or nccl_in_graph or os.environ.get("DVG_BENCH_NOSEL"):   # (NOSEL: debug switch, no selection)
            return held.get("best")
        main = torch.cuda.current_stream()
        snap = held["sc"].clone()                            # ordered after the graph on the compute stream
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(s_sel):
            s_sel.wait_event(ev)
            allsc = shard.gather_scores(snap, world * S)     # the only collective: one all-gather of scores
            sel["best"] = shard.select_best(allsc, higher_is_better=False)
            snap.record_stream(s_sel)
        return sel

    def finish_selection():
        if s_sel is not None:
            torch.cuda.current_stream().wait_stream(s_sel)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        one_step()
    finish_selection()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        best = one_step()
    finish_selection()              # e1 is ordered after the last selection
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms.item() / args.steps
    frames_per_step = world * S * B * w["n_future"]
    value = frames_per_step / (ms_per_step * 1e-3)
    n_trig = int(masks.sum().item())

    # ---- e2e: public streaming API (LatentRolloutPipeline) with pinned HOST buffers.  Every rollout copies its
    #      inputs H2D and its results D2H inside the timed region; copies of neighbouring rollouts overlap compute.
    #      Results that leave the device: the trigger masks, the [S, B] score matrix, the best-of-N choice and the
    #      WINNING futures' decoder inputs [B, T, G] -- as in the reference's output stage (generate_frames.py:185-217)
    #      and SURVEY 8e, the non-selected futures never travel.
    def e2e_post(o):
        sc = score_rollouts(o, target, S, B)
        allsc = shard.gather_scores(sc, world * S)
        best = shard.select_best(allsc, higher_is_better=False)
        winners = shard.gather_winners(o.view(T, S, B, w["G"]).permute(1, 2, 0, 3), best, world * S)   # [B, T, G]
        return sc, best, winners

    pipe = LatentRolloutPipeline(eng, T, post=e2e_post, full_output=False, capture_post=(world == 1 or nccl_in_graph))
    # correctness of the pipelined path first (host-injected noise == the device-resident run), untimed
    graph.replay()
    torch.cuda.synchronize()
    ref_out = out.clone()
    _, masks_chk, extra_chk = pipe.result(pipe.submit(lat_h16, eps_h))
    if world == 1:
        b_ref = held["best"].cpu()
        want = ref_out.view(T, S, B, w["G"])[:, b_ref, torch.arange(B)].permute(1, 0, 2)
        assert torch.equal(extra_chk[1], b_ref) and torch.equal(extra_chk[2], want.cpu()), \
            "pipelined e2e result differs from the device-resident run"
    assert torch.equal(masks_chk, masks.cpu()), "pipelined e2e trigger masks differ from the device-resident run"
    for _ in range(3):
        pipe.submit(lat_h16, None)
    pipe.drain()
    sync_all()
    t0 = time.perf_counter()
    e0.record()
    # timed: latents from pinned host memory every step, rsample noise drawn on the device (as gpytorch does)
    tickets = [pipe.submit(lat_h16, None) for _ in range(args.steps)]
    _, masks_last, extra_last = pipe.result(tickets[-1])
    for st in (pipe.s_in, pipe.s_cmp, pipe.s_out):
        torch.cuda.current_stream().wait_stream(st)      # e1 is ordered after all three pipeline streams
    e1.record()
    sync_all()
    wall_ms = (time.perf_counter() - t0) * 1e3
    ms2 = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = frames_per_step / (ms2.item() / args.steps * 1e-3)
    h2d = lat_h16.numel() * 2
    d2h = pipe.d2h_bytes()

    # ---- roofline: time the dominant kernel (LSTM layer GEMM) live, events between launches ----
    roof = None
    if rank == 0:
        roof = measure_roofline(eng, w, R, lat, args)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_rollout(w, T, budget_s=args.cpu_budget, threads=os.cpu_count())
    stock = pix = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            stock = stock_torch_b200(w, R, T, device, eng, lat)
        except Exception as e:                                    # noqa: BLE001
            stock = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        try:
            del pipe
            torch.cuda.empty_cache()
            pix = pixel_e2e(w, args.workload, device, args.variant, ms_per_step)
        except Exception as e:                                    # noqa: BLE001
            pix = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    if rank == 0:
        f_row, b_row, _ = flops_bytes(w, R)
        line = {
            "metric": "generated frames/sec (N diverse futures), rollout hot path", "value": value,
            "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16x3": "fp32-grade (bf16x3 split products, fp32 accumulate)", "bf16": "bf16",
                      "fp32": "fp32"}[args.variant],
            "data": "synthetic",
            "config": {"workload": args.workload + (" (BASELINE configs[1]: KTH-shaped, B=50 x S=100 futures per GPU, "
                       "10 past + 30 future, g_dim 90, rnn 256x2, GP M=40, trigger window 12)" if args.workload == "kth_s100"
                                                    else " (B=%d x S=%d futures per GPU, %d past + %d future, g_dim %d, rnn %dx%d, GP M=%d, "
                                                    "trigger window %d)" % (B, S, w["n_past"], w["n_future"], w["G"], w["H"], w["L"], w["M"], w["window"])),
                       "rows_per_gpu": R, "time_steps": T, "frames_per_step": frames_per_step,
                       "row_steps_per_s": world * R * T / (ms_per_step * 1e-3),
                       "variant": args.variant, "cuda_graph": True,
                       "step_launches": "chained where the grid covers the GPU (include/dvg_b200.h dvg_lstm_chain_begin: every "
                                        "input of the latent-space rollout exists up front, so consecutive step launches overlap; "
                                        "results identical); stream-ordered figures: roofline.lstm_step_ms_stream_ordered, "
                                        "stock_torch_b200.ours_us_per_step_stream_ordered; DVG_STEP_CHAIN=0 disables",
                       "l2": "inputs+outputs per rollout = %.0f MB > 126 MB L2" % ((lat.numel() + out.numel()) * 4 / 1e6),
                       "triggered_rollout_steps": n_trig, "scope": "hot path only; encoder/decoder convs excluded",
                       "multi_gpu": None if world == 1 else {
                           "collective": "one ncclAllGather of the [S,B] score matrix per rollout (+ one all-reduce of the "
                                         "winners' [B,T,G] decoder inputs in the e2e pipeline), " +
                                         ("captured in the rollout graph, in stream order" if nccl_in_graph else "side stream"),
                           "same_inputs_on_every_rank": True, "numa_binding_rank0": numa}},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms2.item() / args.steps, "wall_ms_per_step": wall_ms / args.steps,
                    "h2d_format": "bf16 on the wire (the synthetic latents are bf16-representable fp32 values: lossless), expanded "
                                  "to fp32 on the device inside the timed region; the pipelined result is asserted equal to the "
                                  "device-resident fp32 run",
                    "how": "LatentRolloutPipeline: pinned host in/out, H2D / compute graph / D2H on three streams, two buffer "
                           "sets each with its own graph (copies of neighbouring rollouts overlap compute); all latents H2D every "
                           "rollout, rsample noise drawn on the device inside the timed region; D2H every rollout: trigger masks, "
                           "score matrix, best-of-N choice and the winning futures' decoder inputs [B,T,G] (non-selected futures "
                           "stay on the device, as in generate_frames.py:185-217)"},
            "gpu_launches": launches_per_rollout(w, T) * args.steps,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "stock_torch_b200": stock,
            "pixel_e2e": pix,
            "hot_path_algorithmic": {"flops_per_row_step": f_row, "bytes_per_row_step": b_row},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # CUDA graphs that captured NCCL work keep the communicator busy: destroy_process_group() was seen to hang
        # for minutes with them alive.  Everything is measured and printed: synchronise, meet once more, leave.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def measure_roofline(eng, w, R, lat, args):
    """Device time of the dominant kernel, ``lstm_step_kernel`` (one launch = one whole LSTM time step of all R rows
    with the GP trigger fused, exactly the launch the timed rollout issues): a CUDA graph holding ONLY the T step
    launches of one rollout (no rsample, no scoring) is replayed and timed with CUDA events on the launching stream;
    average launch duration = replay time / T (it includes the ~1 us launch gaps between dependent kernels, not host
    launch overhead).  The other variants are timed the same way."""
    from dvg_b200 import _capi
    from dvg_b200.rollout import RolloutConfig, RolloutEngine
    pk, how = peaks()
    T = lat.shape[0]
    out = torch.empty(T, R, w["G"], device=lat.device)     # every step writes its own rows, as in the rollout

    def step_time(engine, reps, chained=True):
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            engine.reset()
            for t in range(2):
                engine.step_trigger_mode(lat[t], None, out[t], warmup=t < 1, resample=False)
            engine.reset()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            engine.reset()
            with (engine.chained() if chained else contextlib.nullcontext()):   # chained: the rollout's launch sequence
                for t in range(T):
                    engine.step_trigger_mode(lat[t], None, out[t], warmup=t < w["window"], resample=False)
        engine.cur = 0 if T % 2 == 0 else 1
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (reps * T)

    reps = max(3, min(args.steps, 20))
    step_ms = step_time(eng, reps)
    step_ms_so = step_time(eng, 3, chained=False)     # what a caller with other work between the steps gets
    other = {}
    for name in ("bf16x3", "bf16", "fp32"):
        if name == args.variant:
            continue
        fp2, gp2, lik2 = build_models(w, lat.device, name)
        e2 = RolloutEngine(fp2, gp2, lik2, RolloutConfig(n_points=w["B"], n_rollouts=w["S"], window=w["window"], variant=name))
        other[name] = step_time(e2, 3)
        del e2, fp2, gp2, lik2
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    issued = 3 if args.variant == "bf16x3" else 1
    f_row, b_row, _ = flops_bytes(w, R)
    achieved = f_row * R / (step_ms * 1e-3) / 1e12
    hbm = pk["hbm_gbs"]
    return {"bound": "tensor",
            "kernel": "lstm_step_kernel (whole LSTM time step in one persistent launch: x-pack, %d layers with the embed "
                      "folded in, head, GP trigger; %d rows)" % (w["L"], R),
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": "%s bf16 dense, sustained (kernel timed inside a long step)" % how,
            "traffic": NCU_TRAFFIC_BYTES.get((args.workload, args.variant)),
            "traffic_source": "profiles/r02_lstm_step.md (ncu --set full, dram read+write bytes per launch)",
            "algorithmic_bytes_per_launch": b_row * R,
            "tensor_issue_frac": achieved * issued / peak,
            "hbm_frac_of_state_io": (b_row * R / (step_ms * 1e-3) / 1e9) / hbm,
            "note": "achieved counts ALGORITHMIC flops of the reference step, 2*(G*H + L*2H*4H + H*G) per row; the bf16x3 "
                    "variant issues 3 tcgen05.mma per algorithmic MMA (tensor_issue_frac = tensor-pipe work actually "
                    "issued / peak); launch duration = CUDA-graph replay of the T step launches of one rollout / T "
                    "(the launches are chained as in the timed rollout, i.e. they overlap: this is the step's cost in the "
                    "sequence; lstm_step_ms_stream_ordered = the same without the chain)",
            "kernel_ms": {"lstm_step_kernel": step_ms}, "lstm_step_ms": step_ms, "lstm_step_ms_stream_ordered": step_ms_so,
            "other_variants_lstm_step_ms": other}


# ---------------------------------------------------------------------------------------------------------
def cpu_rollout(w, T, budget_s, threads):
    """Oracle port of the same trigger-mode rollout on the host cores.  The reference loops the S samples
    sequentially (generate_frames.py:143), so a bounded number of samples is timed and frames/s reported
    for that sample."""
    from oracle import gp_ref, lstm_ref, trigger_ref
    from dvg_b200.init import init_gp_state_dicts, init_lstm_state_dict
    import numpy as np
    torch.set_num_threads(threads)
    sd = init_lstm_state_dict(w["G"], w["G"], w["H"], w["L"], 1)
    gsd, lsd = init_gp_state_dicts(w["G"], w["M"], 1)
    B, W = w["B"], w["window"]
    # LSTM stage: the reference's own module where its file is available (models/lstm.py:42-72; its constructor and
    # init_hidden call .cuda(), shimmed to a no-op for this CPU run), else the oracle restatement
    ref_mod, where = reference_lstm_module()
    ref_fp = None
    if ref_mod is not None:
        orig_cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            ref_fp = ref_mod.lstm(w["G"], w["G"], w["H"], w["L"], B)
            ref_fp.load_state_dict(sd)
            ref_fp.eval()
        finally:
            pass          # restored after the timed loop (init_hidden calls .cuda() too)
    g = torch.Generator().manual_seed(5)
    lat = torch.tanh(torch.randn(T, B, w["G"], generator=g))
    done, t0 = 0, time.perf_counter()
    with torch.no_grad():
        while True:
            hid = lstm_ref.init_hidden(w["L"], B, w["H"])
            if ref_fp is not None:
                ref_fp.hidden = ref_fp.init_hidden()
            ctx = []
            for t in range(T):
                pred = gp_ref.predictive(gsd, lsd, gp_ref.latent_to_gp_input(lat[t]), torch.float32, "gpytorch",
                                         full_cov=False)
                v = trigger_ref.trigger_value(pred["variance"].numpy(), min(3, B - 1))
                fired = False
                if t < W:
                    ctx.append(v)
                else:
                    c = trigger_ref.slide(np.array(ctx, dtype=np.float32), v)
                    ctx = list(c)
                    fired = trigger_ref.decide(c, v)
                if fired:
                    pc = gp_ref.predictive(gsd, lsd, gp_ref.latent_to_gp_input(lat[t]), torch.float32, "gpytorch")
                    gp_ref.rsample(pc["mean"], pc["covar"], torch.randn(w["G"], B))
                elif ref_fp is not None:
                    ref_fp(lat[t])
                else:
                    _, hid = lstm_ref.lstm_forward(sd, lat[t], hid)
            done += 1
            el = time.perf_counter() - t0
            if el > budget_s or done >= w["S"]:
                break
    if ref_mod is not None:
        torch.Tensor.cuda = orig_cuda
    return {"value": done * B * w["n_future"] / el, "unit": "frames/s", "cores": threads,
            "kind": "port",       # mixed: LSTM stage = the reference's own module when available, GP stage = port (see below)
            "lstm_stage": ("reference models/lstm.py, unmodified (%s)" % where) if ref_fp is not None else "oracle/lstm_ref.py (%s)" % where,
            "gp_stage": "oracle/gp_ref.py + oracle/trigger_ref.py (port: the reference's GP arithmetic is gpytorch, absent)",
            "sample": "%d of %d samples (sequential in S like generate_frames.py:143), B=%d, %d time steps, %.1f s"
                      % (done, w["S"], B, T, el)}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; /root/reference does
    not exist on the GPU box and gpytorch is not installable) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    T = w["n_past"] + w["n_future"] - 1
    threads = os.cpu_count()
    per_step_budget = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_rollout(w, T, budget_s=min(per_step_budget, 3.0), threads=threads)
    vals, t0 = [], time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_rollout(w, T, budget_s=per_step_budget, threads=threads))
    el = time.perf_counter() - t0
    v = sum(x["value"] for x in vals) / len(vals)
    line = {"impl": "reference", "metric": "generated frames/sec (N diverse futures), rollout hot path",
            "value": v, "unit": "frames/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": args.workload, "scope": "hot path only; the reference's CPU path (LSTM: its own module "
                       "where the file is available, GP: oracle port), bounded sample per step"},
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": threads, "kind": vals[-1]["kind"],
                             "lstm_stage": vals[-1]["lstm_stage"], "gp_stage": vals[-1]["gp_stage"],
                             "sample": vals[-1]["sample"]},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="bf16x3", choices=["bf16x3", "bf16", "fp32"])
    ap.add_argument("--workload", default="kth_s100", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the stock_torch_b200 and pixel_e2e measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    run_ours(args)


if __name__ == "__main__":
    main()
