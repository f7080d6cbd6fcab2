"""Per-step device time of the hot path by step kind (CUDA-graph replay of N identical steps / N):
plain LSTM step, trigger warm-up step, trigger decision step (nothing fires: constant latents).

    python scripts/step_time.py [--workload kth_s100] [--variant bf16x3] [--steps 24] [--reps 20]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, build_models, synth_latents  # noqa: E402
from dvg_b200.rollout import RolloutConfig, RolloutEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=24)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--variant", default="bf16x3")
ap.add_argument("--workload", default="kth_s100")
ap.add_argument("--tag", default="")
a = ap.parse_args()
w = WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
fp, gp, lik = build_models(w, dev, a.variant)
eng = RolloutEngine(fp, gp, lik, RolloutConfig(n_points=w["B"], n_rollouts=w["S"], window=w["window"], variant=a.variant))
R = w["B"] * w["S"]
N = a.steps
lat, eps = synth_latents(w, N, R, dev, 1)
lat = lat.to(dev)
lat[:] = lat[0]                 # constant latents: the variance statistic never moves, nothing fires
eps = eps.to(dev)
out = torch.empty(N, R, w["G"], device=dev)


def timed(fn, label):
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
    best, tot = 1e9, 0.0
    for _ in range(a.reps):
        e0.record()
        g.replay()
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        best = min(best, ms)
        tot += ms
    return {"kind": label, "us_per_step_best": best * 1e3 / N, "us_per_step_mean": tot / a.reps * 1e3 / N}


def plain():
    with eng.chained():
        for t in range(N):
            eng.step_manual_mode(lat[t], None, out[t], resample=False)


def warm():
    eng.reset()
    with eng.chained():
        for t in range(N):
            eng.step_trigger_mode(lat[t], eps[t], out[t], warmup=True)


def decide():
    with eng.chained():
        for t in range(N):
            eng.step_trigger_mode(lat[t], eps[t], out[t], warmup=False)


res = {"workload": a.workload, "variant": a.variant, "rows": R, "tag": a.tag, "steps": []}
res["steps"].append(timed(plain, "plain_lstm_step"))
res["steps"].append(timed(warm, "trigger_warmup_step"))
eng.reset()
with torch.no_grad():
    for t in range(w["window"]):
        eng.step_trigger_mode(lat[t], eps[t], out[t], warmup=True)
res["steps"].append(timed(decide, "trigger_decision_step"))
res["fired_last"] = int(eng.mask.sum().item())
print(json.dumps(res))
