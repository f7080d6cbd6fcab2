"""Warm-cache CUDA-event timing of the individual hot-path calls (eager, no graph)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, build_models, synth_latents
from dvg_b200.rollout import RolloutConfig, RolloutEngine

w = WORKLOADS["kth_s100"]
dev = torch.device("cuda", 0)
fp, gp, lik = build_models(w, dev, "bf16x3")
eng = RolloutEngine(fp, gp, lik, RolloutConfig(n_points=w["B"], n_rollouts=w["S"], window=w["window"]))
R = w["B"] * w["S"]
lat, eps = synth_latents(w, 4, R, dev, 1)
lat, eps = lat.to(dev), eps.to(dev)
out = torch.empty(R, w["G"], device=dev)


def timeit(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


with torch.no_grad():
    print("trigger (decision) us:", timeit(lambda: eng.trigger(lat[0], warmup=False)))
    print("trigger (warmup)   us:", timeit(lambda: eng.trigger(lat[0], warmup=True)))
    print("lstm advance       us:", timeit(lambda: eng.advance(lat[0], out, hold=True)))
    print("rsample (masked)   us:", timeit(lambda: eng.resample(lat[0], eps[0], out, masked=True)))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        eng.trigger(lat[0], warmup=False)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(20):
            eng.trigger(lat[0], warmup=False)
    print("trigger in graph   us:", timeit(lambda: g.replay(), n=50) / 20)
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):
        for i in range(20):
            eng.step_trigger_mode(lat[i % 4], eps[i % 4], out, warmup=False)
    print("full step in graph us:", timeit(lambda: g2.replay(), n=50) / 20)
