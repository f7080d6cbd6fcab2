"""Eager (no CUDA graph) hot-path steps for ncu: `ncu ... python scripts/profile_step.py --steps 4`."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, build_models, synth_latents  # noqa: E402
from dvg_b200.rollout import RolloutConfig, RolloutEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=16)
ap.add_argument("--variant", default="bf16x3")
ap.add_argument("--workload", default="kth_s100")
ap.add_argument("--seed", type=int, default=1)
ap.add_argument("--bench-like", action="store_true", help="warm-up = trigger window, seed 100: the rollout bench.py times")
a = ap.parse_args()
w = WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
fp, gp, lik = build_models(w, dev, a.variant)
eng = RolloutEngine(fp, gp, lik, RolloutConfig(n_points=w["B"], n_rollouts=w["S"], window=w["window"], variant=a.variant))
R = w["B"] * w["S"]
lat, eps = synth_latents(w, a.steps, R, dev, 100 if a.bench_like else a.seed)
lat, eps = lat.to(dev), eps.to(dev)
out = torch.empty(a.steps, R, w["G"], device=dev)
with torch.no_grad():
    masks = torch.zeros(a.steps, w["S"], dtype=torch.uint8, device=dev)
    eng.latent_rollout(lat, eps, out, warmup_steps=w["window"] if a.bench_like else a.steps // 2, masks=masks)
torch.cuda.synchronize()
print("done", a.steps, "fired per step:", masks.sum(1).tolist())
