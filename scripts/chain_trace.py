"""%globaltimer trace of consecutive CHAINED step launches (DVG_TRACE build): eager chain of N trigger steps, the
library dumps launches DVG_TC_TRACE_LAUNCH .. +2 at chain end (stderr); this script prints per-launch summaries on a
common time base.
    DVG_TRACE=1 DVG_LIB_TAG=trace DVG_TC_TRACE=1 python scripts/chain_trace.py [--kind warm|plain|decide] 2> trace.log"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, build_models, synth_latents  # noqa: E402
from dvg_b200.rollout import RolloutConfig, RolloutEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="warm")
ap.add_argument("--workload", default="kth_s100")
ap.add_argument("--steps", type=int, default=12)
ap.add_argument("--first", type=int, default=6)
a = ap.parse_args()
w = WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
fp, gp, lik = build_models(w, dev, "bf16x3")
eng = RolloutEngine(fp, gp, lik, RolloutConfig(n_points=w["B"], n_rollouts=w["S"], window=w["window"], variant="bf16x3"))
R = w["B"] * w["S"]
N = a.steps
lat, eps = synth_latents(w, N, R, dev, 1)
lat = lat.to(dev)
lat[:] = lat[0]
eps = eps.to(dev)
out = torch.empty(N, R, w["G"], device=dev)
def chain():
    with eng.chained():
        for t in range(N):
            if a.kind == "plain":
                eng.step_manual_mode(lat[t], None, out[t], resample=False)
            else:
                eng.step_trigger_mode(lat[t], eps[t], out[t], warmup=a.kind == "warm")


with torch.no_grad():
    eng.reset()
    if a.kind == "decide":
        for t in range(w["window"]):
            eng.step_trigger_mode(lat[0], eps[0], out[0], warmup=True)
    torch.cuda.synchronize()
    # the launches are captured in a CUDA graph (eager launches from Python are slower than the kernels: the chain would
    # be launch bound); the per-launch trace slabs are baked into the graph and dumped by the next eager chain end
    os.environ["DVG_TC_TRACE_LAUNCH"] = "99"
    chain()                       # warm (dump suppressed: launch 99 does not exist)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            chain()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    os.environ["DVG_TC_TRACE_LAUNCH"] = str(a.first)
    with eng.chained():           # no launches: the chain end dumps what the last replay recorded
        pass
    torch.cuda.synchronize()
print("done")
