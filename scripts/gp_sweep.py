"""GP trigger scaling sweep (BASELINE configs[4]): predictive variance + trigger decision for `samples` rollouts, D = 90,
inducing set M in {128 .. 4096}.  Prints one JSON line per (M, samples) with the device time of dvg_gp_trigger and the
achieved FP32 rate of the tiled kernel (2 * 2 M^2 flops per (sample, dim), triangular halves counted as stored).
Usage: python scripts/gp_sweep.py [--M 128,512,2048] [--samples 16,256,4096]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dvg_b200 import _capi  # noqa: E402
from dvg_b200.models.gp_models import GaussianLikelihood, GPRegressionLayer1  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--M", default="128,256,512,1024,2048,4096")
ap.add_argument("--samples", default="16,256,4096")
ap.add_argument("--D", type=int, default=90)
a = ap.parse_args()
dev = torch.device("cuda", 0)
D, W = a.D, 12
for M in [int(v) for v in a.M.split(",")]:
    torch.manual_seed(M)
    gp = GPRegressionLayer1(D, M)
    with torch.no_grad():
        gp.variational_strategy.variational_distribution.variational_mean.normal_(0, 0.3)
        gp.variational_strategy.variational_distribution.chol_variational_covar.copy_(
            torch.tril(0.5 * torch.eye(M) + 0.05 * torch.randn(D, M, M) / M ** 0.5))
    gp = gp.to(dev).eval()
    lik = GaussianLikelihood(D).to(dev).eval()
    rt = gp._runtime(lik)
    for S in [int(v) for v in a.samples.split(",")]:
        x = torch.tanh(torch.randn(S, D, device=dev))
        rows = torch.arange(S, dtype=torch.int32, device=dev)
        window = torch.zeros(S, W, device=dev)
        count = torch.zeros(1, dtype=torch.int32, device=dev)
        value, thr = torch.empty(S, device=dev), torch.empty(S, device=dev)
        mask = torch.empty(S, dtype=torch.uint8, device=dev)

        def call(warm):
            _capi.check(rt.lib.dvg_gp_trigger(rt.handle, S, _capi.ptr(x), D, _capi.ptr(rows), _capi.ptr(window), W,
                                              _capi.ptr(count), warm, 2.01, _capi.ptr(value), _capi.ptr(thr),
                                              _capi.ptr(mask), _capi.stream_ptr()), "dvg_gp_trigger")
        for _ in range(W):
            call(1)
        torch.cuda.synchronize()
        reps = 3 if M * M * S > 1 << 32 else 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call(0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        flops = 2.0 * (M * M + 2 * M) * S * D * 2 / 2           # lower + upper triangles, 2 flops per FMA
        print(json.dumps({"M": M, "samples": S, "D": D, "ms": round(ms, 4), "samples_per_s": round(S / ms * 1e3),
                          "fp32_tflops": round(flops / ms / 1e9, 2), "path": "tensor-core tiles" if M > 64 else "smem"}), flush=True)
    del gp, lik, rt
    torch.cuda.empty_cache()
