"""Throughput of the on-device bouncing-digit generator (dvg_moving_mnist): frames/s and achieved HBM write bandwidth
(4 B per pixel is the algorithmic traffic; the digit bank and trajectories stay in cache).

    python scripts/data_bench.py [--out gpurun_out/data_bench.json]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dvg_b200.data import draws_per_seq, moving_mnist_batch, synthetic_digit_bank  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/data_bench.json")
    args = ap.parse_args()
    peaks = {}
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    bank = synthetic_digit_bank(64, seed=0).cuda()
    rows = []
    for B, T, W in [(16, 15, 64), (1600, 15, 64), (5000, 15, 64), (5000, 40, 64)]:
        draws = torch.randint(-2 ** 31, 2 ** 31 - 1, (B, draws_per_seq(T)), device="cuda").to(torch.int32)
        out = torch.empty(T, B, 1, W, W, device="cuda")
        for _ in range(3):
            moving_mnist_batch(bank, B, T, W, draws=draws, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            moving_mnist_batch(bank, B, T, W, draws=draws, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        nbytes = out.numel() * 4
        row = {"n_seq": B, "n_frames": T, "image_size": W, "ms": round(ms, 4), "frames_per_s": round(B * T / ms * 1e3),
               "write_gbs": round(nbytes / ms / 1e6, 1), "bytes": nbytes}
        if "hbm_gbs" in peaks:
            row["frac_of_hbm_peak"] = round(row["write_gbs"] / peaks["hbm_gbs"], 3)
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump(rows, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
