"""Pixel-space end-to-end throughput of the N-diverse-futures rollout (encoder + hot path + decoder,
stock PyTorch convolutions) -- generated frames/s = S * B * n_future / time of one ``diverse_rollout``.

Arms (same models, same inputs, all on one GPU):
  plain       reference-shaped conv execution: eval-mode nets as they are (BatchNorm separate, NCHW, skips replicated
              S times), S samples batched, hot path through the C-ABI engine
  codec       dvg_b200.codec.BatchedCodec fp32 (cuDNN TF32 as torch's default allows): folded BN, channels-last,
              chunks, shared-skip decoder
  codec_bf16  the same in bf16
  graph       codec (fp32) captured into one CUDA graph (PixelRollout)
  graph_bf16  codec_bf16 captured
Each arm is wrapped in try/except so one failure does not lose the others.  Not a bench.py contract line: the
contract metric is the hot path; this explains where end-to-end time goes.

    python scripts/pixel_bench.py [--workloads smmnist_b16 kth_s100] [--reps 2] [--out gpurun_out/pixel_bench.json]
"""
import argparse
import json
import os
import sys
import time
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (model builders / workload table only)

PIXEL = {
    # name: (codec model, channels, width, n_past, n_eval, S override or None)
    "smmnist_b16": ("dcgan_64", 1, 64, 5, 15, None),
    "smmnist_s100": ("dcgan_64", 1, 64, 5, 15, 100),
    "kth_s100": ("vgg_64", 1, 64, 10, 40, None),
    "bair_s32": ("vgg_64", 3, 64, 2, 30, None),
}


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", nargs="+", default=["smmnist_b16", "kth_s100"])
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--arms", nargs="+", default=["plain", "codec", "codec_bf16", "graph", "graph_bf16"])
    ap.add_argument("--out", default="gpurun_out/pixel_bench.json")
    ap.add_argument("--cudnn-benchmark", action="store_true", help="torch.backends.cudnn.benchmark = True (autotuned convs)")
    args = ap.parse_args()
    from dvg_b200.codec import BatchedCodec
    from dvg_b200.convnets import make_codec
    from dvg_b200.rollout import PixelRollout, RolloutConfig, RolloutEngine, diverse_rollout, resample_steps
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = args.cudnn_benchmark
    results = []
    for name in args.workloads:
        model, nc, width, n_past, n_eval, s_over = PIXEL[name]
        w = dict(bench.WORKLOADS[name.replace("smmnist_s100", "smmnist_b16")])
        S = s_over if s_over is not None else w["S"]
        B = w["B"]
        fp, gp, lik = bench.build_models(w, dev, "bf16x3")
        torch.manual_seed(1)
        enc, dec = make_codec(model, w["G"], nc)
        enc, dec = enc.to(dev).eval(), dec.to(dev).eval()
        g = torch.Generator().manual_seed(2)
        x = [torch.rand(B, nc, width, width, generator=g).to(dev) for _ in range(n_eval)]
        hits = resample_steps(n_past, n_eval, 15)
        eps_dev = torch.randn(max(1, len(hits)), S, w["G"], B, generator=g).to(dev)
        frames = S * B * (n_eval - n_past)
        row = {"workload": name, "codec": model, "B": B, "S": S, "rows": S * B, "n_past": n_past, "n_eval": n_eval,
               "frames_per_rollout": frames, "cudnn_benchmark": args.cudnn_benchmark, "arms": {}}
        eng = RolloutEngine(fp, gp, lik, RolloutConfig(n_points=B, n_rollouts=S))
        fbuf = torch.empty(n_eval - min(n_past, n_eval), S * B, nc, width, width, device=dev)

        def arm_plain():
            diverse_rollout(fp, gp, lik, enc, dec, x, n_past, n_eval, S, engine=eng, eps_dev=eps_dev, frames_out=fbuf)

        def make_codec_arm(dtype):
            codec = BatchedCodec(enc, dec, n_points=B, dtype=dtype)
            return lambda: diverse_rollout(fp, gp, lik, enc, dec, x, n_past, n_eval, S, codec=codec, engine=eng,
                                           eps_dev=eps_dev, frames_out=fbuf)

        def make_graph_arm(dtype):
            pr = PixelRollout(fp, gp, lik, enc, dec, (nc, width, width), B, S, n_past, n_eval, codec_dtype=dtype, graph=True)
            pr.x.copy_(torch.stack(x[:pr.n_ctx]))
            pr.eps.copy_(eps_dev)
            return lambda: pr.run()

        makers = {"plain": lambda: arm_plain, "codec": lambda: make_codec_arm(torch.float32),
                  "codec_bf16": lambda: make_codec_arm(torch.bfloat16), "graph": lambda: make_graph_arm(torch.float32),
                  "graph_bf16": lambda: make_graph_arm(torch.bfloat16)}
        for arm in args.arms:
            t0 = time.time()
            try:
                fn = makers[arm]()
                ms = timed(fn, args.reps)
                row["arms"][arm] = {"ms_per_rollout": round(ms, 3), "frames_per_s": round(frames / ms * 1e3, 1),
                                    "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}
            except Exception as e:                               # noqa: BLE001
                row["arms"][arm] = {"error": f"{type(e).__name__}: {e}"[:400]}
                traceback.print_exc()
            torch.cuda.synchronize()
            torch.cuda.reset_peak_memory_stats()
            print(name, arm, row["arms"][arm], f"({time.time() - t0:.1f} s wall)", flush=True)
            del fn
            torch.cuda.empty_cache()
        # hot path alone on the same rows (latent space): (n_eval - n_ctx) manual-mode steps
        try:
            lat = torch.tanh(torch.randn(S * B, w["G"], device=dev))
            out = torch.empty(S * B, w["G"], device=dev)

            def hot():
                for i in range(min(n_past, n_eval), n_eval):
                    hit = i in hits
                    eng.step_manual_mode(lat, eps_dev[0] if hit else None, out, resample=hit)
            row["hot_path_ms"] = round(timed(hot, 5), 4)
        except Exception as e:                                   # noqa: BLE001
            row["hot_path_ms"] = f"{type(e).__name__}: {e}"[:200]
        results.append(row)
        print(json.dumps(row), flush=True)
        del eng, fbuf, fp, gp, lik, enc, dec
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
