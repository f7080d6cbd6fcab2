"""Golden vectors for the frame metrics from the REAL reference functions (run in the build container only).

``/root/reference/utils.py`` cannot be imported (scipy.misc, skimage, matplotlib are absent) but finn_psnr,
fspecial_gauss and finn_ssim are self-contained: their definitions are extracted with ``ast`` and executed in a
namespace that provides numpy / scipy.signal -- no reference source is copied into the repo.
"""
import ast
import os

import numpy as np
import torch
from scipy import signal

HERE = os.path.dirname(os.path.abspath(__file__))
src = open("/root/reference/utils.py").read()
tree = ast.parse(src)
wanted = {"finn_psnr", "fspecial_gauss", "finn_ssim"}
mod = ast.Module(body=[n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in wanted], type_ignores=[])
ns = {"np": np, "signal": signal}
exec(compile(mod, "reference_utils_subset", "exec"), ns)

g = torch.Generator().manual_seed(7)
cases = []
for (H, note) in ((64, "noise"), (64, "smooth"), (128, "smooth"), (32, "smooth")):
    if note == "noise":
        a = torch.rand(H, H, generator=g)
        b = torch.rand(H, H, generator=g)
    else:
        yy, xx = torch.meshgrid(torch.linspace(0, 1, H), torch.linspace(0, 1, H), indexing="ij")
        a = (0.5 + 0.5 * torch.sin(6 * xx + 3 * yy)).float()
        b = (a + 0.05 * torch.randn(H, H, generator=g)).clamp(0, 1)
    ssim_map = ns["finn_ssim"](a, b)
    cases.append({"a": a, "b": b, "ssim_mean": float(ssim_map.mean()),
                  "psnr": float(ns["finn_psnr"](a.numpy().astype(np.float64), b.numpy().astype(np.float64)))})
torch.save(cases, os.path.join(HERE, "metrics_finn.pt"))
for c in cases:
    print(c["a"].shape, c["ssim_mean"], c["psnr"])
