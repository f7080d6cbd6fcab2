"""Once-per-weight-load factorisation of the GP constants for large inducing sets: the library's blocked fp64 kernels
(dvg_gp_factorize) next to torch.linalg (cuSOLVER) on the same GPU.   python scripts/factor_time.py"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dvg_b200 import _capi  # noqa: E402

lib = _capi.load()
dev = torch.device("cuda")
D = 90
for M in (128, 256, 512, 1024, 2048, 4096):
    g = torch.Generator().manual_seed(M)
    z = torch.rand(D, M, generator=g).to(dev)
    m_q = (0.3 * torch.randn(D, M, generator=g)).to(dev)
    c = torch.zeros(D, device=dev)
    raw = torch.zeros(D, device=dev)
    linv = torch.empty(D, M, M, device=dev)
    beta = torch.empty(D, M, device=dev)
    dims = _capi.GpDims(D, M, 1e-3, 1e-4)

    def native():
        batch = max(1, min(D, (1 << 28) // (M * M)))
        ws = torch.empty(lib.dvg_gp_factorize_workspace(_capi.ctypes.byref(dims), batch), dtype=torch.uint8, device=dev)
        _capi.check(lib.dvg_gp_factorize(_capi.ctypes.byref(dims), _capi.ptr(z), _capi.ptr(m_q), _capi.ptr(c), _capi.ptr(raw),
                                         _capi.ptr(raw), _capi.ptr(linv), _capi.ptr(beta), _capi.ptr(ws), ws.numel(),
                                         _capi.stream_ptr()), "dvg_gp_factorize")

    def stock():
        f64 = torch.float64
        ell = torch.nn.functional.softplus(raw.to(f64))
        eye = torch.eye(M, dtype=f64, device=dev)
        step = max(1, min(D, (1 << 28) // (M * M)))
        for d0 in range(0, D, step):
            d1 = min(D, d0 + step)
            zz = z[d0:d1].to(f64)
            t = (zz[:, :, None] - zz[:, None, :]) / ell[d0:d1, None, None]
            K = ell[d0:d1, None, None] * torch.exp(-0.5 * t * t) + 1e-3 * eye
            L = torch.linalg.cholesky(K)
            Li = torch.linalg.solve_triangular(L, eye.expand(d1 - d0, M, M), upper=False)
            linv[d0:d1] = Li.float()
            beta[d0:d1] = torch.einsum("dij,dj->di", Li, m_q[d0:d1].to(f64)).float()

    res = {"M": M, "D": D}
    for name, fn in (("native_ms", native), ("torch_linalg_ms", stock)):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        res[name] = round((time.perf_counter() - t0) * 1e3, 2)
    res["fp64_gflop"] = round(D * (M ** 3) * (1 / 3 + 1 / 3) * 2 / 1e9, 1)       # Cholesky + block-row inverse (zero blocks skipped)
    res["native_tflops"] = round(res["fp64_gflop"] / res["native_ms"], 2)
    print(json.dumps(res), flush=True)
