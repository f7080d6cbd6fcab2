"""GPU parity of the GP predictive / trigger / rsample (through the C ABI) against the CPU oracle.
The GP oracle is a restatement of gpytorch 0.3.x (PARITY UNPINNED, see oracle/gp_ref.py); the bar is
<= 1e-4 relative against its fp64 evaluation."""
import numpy as np
import pytest
import torch

from oracle import gp_ref, trigger_ref
from util import make_gp, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("params", ["init", "trained_smooth", "trained_iid"])
@pytest.mark.parametrize("D,M,N", [(90, 40, 50), (90, 40, 333), (17, 24, 9), (8, 64, 40), (5, 128, 30)])
def test_predict(params, D, M, N):
    """mean / variance vs the fp64 oracle.  Variance: <= 1e-4 always.  Mean: <= 1e-4 for the random-init
    and the smooth trained-like sets; for the i.i.d. m_q set, where K_ZZ^-1(m_q - c) is
    ill-conditioned and the reference's own fp32 arithmetic is 4e-4..7e-3 off fp64, the CUDA path must be
    at least as close to fp64 as the fp32 oracle is."""
    trained = params != "init"
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(D, M, seed=D + M, trained_like=trained,
                                                 smooth_mean=params == "trained_smooth")
    gp, lik = make_gp(gp_sd, lik_sd)
    h = torch.tanh(torch.randn(N, D, generator=torch.Generator().manual_seed(N)))
    xin = gp_ref.latent_to_gp_input(h)
    ref = gp_ref.predictive(gp_sd, lik_sd, xin, torch.float64, "direct", full_cov=False)
    ref32 = gp_ref.predictive(gp_sd, lik_sd, xin, torch.float32, "gpytorch", full_cov=False)
    hc = h.cuda()
    with torch.no_grad():
        pred = lik(gp(hc.transpose(0, 1).view(D, N, 1)))
        mean, var = pred.mean, pred.variance
    assert mean.shape == (D, N) and var.shape == (D, N)
    assert relerr(var, ref["variance"]) < 1e-4
    if params == "init":
        assert mean.abs().max().item() < 1e-6
    elif params == "trained_smooth":
        assert relerr(mean, ref["mean"]) < 1e-4
    else:
        assert relerr(mean, ref["mean"]) < max(1e-4, relerr(ref32["mean"], ref["mean"]))
    # latent (noise-free) predictive: gp_layer(x).variance == variance - noise
    with torch.no_grad():
        f = gp(hc.transpose(0, 1).view(D, N, 1))
        _, _, _, noise = gp_ref.effective_hypers(gp_sd, lik_sd, torch.float64)
        assert relerr(f.variance, ref["variance"] - noise.reshape(-1, 1)) < 2e-4
        assert relerr(f.mean, mean) < 1e-6 or params == "init"


def test_prepared_factors():
    from dvg_b200 import _capi
    D, M = 12, 40
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(D, M, seed=3, trained_like=True)
    gp, lik = make_gp(gp_sd, lik_sd)
    rt = gp._runtime(lik)
    linv = torch.empty(D, M, M, device="cuda")
    lqt = torch.empty(D, M, M, device="cuda")
    alpha = torch.empty(D, M, device="cuda")
    hyp = torch.empty(D, 4, device="cuda")
    _capi.check(rt.lib.dvg_gp_export(rt.handle, _capi.ptr(linv), _capi.ptr(lqt), _capi.ptr(alpha), _capi.ptr(hyp),
                                     _capi.stream_ptr()))
    ell, s, c, noise = gp_ref.effective_hypers(gp_sd, lik_sd, torch.float64)
    assert relerr(hyp, torch.stack([ell, s, c, noise], 1)) < 1e-6
    Z = gp_sd[gp_ref.K_INDUCING].double()
    K = gp_ref.kernel(Z, Z, ell, s, "direct") + 1e-3 * torch.eye(M, dtype=torch.float64)
    L = torch.linalg.cholesky(K)
    assert relerr(linv, torch.linalg.inv(L)) < 1e-5
    assert relerr(lqt, torch.tril(gp_sd[gp_ref.K_VCHOL].double()).transpose(1, 2)) < 1e-6
    b = torch.linalg.solve_triangular(L, (gp_sd[gp_ref.K_VMEAN].double() - c.reshape(-1, 1)).unsqueeze(-1),
                                      upper=False).squeeze(-1)
    assert relerr(alpha, b) < 1e-5


@pytest.mark.parametrize("trained", [False, True])
def test_rsample(trained):
    D, M, N = 90, 40, 50
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(D, M, seed=21, trained_like=trained, smooth_mean=True)
    gp, lik = make_gp(gp_sd, lik_sd)
    g = torch.Generator().manual_seed(2)
    h = torch.tanh(torch.randn(N, D, generator=g))
    eps = torch.randn(D, N, generator=g)
    ref = gp_ref.predictive(gp_sd, lik_sd, gp_ref.latent_to_gp_input(h), torch.float64, "direct")
    want = gp_ref.rsample(ref["mean"], ref["covar"], eps.double())
    with torch.no_grad():
        got = lik(gp(h.cuda().transpose(0, 1).view(D, N, 1))).rsample(eps=eps.cuda())
    assert got.shape == (D, N)
    assert relerr(got, want) < 1e-4


@pytest.mark.parametrize("D,M,N", [(5, 40, 1), (7, 24, 2), (3, 40, 63), (3, 40, 64), (2, 64, 100), (2, 40, 128)])
def test_rsample_sizes(D, M, N):
    """Point counts around the register-resident Cholesky's limit (N <= 63 with 256 threads in the stand-alone kernel)
    and up to the documented maximum of 128."""
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(D, M, seed=5 + N, trained_like=True, smooth_mean=True)
    gp, lik = make_gp(gp_sd, lik_sd)
    g = torch.Generator().manual_seed(N)
    h = torch.tanh(torch.randn(N, D, generator=g))
    eps = torch.randn(D, N, generator=g)
    ref = gp_ref.predictive(gp_sd, lik_sd, gp_ref.latent_to_gp_input(h), torch.float64, "direct")
    want = gp_ref.rsample(ref["mean"], ref["covar"], eps.double())
    with torch.no_grad():
        got = lik(gp(h.cuda().transpose(0, 1).view(D, N, 1))).rsample(eps=eps.cuda())
    assert relerr(got, want) < 1e-4


def test_rsample_rejects_more_than_128_points():
    from dvg_b200 import _capi
    D, M, N = 2, 40, 129
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(D, M, seed=1)
    gp, lik = make_gp(gp_sd, lik_sd)
    h = torch.tanh(torch.randn(N, D))
    with torch.no_grad(), pytest.raises(_capi.DvgError):
        lik(gp(h.cuda().transpose(0, 1).view(D, N, 1))).rsample(eps=torch.randn(D, N).cuda())


def test_trigger_sequence_matches_numpy_oracle():
    """Device window / threshold / decision vs oracle/trigger_ref.py over a 60-step sequence for S=7
    independent rollouts (decisions bit-exact outside a 1e-5 band around the threshold)."""
    from dvg_b200 import _capi
    D, M, N, S, W = 90, 40, 50, 7, 12
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(D, M, seed=33, trained_like=True)
    gp, lik = make_gp(gp_sd, lik_sd)
    rt = gp._runtime(lik)
    g = torch.Generator().manual_seed(5)
    T = 60
    lat = torch.tanh(torch.randn(T, S * N, D, generator=g) * torch.linspace(0.3, 1.5, T).reshape(T, 1, 1))
    stat_cols = torch.tensor([3, 0, 7, 49, 3, 21, 10])
    stat_rows = (torch.arange(S) * N + stat_cols).int().cuda()
    window = torch.zeros(S, W, device="cuda")
    count = torch.zeros(1, dtype=torch.int32, device="cuda")
    value = torch.empty(S, device="cuda")
    thr = torch.empty(S, device="cuda")
    mask = torch.empty(S, dtype=torch.uint8, device="cuda")
    ctx = [[] for _ in range(S)]
    n_checked = n_fired = 0
    for t in range(T):
        x = lat[t].cuda()
        warm = 1 if t < W else 0
        _capi.check(rt.lib.dvg_gp_trigger(rt.handle, S, _capi.ptr(x), D, _capi.ptr(stat_rows), _capi.ptr(window), W,
                                          _capi.ptr(count), warm, float(np.float32(trigger_ref.FACTOR)),
                                          _capi.ptr(value), _capi.ptr(thr), _capi.ptr(mask), _capi.stream_ptr()))
        v_gpu, thr_gpu, m_gpu = value.cpu().numpy(), thr.cpu().numpy(), mask.cpu().numpy()
        for s in range(S):
            ref = gp_ref.predictive(gp_sd, lik_sd, gp_ref.latent_to_gp_input(lat[t, s * N:(s + 1) * N]),
                                    torch.float32, "gpytorch", full_cov=False)
            v = trigger_ref.trigger_value(ref["variance"].numpy(), int(stat_cols[s]))
            assert abs(v_gpu[s] - v) <= 1e-4 * abs(v)
            if warm:
                ctx[s].append(v)
                assert m_gpu[s] == 0
            else:
                c = trigger_ref.slide(np.array(ctx[s], dtype=np.float32), v)
                ctx[s] = list(c)
                th = trigger_ref.threshold(c)
                assert abs(thr_gpu[s] - th) <= 1e-4 * abs(th)
                if abs(float(v) - float(th)) > 1e-4 * abs(float(th)):
                    assert bool(m_gpu[s]) == bool(v > th)
                    n_checked += 1
                    n_fired += int(v > th)
        np.testing.assert_allclose(window.cpu().numpy()[:, :len(ctx[0])] if warm else window.cpu().numpy(),
                                   np.array(ctx, dtype=np.float32), rtol=1e-4)
    assert n_checked > 300 and 0 < n_fired < n_checked


def test_rsample_mask_leaves_rows_untouched():
    from dvg_b200 import _capi
    D, M, N, S = 30, 40, 20, 5
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(D, M, seed=8, trained_like=True, smooth_mean=True)
    gp, lik = make_gp(gp_sd, lik_sd)
    rt = gp._runtime(lik)
    g = torch.Generator().manual_seed(1)
    x = torch.tanh(torch.randn(S * N, D, generator=g)).cuda()
    eps = torch.randn(S, D, N, generator=g).cuda()
    out = torch.full((S * N, D), 7.0, device="cuda")
    mask = torch.tensor([1, 0, 0, 1, 0], dtype=torch.uint8, device="cuda")
    _capi.check(rt.lib.dvg_gp_rsample(rt.handle, S, N, _capi.ptr(x), D, _capi.ptr(eps), _capi.ptr(mask),
                                      _capi.ptr(out), D, _capi.stream_ptr()))
    out = out.cpu()
    for s in range(S):
        blk = out[s * N:(s + 1) * N]
        if mask[s]:
            ref = gp_ref.predictive(gp_sd, lik_sd, gp_ref.latent_to_gp_input(x[s * N:(s + 1) * N].cpu()),
                                    torch.float64, "direct")
            want = gp_ref.rsample(ref["mean"], ref["covar"], eps[s].cpu().double()).transpose(0, 1)
            assert relerr(blk, want) < 1e-4
        else:
            assert torch.all(blk == 7.0)


# ---- large inducing sets (M > 64: pre-computed factors + tcgen05 tiles, gp_tc.cu / gp_big.cu; BASELINE configs[4]) -------
@pytest.mark.parametrize("params", ["init", "trained_smooth"])
@pytest.mark.parametrize("D,M,N", [(6, 129, 70), (4, 200, 33), (3, 512, 130)])
def test_predict_large_inducing_set(params, D, M, N):
    """Same bar as test_predict (variance and smooth-set mean <= 1e-4 vs the fp64 oracle) on the large-M path."""
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(D, M, seed=D + M, trained_like=params != "init", smooth_mean=True)
    gp, lik = make_gp(gp_sd, lik_sd)
    h = torch.tanh(torch.randn(N, D, generator=torch.Generator().manual_seed(N)))
    ref = gp_ref.predictive(gp_sd, lik_sd, gp_ref.latent_to_gp_input(h), torch.float64, "direct", full_cov=False)
    hc = h.cuda()
    with torch.no_grad():
        pred = lik(gp(hc.transpose(0, 1).view(D, N, 1)))
        mean, var = pred.mean, pred.variance
    assert gp._runtime(lik).M == M and M > 64
    assert relerr(var, ref["variance"]) < 1e-4
    if params == "init":
        assert mean.abs().max().item() < 1e-6
    else:
        assert relerr(mean, ref["mean"]) < 1e-4


@pytest.mark.parametrize("D,M,N", [(3, 256, 50), (2, 129, 7), (2, 512, 128), (2, 200, 65)])
def test_rsample_large_inducing_set(D, M, N):
    """.rsample() on a handle with pre-computed factors (M > 64, BASELINE configs[4]; gp_big_rsample_kernel):
    the same bar as test_rsample, plus the mask semantics (unmasked rollouts untouched)."""
    from dvg_b200 import _capi
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(D, M, seed=3 + M, trained_like=True, smooth_mean=True)
    gp, lik = make_gp(gp_sd, lik_sd)
    g = torch.Generator().manual_seed(N)
    h = torch.tanh(torch.randn(N, D, generator=g))
    eps = torch.randn(D, N, generator=g)
    ref = gp_ref.predictive(gp_sd, lik_sd, gp_ref.latent_to_gp_input(h), torch.float64, "direct")
    want = gp_ref.rsample(ref["mean"], ref["covar"], eps.double())
    with torch.no_grad():
        got = lik(gp(h.cuda().transpose(0, 1).view(D, N, 1))).rsample(eps=eps.cuda())
    assert gp._runtime(lik).M == M and M > 64
    assert relerr(got, want) < 1e-4
    # three rollouts, only the middle one masked
    rt = gp._runtime(lik)
    S = 3
    x = torch.tanh(torch.randn(S * N, D, generator=g))
    x[N:2 * N] = h
    e3 = torch.randn(S, D, N, generator=g)
    e3[1] = eps
    out = torch.full((S * N, D), 7.0, device="cuda")
    mask = torch.tensor([0, 1, 0], dtype=torch.uint8, device="cuda")
    xc, ec = x.cuda(), e3.cuda()
    _capi.check(rt.lib.dvg_gp_rsample(rt.handle, S, N, _capi.ptr(xc), D, _capi.ptr(ec), _capi.ptr(mask), _capi.ptr(out), D,
                                      _capi.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.all(out[:N] == 7.0) and torch.all(out[2 * N:] == 7.0)
    assert relerr(out[N:2 * N].transpose(0, 1), want) < 1e-4


def test_large_and_small_paths_agree_at_the_boundary():
    """M = 64 runs the shared-memory kernels, the same parameters padded to M = 65 (one extra far-away inducing
    point with zero variational weight) run the tiled path: both must give the same predictive."""
    D, M, N = 5, 64, 90
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(D, M, seed=77, trained_like=True, smooth_mean=True)
    gp, lik = make_gp(gp_sd, lik_sd)
    h = torch.tanh(torch.randn(N, D, generator=torch.Generator().manual_seed(4))).cuda()
    with torch.no_grad():
        p = lik(gp(h.transpose(0, 1).view(D, N, 1)))
        m_small, v_small = p.mean.clone(), p.variance.clone()
    ref = gp_ref.predictive(gp_sd, lik_sd, gp_ref.latent_to_gp_input(h.cpu()), torch.float64, "direct", full_cov=False)
    # extra inducing point at z = 40 (k(z, x) == 0 for |x| < 1), m_q = c there, unit L_q diagonal
    sd2 = {k: v.clone() for k, v in gp_sd.items()}
    sd2[gp_ref.K_INDUCING] = torch.cat([gp_sd[gp_ref.K_INDUCING], torch.full((D, 1, 1), 40.0)], 1)
    sd2[gp_ref.K_VMEAN] = torch.cat([gp_sd[gp_ref.K_VMEAN], gp_sd["mean_module.constant"].reshape(D, 1)], 1)
    vc = torch.zeros(D, M + 1, M + 1)
    vc[:, :M, :M] = gp_sd[gp_ref.K_VCHOL]
    vc[:, M, M] = 1.0
    sd2[gp_ref.K_VCHOL] = vc
    gp2, lik2 = make_gp(sd2, lik_sd)
    with torch.no_grad():
        p2 = lik2(gp2(h.transpose(0, 1).view(D, N, 1)))
        m_big, v_big = p2.mean, p2.variance
    assert relerr(v_small, ref["variance"]) < 1e-4 and relerr(v_big, ref["variance"]) < 1e-4
    assert relerr(m_small, ref["mean"]) < 1e-4 and relerr(m_big, ref["mean"]) < 1e-4
    assert relerr(v_big, v_small) < 2e-5


def test_trigger_large_inducing_set():
    """dvg_gp_trigger on the large-M path: values / thresholds / decisions vs oracle/trigger_ref.py."""
    from dvg_b200 import _capi
    D, M, N, S, W, T = 6, 256, 8, 37, 5, 14
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(D, M, seed=12, trained_like=True, smooth_mean=True)
    gp, lik = make_gp(gp_sd, lik_sd)
    rt = gp._runtime(lik)
    g = torch.Generator().manual_seed(6)
    lat = torch.tanh(torch.randn(T, S * N, D, generator=g) * torch.linspace(0.3, 1.5, T).reshape(T, 1, 1))
    stat_rows = (torch.arange(S) * N + 3).int().cuda()
    window = torch.zeros(S, W, device="cuda")
    count = torch.zeros(1, dtype=torch.int32, device="cuda")
    value, thr = torch.empty(S, device="cuda"), torch.empty(S, device="cuda")
    mask = torch.empty(S, dtype=torch.uint8, device="cuda")
    ctx = [[] for _ in range(S)]
    n_checked = 0
    for t in range(T):
        x = lat[t].cuda()
        warm = 1 if t < W else 0
        _capi.check(rt.lib.dvg_gp_trigger(rt.handle, S, _capi.ptr(x), D, _capi.ptr(stat_rows), _capi.ptr(window), W,
                                          _capi.ptr(count), warm, float(np.float32(trigger_ref.FACTOR)),
                                          _capi.ptr(value), _capi.ptr(thr), _capi.ptr(mask), _capi.stream_ptr()))
        v_gpu, thr_gpu, m_gpu = value.cpu().numpy(), thr.cpu().numpy(), mask.cpu().numpy()
        ref = gp_ref.predictive(gp_sd, lik_sd, gp_ref.latent_to_gp_input(lat[t, 3::N]), torch.float64, "direct",
                                full_cov=False)["variance"].float().numpy()          # [D, S]
        for s in range(S):
            v = trigger_ref.trigger_value(ref[:, s:s + 1], 0)
            assert abs(v_gpu[s] - v) <= 1e-4 * abs(v)
            if warm:
                ctx[s].append(v)
                assert m_gpu[s] == 0
            else:
                c = trigger_ref.slide(np.array(ctx[s], dtype=np.float32), v)
                ctx[s] = list(c)
                th = trigger_ref.threshold(c)
                assert abs(thr_gpu[s] - th) <= 1e-4 * abs(th)
                if abs(float(v) - float(th)) > 1e-4 * abs(float(th)):
                    assert bool(m_gpu[s]) == bool(v > th)
                    n_checked += 1
    assert n_checked > 200


@pytest.mark.parametrize("D,M", [(3, 65), (4, 200), (2, 512), (2, 1000)])
def test_native_factorisation_matches_torch_linalg(D, M):
    """dvg_gp_factorize (blocked fp64 Cholesky + triangular inverse, csrc/gp_factor.cu) against torch.linalg in fp64 on
    the same parameters: Linv [D,M,M] and beta [D,M] as the C ABI receives them (fp32), plus L Linv = I in fp64."""
    from dvg_b200 import _capi
    gp_sd, lik_sd = gp_ref.random_gp_state_dicts(D, M, seed=5 + M, trained_like=True, smooth_mean=True)
    dev = torch.device("cuda")
    z = gp_sd[gp_ref.K_INDUCING].reshape(D, M).float().to(dev).contiguous()
    m_q = gp_sd[gp_ref.K_VMEAN].float().to(dev).contiguous()
    c = gp_sd["mean_module.constant"].reshape(D).float().to(dev).contiguous()
    raw_os = gp_sd["covar_module.raw_outputscale"].reshape(D).float().to(dev).contiguous()
    raw_ls = gp_sd["covar_module.base_kernel.raw_lengthscale"].reshape(D).float().to(dev).contiguous()
    linv = torch.full((D, M, M), float("nan"), device=dev)
    beta = torch.full((D, M), float("nan"), device=dev)
    lib = _capi.load()
    dims = _capi.GpDims(D, M, 1e-3, 1e-4)
    ws = torch.empty(lib.dvg_gp_factorize_workspace(_capi.ctypes.byref(dims), 2), dtype=torch.uint8, device=dev)   # 2 dims per pass
    _capi.check(lib.dvg_gp_factorize(_capi.ctypes.byref(dims), _capi.ptr(z), _capi.ptr(m_q), _capi.ptr(c), _capi.ptr(raw_os),
                                     _capi.ptr(raw_ls), _capi.ptr(linv), _capi.ptr(beta), _capi.ptr(ws), ws.numel(),
                                     _capi.stream_ptr()), "dvg_gp_factorize")
    torch.cuda.synchronize()
    f64 = torch.float64
    sp = torch.nn.functional.softplus
    ell, sc = sp(raw_ls.to(f64)), sp(raw_os.to(f64))
    t = (z.to(f64)[:, :, None] - z.to(f64)[:, None, :]) / ell[:, None, None]
    K = sc[:, None, None] * torch.exp(-0.5 * t * t) + 1e-3 * torch.eye(M, dtype=f64, device=dev)
    L = torch.linalg.cholesky(K)
    Li = torch.linalg.solve_triangular(L, torch.eye(M, dtype=f64, device=dev).expand(D, M, M), upper=False)
    want_beta = torch.einsum("dij,dj->di", Li, m_q.to(f64) - c.to(f64)[:, None])
    assert torch.isfinite(linv).all() and torch.isfinite(beta).all()
    assert torch.all(torch.triu(linv, 1) == 0)
    scale = Li.abs().amax(dim=(1, 2), keepdim=True)
    assert ((linv.to(f64) - Li).abs() / scale).max().item() < 2e-7          # fp32 rounding of an fp64 result
    assert relerr(beta, want_beta) < 2e-6
    resid = (L @ linv.to(f64) - torch.eye(M, dtype=f64, device=dev)).abs().max().item()
    assert resid < 1e-3, resid                                               # cond(L) x fp32 rounding


def test_fp32_tiled_kernels_cross_check():
    """The FP32 FFMA tiles of gp_big.cu (the path DVG_GP_TC=0 selects; the switch is read once per process) must hold
    the same bar as the tensor-core tiles: re-run the large-M predictive and trigger tests in a child with it set."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, DVG_GP_TC="0")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", __file__, "-k",
                        "predict_large_inducing_set or trigger_large_inducing_set or agree_at_the_boundary"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
