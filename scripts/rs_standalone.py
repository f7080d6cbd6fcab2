import sys, os, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from oracle import gp_ref
from util import make_gp
D, M, N = 90, 40, 50
gp_sd, lik_sd = gp_ref.random_gp_state_dicts(D, M, seed=21, trained_like=True, smooth_mean=True)
gp, lik = make_gp(gp_sd, lik_sd)
h = torch.tanh(torch.randn(N, D)).cuda()
eps = torch.randn(D, N).cuda()
with torch.no_grad():
    for _ in range(3):
        got = lik(gp(h.transpose(0, 1).view(D, N, 1))).rsample(eps=eps)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
with torch.no_grad():
    for _ in range(20):
        got = lik(gp(h.transpose(0, 1).view(D, N, 1))).rsample(eps=eps)
e1.record(); torch.cuda.synchronize()
print("standalone rsample us per call (incl. python):", e0.elapsed_time(e1) / 20 * 1e3)
