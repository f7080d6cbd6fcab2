"""Timeline of lstm_small_kernel (DVG_SMALL_TRACE=1): eager steps at 16 rows, the 7th launch is dumped to stderr."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, build_models
w = WORKLOADS["smmnist_b16"]
dev = torch.device("cuda", 0)
fp, gp, lik = build_models(w, dev, "bf16x3")
x = torch.tanh(torch.randn(w["B"], w["G"], device=dev))
with torch.no_grad():
    fp.hidden = fp.init_hidden()
    for _ in range(10):
        y = fp(x)
torch.cuda.synchronize()
print("ok")
