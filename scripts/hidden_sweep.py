"""Hidden-size sweep of the LSTM step (BASELINE configs[4]: hidden 256-1024, samples 16-4096): CUDA-graph replay of 12
plain steps per (H, rows), ours (bf16x3) next to the stock torch modules (fp32 / TF32 matmuls) on the same GPU.
    python scripts/hidden_sweep.py > gpurun_out/r02_hidden_sweep.jsonl"""
import json
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import graph_time_us  # noqa: E402
from dvg_b200.init import init_lstm_state_dict  # noqa: E402
from dvg_b200.models.lstm import lstm  # noqa: E402

dev = torch.device("cuda", 0)
G, L, T = 90, 2, 12
for H in (256, 512, 1024):
    sd = init_lstm_state_dict(G, G, H, L, 1)
    for R in (16, 256, 4096):
        m = lstm(G, G, H, L, R)
        m.load_state_dict(sd)
        m = m.to(dev).eval()
        x = torch.tanh(torch.randn(T, R, G, device=dev))

        def ours():
            for t in range(T):
                m(x[t])
        with torch.no_grad():
            m.hidden = m.init_hidden()
            ours()
        us_ours = graph_time_us(ours, T)[0]

        class Stock(nn.Module):                       # the reference module's ops (models/lstm.py:42-72)
            def __init__(s_):
                super().__init__()
                s_.embed = nn.Linear(G, H)
                s_.lstm = nn.ModuleList([nn.LSTMCell(H, H) for _ in range(L)])
                s_.output = nn.Sequential(nn.Linear(H, G), nn.Tanh())

            def forward(s_, xx):
                h_in = s_.embed(xx)
                for i in range(L):
                    s_.hidden[i] = s_.lstm[i](h_in, s_.hidden[i])
                    h_in = s_.hidden[i][0]
                return s_.output(h_in)
        st = Stock()
        st.load_state_dict(sd)
        st = st.to(dev).eval()
        h0 = [(torch.zeros(R, H, device=dev), torch.zeros(R, H, device=dev)) for _ in range(L)]

        def stock():
            st.hidden = list(h0)
            for t in range(T):
                st(x[t])
        res = {"H": H, "rows": R, "ours_us_per_step": round(us_ours, 2)}
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            res["stock_%s_us_per_step" % name] = round(graph_time_us(stock, T)[0], 2)
        torch.backends.cuda.matmul.allow_tf32 = False
        flops = 2 * (G * H + L * 2 * H * 4 * H + H * G) * R
        res["ours_tflops_algorithmic"] = round(flops / us_ours / 1e6, 1)
        print(json.dumps(res), flush=True)
        del m, st
