"""Oracle: plain-torch CPU restatement of the reference LSTM cells.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Pinned against the real
reference classes through ``tests/golden/lstm_*.pt``.

Restates (reference paths relative to /root/reference):

* ``models/lstm.py:65-72``   ``lstm.forward``: embed -> L x LSTMCell -> Linear+Tanh
* ``models/lstm.py:58-63``   ``lstm.init_hidden``: L tuples of zeros [B,H]
* ``models/lstm.py:161-164`` ``gaussian_lstm.reparameterize``
* ``models/lstm.py:166-175`` ``gaussian_lstm.forward``

The LSTMCell arithmetic (torch.nn.LSTMCell, gate chunk order i,f,g,o) is
written out explicitly so the CUDA kernels can be compared op by op.
"""
from __future__ import annotations

import torch


def init_hidden(n_layers: int, rows: int, hidden_size: int, dtype=torch.float32):
    """models/lstm.py:58-63 (without the .cuda())."""
    return [(torch.zeros(rows, hidden_size, dtype=dtype),
             torch.zeros(rows, hidden_size, dtype=dtype)) for _ in range(n_layers)]


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.LSTMCell as called from models/lstm.py:69.

    gates = x W_ih^T + b_ih + h W_hh^T + b_hh ; chunks i,f,g,o ;
    c' = sigmoid(f) c + sigmoid(i) tanh(g) ; h' = sigmoid(o) tanh(c').
    """
    gates = torch.addmm(b_ih, x, w_ih.t()) + torch.addmm(b_hh, h, w_hh.t())
    i, f, g, o = gates.chunk(4, dim=1)
    c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    h_new = torch.sigmoid(o) * torch.tanh(c_new)
    return h_new, c_new


def _trunk(sd, x, hidden, n_layers):
    in_size = sd["embed.weight"].shape[1]
    h_in = torch.addmm(sd["embed.bias"], x.reshape(-1, in_size), sd["embed.weight"].t())
    new_hidden = []
    for l in range(n_layers):
        h, c = hidden[l]
        h2, c2 = lstm_cell(h_in, h, c,
                           sd[f"lstm.{l}.weight_ih"], sd[f"lstm.{l}.weight_hh"],
                           sd[f"lstm.{l}.bias_ih"], sd[f"lstm.{l}.bias_hh"])
        new_hidden.append((h2, c2))
        h_in = h2
    return h_in, new_hidden


def n_layers_of(sd) -> int:
    return len([k for k in sd if k.startswith("lstm.") and k.endswith(".weight_ih")])


def lstm_forward(sd, x, hidden):
    """models/lstm.py:65-72.  ``sd`` = reference state_dict (fp32 or fp64).

    Returns (y [R,G] in (-1,1), new_hidden)."""
    h_top, new_hidden = _trunk(sd, x, hidden, n_layers_of(sd))
    y = torch.tanh(torch.addmm(sd["output.0.bias"], h_top, sd["output.0.weight"].t()))
    return y, new_hidden


def gaussian_lstm_forward(sd, x, hidden, eps):
    """models/lstm.py:166-175 with the noise of :163 injected as ``eps`` [R,Z].

    Returns (z, mu, logvar, new_hidden); z = eps * exp(0.5 logvar) + mu."""
    h_top, new_hidden = _trunk(sd, x, hidden, n_layers_of(sd))
    mu = torch.addmm(sd["mu_net.bias"], h_top, sd["mu_net.weight"].t())
    logvar = torch.addmm(sd["logvar_net.bias"], h_top, sd["logvar_net.weight"].t())
    sigma = logvar.mul(0.5).exp()
    z = eps.mul(sigma).add(mu)
    return z, mu, logvar, new_hidden


def to_dtype(sd, dtype):
    return {k: v.detach().to(dtype) for k, v in sd.items()}


def random_lstm_state_dict(g_in, g_out, hidden, n_layers, seed=1, gaussian=False, dtype=torch.float32):
    """Random-init weights with the reference's semantics: Linear layers per
    utils.py:304-311 (N(0,0.02), zero bias), LSTMCell torch default U(+-1/sqrt(H))."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    sd["embed.weight"] = torch.randn(hidden, g_in, generator=gen) * 0.02
    sd["embed.bias"] = torch.zeros(hidden)
    k = 1.0 / hidden ** 0.5
    for l in range(n_layers):
        for name, shape in (("weight_ih", (4 * hidden, hidden)), ("weight_hh", (4 * hidden, hidden)),
                            ("bias_ih", (4 * hidden,)), ("bias_hh", (4 * hidden,))):
            sd[f"lstm.{l}.{name}"] = (torch.rand(*shape, generator=gen) * 2 - 1) * k
    if gaussian:
        for head in ("mu_net", "logvar_net"):
            sd[f"{head}.weight"] = torch.randn(g_out, hidden, generator=gen) * 0.02
            sd[f"{head}.bias"] = torch.zeros(g_out)
    else:
        sd["output.0.weight"] = torch.randn(g_out, hidden, generator=gen) * 0.02
        sd["output.0.bias"] = torch.zeros(g_out)
    return {k_: v.to(dtype) for k_, v in sd.items()}
