"""Generate golden vectors from the REAL reference modules (run in the build container only).

    python tests/golden/make_golden.py

Imports ``/root/reference/models/lstm.py`` unchanged (the constructors call ``.cuda()``,
models/lstm.py:61-62, so ``torch.Tensor.cuda`` is shimmed to a no-op for this CPU run), applies the
reference's ``utils.init_weights`` semantics (utils.py:304-311: Linear ~ N(0,0.02), zero bias;
LSTMCell keeps the torch default), rolls a few steps and stores weights, inputs, noise and outputs.

The fixtures pin ``oracle/lstm_ref.py`` (tests/test_oracle_lstm.py, CPU) and the CUDA path
(tests/test_gpu_lstm.py).  /root/reference is NOT read by any test at run time.
"""
import os
import sys

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def init_weights(m):  # utils.py:304-311 semantics (utils.py itself cannot be imported: scipy.misc/skimage)
    name = m.__class__.__name__
    if name.find("Conv") != -1 or name.find("Linear") != -1:
        m.weight.data.normal_(0.0, 0.02)
        m.bias.data.fill_(0)


def roll_lstm(mod, xs):
    mod.hidden = mod.init_hidden()
    ys, hs = [], []
    with torch.no_grad():
        for x in xs:
            ys.append(mod(x).clone())
            hs.append([(h.clone(), c.clone()) for h, c in mod.hidden])
    return ys, hs


def sd_checksum(sd):
    import hashlib
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def big_cases():
    """Reference-class goldens at the REAL sizes (G90 / H256 / L2, 300 rows, 12 free-running steps) so that they reach
    the persistent tcgen05 step kernel (>= 2 row tiles).  The 4.4 MB of weights are not stored: they come from
    ``oracle.lstm_ref.random_lstm_state_dict(seed)``, are loaded into the reference class with ``load_state_dict`` and
    pinned by a SHA-256 (the test regenerates them and checks the hash).  Stored: inputs' seed, y at steps 0 / 5 / 11,
    the final hidden state (h of the top layer, c of layer 0), and for gaussian_lstm the noise and (z, mu, logvar)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import lstm_ref
    import models.lstm as ref
    G, H, L, R, T = 90, 256, 2, 300, 12
    keep = (0, 5, 11)
    sd = lstm_ref.random_lstm_state_dict(G, G, H, L, seed=21)
    sd["embed.bias"].uniform_(-0.1, 0.1, generator=torch.Generator().manual_seed(5))
    sd["output.0.bias"].uniform_(-0.1, 0.1, generator=torch.Generator().manual_seed(6))
    mod = ref.lstm(G, G, H, L, R)
    mod.load_state_dict(sd)
    mod.eval()
    gen = torch.Generator().manual_seed(77)
    xs = [torch.tanh(torch.randn(R, G, generator=gen)) for _ in range(T)]
    ys, hs = roll_lstm(mod, xs)
    torch.save({"kind": "lstm_big", "dims": (G, G, H, L, R), "steps": T, "weights_seed": 21, "x_seed": 77,
                "sha256": sd_checksum(sd), "keep": keep, "y": {t: ys[t] for t in keep},
                "h_top": hs[-1][L - 1][0], "c0": hs[-1][0][1]}, os.path.join(HERE, "big_lstm_g90_h256_r300.pt"))
    Z = 10
    gsd = lstm_ref.random_lstm_state_dict(G, Z, H, L, seed=22, gaussian=True)
    for k, sdd in (("embed.bias", 7), ("mu_net.bias", 8), ("logvar_net.bias", 9)):
        gsd[k].uniform_(-0.2, 0.2, generator=torch.Generator().manual_seed(sdd))
    gm = ref.gaussian_lstm(G, Z, H, L, R)
    gm.load_state_dict(gsd)
    gm.eval()
    gm.hidden = gm.init_hidden()
    gen = torch.Generator().manual_seed(78)
    xs = [torch.tanh(torch.randn(R, G, generator=gen)) for _ in range(T)]
    outs, epss = [], []
    with torch.no_grad():
        for t, x in enumerate(xs):
            torch.manual_seed(300 + t)
            z, mu, logvar = gm(x)
            torch.manual_seed(300 + t)
            eps = torch.empty(R, Z).normal_()          # models/lstm.py:163 draws eps with .normal_() first
            assert torch.equal(z, eps.mul(logvar.mul(0.5).exp()).add(mu))
            outs.append((z.clone(), mu.clone(), logvar.clone()))
            epss.append(eps)
    torch.save({"kind": "gauss_big", "dims": (G, Z, H, L, R), "steps": T, "weights_seed": 22, "x_seed": 78,
                "sha256": sd_checksum(gsd), "eps": epss, "out": outs,
                "h_top": gm.hidden[L - 1][0].clone(), "c0": gm.hidden[0][1].clone()},
               os.path.join(HERE, "big_gauss_g90_z10_h256_r300.pt"))


def main():
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self
    import models.lstm as ref

    cases = {
        # name: (G_in, G_out, H, L, B, steps)
        "tiny": (12, 12, 32, 2, 5, 4),
        "g90_h64": (90, 90, 64, 2, 7, 3),
        "g90_h64_l3": (90, 90, 64, 3, 3, 2),
    }
    for name, (gi, go, H, L, B, T) in cases.items():
        torch.manual_seed(1)
        mod = ref.lstm(gi, go, H, L, B)
        mod.apply(init_weights)
        # non-zero biases so the bias path is exercised
        for p in (mod.embed.bias, mod.output[0].bias):
            p.data.uniform_(-0.1, 0.1)
        mod.eval()
        xs = [torch.tanh(torch.randn(B, gi)) for _ in range(T)]
        ys, hs = roll_lstm(mod, xs)
        torch.save({"kind": "lstm", "dims": (gi, go, H, L, B), "state_dict": mod.state_dict(),
                    "x": xs, "y": ys, "hidden": hs}, os.path.join(HERE, f"lstm_{name}.pt"))

    gcases = {"tiny": (12, 6, 32, 1, 5, 3), "g90_z10_h64": (90, 10, 64, 1, 6, 3), "z10_l2": (20, 10, 32, 2, 4, 2)}
    for name, (gi, Z, H, L, B, T) in gcases.items():
        torch.manual_seed(2)
        mod = ref.gaussian_lstm(gi, Z, H, L, B)
        mod.apply(init_weights)
        for p in (mod.embed.bias, mod.mu_net.bias, mod.logvar_net.bias):
            p.data.uniform_(-0.2, 0.2)
        mod.eval()
        mod.hidden = mod.init_hidden()
        xs = [torch.tanh(torch.randn(B, gi)) for _ in range(T)]
        out, eps_list, hs = [], [], []
        with torch.no_grad():
            for t, x in enumerate(xs):
                torch.manual_seed(100 + t)
                z, mu, logvar = mod(x)
                torch.manual_seed(100 + t)          # models/lstm.py:163 draws eps with .normal_() first
                eps = torch.empty(B, Z).normal_()
                assert torch.equal(z, eps.mul(logvar.mul(0.5).exp()).add(mu))
                out.append((z.clone(), mu.clone(), logvar.clone()))
                eps_list.append(eps)
                hs.append([(h.clone(), c.clone()) for h, c in mod.hidden])
        torch.save({"kind": "gaussian_lstm", "dims": (gi, Z, H, L, B), "state_dict": mod.state_dict(),
                    "x": xs, "eps": eps_list, "out": out, "hidden": hs},
                   os.path.join(HERE, f"gauss_{name}.pt"))
    big_cases()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
