"""Oracle: sequential CPU restatement of the reference rollout loops.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

Restates, with injectable noise and config-driven sizes instead of the
hard-coded ``view(90,50,1)`` / 12 / 105 / 15:

* ``generate_frames.py:111-134``  make_gifs pass A ("approx. posterior": GP mean on the LSTM output)
* ``generate_frames.py:138-178``  make_gifs pass B (S diverse futures, sequential python loop over s,
                                  resample when ``i % resample_every == 0``; LSTM *is* advanced on a
                                  resample step, :166 precedes the test)
* ``train.py:262-289``            plot(): same as pass B with S=5 and resample only at ``i == 10``
* ``generate_frames.py:249-300``  GPtrigger_gen for ONE ``index`` (variance trigger; LSTM is *not*
                                  advanced on a triggered step, :289-295)

``encoder(x) -> (h [N,G], skip)`` and ``decoder([vec [N,G], skip]) -> x`` are any callables (the
reference conv nets, or latent-space stand-ins for hot-path-only tests).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import gp_ref, lstm_ref, trigger_ref


@dataclass
class OracleModels:
    lstm_sd: Dict[str, torch.Tensor]
    gp_sd: Dict[str, torch.Tensor]
    lik_sd: Dict[str, torch.Tensor]
    encoder: Callable
    decoder: Callable
    dtype: torch.dtype = torch.float32
    gp_mode: str = "gpytorch"
    hidden: Optional[list] = None

    def init_hidden(self, rows):
        H = self.lstm_sd["embed.weight"].shape[0]
        self.hidden = lstm_ref.init_hidden(lstm_ref.n_layers_of(self.lstm_sd), rows, H, self.dtype)

    def frame_predictor(self, h):
        y, self.hidden = lstm_ref.lstm_forward(self.lstm_sd, h.to(self.dtype), self.hidden)
        return y

    def gp(self, h, full_cov=True):
        """likelihood(gp_layer(h.transpose(0,1).view(D,N,1)))."""
        return gp_ref.predictive(self.gp_sd, self.lik_sd, gp_ref.latent_to_gp_input(h), self.dtype,
                                 self.gp_mode, full_cov=full_cov)


def posterior_rollout(m: OracleModels, x: Sequence[torch.Tensor], n_past: int, n_eval: int,
                      last_frame_skip: bool = False):
    """generate_frames.py:111-134.  Returns posterior_gen (list of n_eval frames)."""
    m.init_hidden(x[0].shape[0])
    gen = [x[0]]
    x_in = x[0]
    skip = None
    for i in range(1, n_eval):
        h, sk = m.encoder(x_in)
        if last_frame_skip or i < n_past:
            skip = sk
        if i < n_past:
            m.frame_predictor(h)
            x_in = x[i]
        else:
            h_pred = m.frame_predictor(h)
            pred = m.gp(h_pred, full_cov=False)
            x_in = m.decoder([pred["mean"].transpose(0, 1), skip])
        gen.append(x_in)
    return gen


def diverse_rollout(m: OracleModels, x: Sequence[torch.Tensor], n_past: int, n_eval: int,
                    nsample: int, eps: Dict, resample_every: Optional[int] = 15,
                    resample_at: Optional[Sequence[int]] = None, last_frame_skip: bool = False,
                    record_latents: bool = False):
    """generate_frames.py:138-178 (resample_every=15) / train.py:262-289 (resample_at=[10]).

    ``eps[(s, i)]`` is the [D,N] standard-normal draw used by ``rsample`` of sample s at step i.
    Returns all_gen[s][t] (frames), and if record_latents the decoder inputs lat[s][t]."""
    all_gen, all_lat = [], []
    for s in range(nsample):
        m.init_hidden(x[0].shape[0])
        x_in = x[0]
        gen, lat = [x_in], [None]
        skip = None
        for i in range(1, n_eval):
            h, sk = m.encoder(x_in)
            if last_frame_skip or i < n_past:
                skip = sk
            if i < n_past:
                m.frame_predictor(h)
                x_in = x[i]
                lat.append(None)
            else:
                h_pred = m.frame_predictor(h)
                hit = (resample_every is not None and i % resample_every == 0) or \
                      (resample_at is not None and i in resample_at)
                if hit:
                    pred = m.gp(h, full_cov=True)
                    vec = gp_ref.rsample(pred["mean"], pred["covar"], eps[(s, i)]).transpose(0, 1)
                else:
                    vec = h_pred
                lat.append(vec)
                x_in = m.decoder([vec, skip])
            gen.append(x_in)
        all_gen.append(gen)
        all_lat.append(lat)
    return (all_gen, all_lat) if record_latents else all_gen


def trigger_rollout(m: OracleModels, x0: torch.Tensor, eps: Dict, warmup: int = 12, n_steps: int = 105,
                    stat_col_warmup: int = 0, stat_col: int = 3, skip_until: int = 5):
    """generate_frames.py:252-298 for one ``index`` (= stat_col_warmup).

    ``eps[i]`` is the [D,N] draw for a triggered step i.  Returns dict(gen_seq, values, triggers,
    thresholds, latents)."""
    m.init_hidden(x0.shape[0])
    x_in = x0
    ctx, values, triggers, thresholds, gen_seq, latents = [], [], [], [], [], []
    skip = None
    for i in range(warmup):
        h, sk = m.encoder(x_in)
        if i < skip_until:
            skip = sk
        pred = m.gp(h, full_cov=False)
        value = trigger_ref.trigger_value(pred["variance"].detach().to(torch.float32).numpy(), stat_col_warmup)
        ctx.append(value)
        vec = m.frame_predictor(m.encoder(x_in)[0])          # generation(), :220-224
        x_out = m.decoder([vec, skip])
        values.append(value); gen_seq.append(x_out); latents.append(vec); triggers.append(False)
        thresholds.append(np.float32("nan"))
        x_in = x_out
    ctx = np.array(ctx, dtype=np.float32)
    for i in range(warmup, n_steps):
        h = m.encoder(x_in)[0]
        pred = m.gp(h, full_cov=False)
        value = trigger_ref.trigger_value(pred["variance"].detach().to(torch.float32).numpy(), stat_col)
        ctx = trigger_ref.slide(ctx, value)
        thr = trigger_ref.threshold(ctx)
        fired = bool(value > thr)
        if fired:
            predc = m.gp(h, full_cov=True)
            vec = gp_ref.rsample(predc["mean"], predc["covar"], eps[i]).transpose(0, 1)
        else:
            vec = m.frame_predictor(h)
        x_out = m.decoder([vec, skip])
        values.append(value); triggers.append(fired); thresholds.append(thr)
        gen_seq.append(x_out); latents.append(vec)
        x_in = x_out
    return {"gen_seq": gen_seq, "values": np.array(values, dtype=np.float32), "triggers": triggers,
            "thresholds": np.array(thresholds, dtype=np.float32), "latents": latents}
