"""Sample-batched execution of the encoder / decoder conv stacks.

The convolutions stay on the stock PyTorch / cuDNN path (north_star); what changes is how they are driven when the
S diverse futures of generate_frames.py:138-178 are batched into S*B rows:

* eval-mode BatchNorm is folded into the preceding conv (``fold_batchnorm``) -- one launch less per block, same
  arithmetic up to fp32 rounding;
* weights and activations are channels-last (cuDNN's native tensor-core layout), optionally bf16;
* rows are processed in chunks that are multiples of B so activations stay bounded
  (vgg_64 at 5000 rows: 5.2 GB per 64-channel full-resolution activation in fp32);
* the skip connections come from the last CONTEXT frame (generate_frames.py:154-157 with ``last_frame_skip`` off), so
  they are identical for the S samples of a sequence: the first conv of every decoder stage is linear in its input
  ``cat([d, skip])`` and is split into ``conv_d(d) + conv_s(skip)``; the skip half is computed ONCE for the B context
  rows and broadcast-added over S instead of replicating the skips S times and convolving them S times
  (``SharedSkipDecoder``).  That removes half the MACs of those convs and the S-fold skip copies.

Works on the harness nets of ``dvg_b200.convnets`` and on the reference's own model classes (same sub-module names:
``c1..``, ``upc1..``, ``up``, ``mp``; models/{dcgan,vgg}_{64,128}.py).
"""
from __future__ import annotations

import copy
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
from torch.nn.utils.fusion import fuse_conv_bn_eval

_CONVS = (nn.Conv2d, nn.ConvTranspose2d)


def _fold(m: nn.Module) -> nn.Module:
    if isinstance(m, nn.Sequential):
        mods = list(m.children())
        out = []
        i = 0
        while i < len(mods):
            a = mods[i]
            if isinstance(a, _CONVS) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm2d):
                out.append(fuse_conv_bn_eval(a, mods[i + 1], transpose=isinstance(a, nn.ConvTranspose2d)))
                out.append(nn.Identity())                      # keeps the positions of the following modules
                i += 2
            else:
                out.append(_fold(a))
                i += 1
        return nn.Sequential(*out)
    for name, child in list(m.named_children()):
        setattr(m, name, _fold(child))
    return m


def fold_batchnorm(net: nn.Module) -> nn.Module:
    """Deep copy of ``net`` in eval mode with every (conv, BatchNorm2d) pair of a Sequential fused into one conv."""
    net = copy.deepcopy(net).eval()
    for p in net.parameters():
        p.requires_grad_(False)
    return _fold(net)


class _SplitConv(nn.Module):
    """First conv of a decoder stage, split over its input channels: y = conv_d(d) + partial, with
    ``partial = conv_s(skip) + bias`` computed once per sequence and broadcast over the samples."""

    def __init__(self, conv: nn.Module, c_d: int):
        super().__init__()
        tr = isinstance(conv, nn.ConvTranspose2d)
        w = conv.weight.detach()
        wd, ws = (w[:c_d], w[c_d:]) if tr else (w[:, :c_d], w[:, c_d:])
        kw = dict(kernel_size=conv.kernel_size, stride=conv.stride, padding=conv.padding)
        cls = type(conv)
        o = conv.out_channels
        self.conv_d = cls(c_d, o, bias=False, **kw)
        self.conv_s = cls(w.shape[0 if tr else 1] - c_d, o, bias=conv.bias is not None, **kw)
        self.conv_d.weight = nn.Parameter(wd.clone().contiguous(), requires_grad=False)
        self.conv_s.weight = nn.Parameter(ws.clone().contiguous(), requires_grad=False)
        if conv.bias is not None:
            self.conv_s.bias = nn.Parameter(conv.bias.detach().clone(), requires_grad=False)
        self.partial: Optional[torch.Tensor] = None

    def forward(self, d):
        y = self.conv_d(d)
        p = self.partial                                       # [B, o, h, w]
        y5 = y.view(-1, p.shape[0], *y.shape[1:])              # [s, B, o, h, w] (a view: only dim 0 is split)
        y5 += p
        return y


def _first_conv_path(stage: nn.Module):
    """(parent, attribute name) of the first conv executed by ``stage`` (Sequentials run in registration order)."""
    for parent in stage.modules():
        for name, child in parent.named_children():
            if isinstance(child, _CONVS):
                return parent, name
            break                                              # only the FIRST child of each container is on the path
    raise ValueError("decoder stage without a leading convolution")


class SharedSkipDecoder(nn.Module):
    """Decoder whose skip inputs are shared by the S samples of each sequence (see module docstring)."""

    def __init__(self, decoder: nn.Module):
        super().__init__()
        dec = fold_batchnorm(decoder)
        self.dim = dec.dim
        self.upc1 = dec.upc1
        self.up = getattr(dec, "up", None)
        stages = []
        k = 2
        while hasattr(dec, f"upc{k}"):
            stages.append(getattr(dec, f"upc{k}"))
            k += 1
        self.splits: List[_SplitConv] = []
        for st in stages:
            parent, name = _first_conv_path(st)
            conv = getattr(parent, name)
            cin = conv.in_channels
            sp = _SplitConv(conv, cin // 2)
            setattr(parent, name, sp)
            self.splits.append(sp)
        self.stages = nn.ModuleList(stages)

    @torch.no_grad()
    def set_skips(self, skips: Sequence[torch.Tensor]):
        """``skips``: the encoder's skip list for the B context rows (finest first, as the encoders return it)."""
        n = len(self.stages)
        assert len(skips) == n, (len(skips), n)
        for k, sp in enumerate(self.splits):
            sp.partial = sp.conv_s(skips[n - 1 - k])

    def forward(self, vec):
        d = self.upc1(vec.view(-1, self.dim, 1, 1))
        for st in self.stages:
            if self.up is not None:
                d = self.up(d)
            d = st(d)
        return d


class BatchedCodec:
    """Inference-time driver of one (encoder, decoder) pair for S*B-row rollouts.

    ``encode(x)`` -> (latent [R, G] fp32 contiguous, skips | None); ``set_shared_skips(skips_B)`` then
    ``decode_shared(vec [S*B, G])`` -> frames [S*B, C, W, W] fp32; ``decode(vec, skips)`` is the general per-row-skip
    form (``last_frame_skip``)."""

    def __init__(self, encoder: nn.Module, decoder: nn.Module, n_points: int, dtype=torch.float32,
                 channels_last: bool = True, chunk_rows: Optional[int] = None):
        self.B = n_points
        self.dtype = dtype
        self.mf = torch.channels_last if channels_last else torch.contiguous_format
        self.enc = fold_batchnorm(encoder).to(dtype=dtype, memory_format=self.mf)
        self.dec = fold_batchnorm(decoder).to(dtype=dtype, memory_format=self.mf)
        self.sdec = SharedSkipDecoder(decoder).to(dtype=dtype, memory_format=self.mf)
        self.chunk_rows = chunk_rows

    def _chunk(self, rows: int, width: int) -> int:
        if self.chunk_rows is not None:
            c = self.chunk_rows
        else:                                                  # largest activation: 64..128 channels at full resolution
            c = (1 << 30) // (128 * width * width * torch.empty((), dtype=self.dtype).element_size())
        c = max(self.B, c // self.B * self.B)
        return min(rows, c)

    def _in(self, x):
        return x.to(dtype=self.dtype).contiguous(memory_format=self.mf)

    @torch.no_grad()
    def encode(self, x: torch.Tensor, want_skips: bool = True):
        R = x.shape[0]
        c = self._chunk(R, x.shape[-1])
        hs, sks = [], []
        for r0 in range(0, R, c):
            h, sk = self.enc(self._in(x[r0:r0 + c]))
            hs.append(h.float())
            if want_skips:
                sks.append(sk)
        h = hs[0] if len(hs) == 1 else torch.cat(hs)
        if not want_skips:
            return h.contiguous(), None
        skips = list(sks[0]) if len(sks) == 1 else [torch.cat([s[k] for s in sks]) for k in range(len(sks[0]))]
        return h.contiguous(), skips

    @torch.no_grad()
    def set_shared_skips(self, skips: Sequence[torch.Tensor]):
        assert skips[0].shape[0] == self.B, "shared skips are the B context rows"
        self.sdec.set_skips([self._in(s) for s in skips])

    @torch.no_grad()
    def decode_shared(self, vec: torch.Tensor, out: Optional[torch.Tensor] = None):
        R = vec.shape[0]
        assert R % self.B == 0
        width = self.sdec.splits[-1].partial.shape[-1]
        c = self._chunk(R, width)
        outs = []
        for r0 in range(0, R, c):
            y = self.sdec(vec[r0:r0 + c].to(self.dtype))
            if out is not None:
                out[r0:r0 + c].copy_(y)
            else:
                outs.append(y.float())
        if out is not None:
            return out
        y = outs[0] if len(outs) == 1 else torch.cat(outs)
        return y.contiguous()

    @torch.no_grad()
    def decode(self, vec: torch.Tensor, skips: Sequence[torch.Tensor]):
        R = vec.shape[0]
        c = self._chunk(R, skips[0].shape[-1])
        outs = []
        for r0 in range(0, R, c):
            y = self.dec([vec[r0:r0 + c].to(self.dtype), [self._in(s[r0:r0 + c]) for s in skips]])
            outs.append(y.float())
        y = outs[0] if len(outs) == 1 else torch.cat(outs)
        return y.contiguous()
