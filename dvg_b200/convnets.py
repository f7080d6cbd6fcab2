"""Encoder / decoder conv stacks of the reference on the STOCK PyTorch path (north_star: "the DCGAN/VGG-64/128
encoder and decoder convolutions stay on the reference PyTorch path").  Harness only -- they are not part of the
hot path and carry no kernels of ours.  Table-driven re-statement with the reference's sub-module names so that
``state_dict``s interchange with models/{dcgan_64,dcgan_128,vgg_64,vgg_128}.py:

    encoder(x [N,C,W,W]) -> (latent [N, dim], [skip_1 .. skip_k])      decoder([latent, skips]) -> frame

Architectures (reference file: lines): dcgan_64 (models/dcgan_64.py:28-88), dcgan_128 (models/dcgan_128.py:28-94),
vgg_64 (models/vgg_64.py:17-106), vgg_128 (models/vgg_128.py:16-120).
"""
from __future__ import annotations

import torch
import torch.nn as nn

NF = 64


class _Block(nn.Module):
    """conv/upconv + BatchNorm + LeakyReLU(0.2) held in ``.main`` (dcgan_conv / dcgan_upconv / vgg_layer)."""

    def __init__(self, conv):
        super().__init__()
        self.main = nn.Sequential(conv, nn.BatchNorm2d(conv.out_channels), nn.LeakyReLU(0.2, inplace=True))

    def forward(self, x):
        return self.main(x)


def _down(i, o):
    return _Block(nn.Conv2d(i, o, 4, 2, 1))


def _upconv(i, o):
    return _Block(nn.ConvTranspose2d(i, o, 4, 2, 1))


def _vgg(*chans):
    return nn.Sequential(*[_Block(nn.Conv2d(i, o, 3, 1, 1)) for i, o in zip(chans[:-1], chans[1:])])


_DCGAN = {64: [NF, 2 * NF, 4 * NF, 8 * NF], 128: [NF, 2 * NF, 4 * NF, 8 * NF, 8 * NF]}
_VGG_ENC = {64: [(64, 64), (128, 128), (256, 256, 256), (512, 512, 512)],
            128: [(64, 64), (128, 128), (256, 256, 256), (512, 512, 512), (512, 512, 512)]}
# decoder stages after the 1x1 -> 4x4 stem: channel chains (input is 2x the first number: upsampled + skip)
_VGG_DEC = {64: [(512, 512, 512, 256), (256, 256, 256, 128), (128, 128, 64)],
            128: [(512, 512, 512, 512), (512, 512, 512, 256), (256, 256, 256, 128), (128, 128, 64)]}


class Encoder(nn.Module):
    def __init__(self, arch: str, width: int, dim: int, nc: int = 1):
        super().__init__()
        self.arch, self.dim = arch, dim
        if arch == "dcgan":
            chans = [nc] + _DCGAN[width]
            self.stages = len(chans) - 1
            for k in range(self.stages):
                setattr(self, f"c{k + 1}", _down(chans[k], chans[k + 1]))
            last = chans[-1]
        else:
            spec = _VGG_ENC[width]
            self.stages = len(spec)
            prev = nc
            for k, st in enumerate(spec):
                setattr(self, f"c{k + 1}", _vgg(prev, *st))
                prev = st[-1]
            last = prev
            self.mp = nn.MaxPool2d(kernel_size=2, stride=2, padding=0)
        setattr(self, f"c{self.stages + 1}", nn.Sequential(nn.Conv2d(last, dim, 4, 1, 0), nn.BatchNorm2d(dim), nn.Tanh()))

    def forward(self, x):
        skips = []
        h = x
        for k in range(self.stages):
            if self.arch == "vgg" and k > 0:
                h = self.mp(h)
            h = getattr(self, f"c{k + 1}")(h)
            skips.append(h)
        if self.arch == "vgg":
            h = self.mp(h)
        h = getattr(self, f"c{self.stages + 1}")(h)
        return h.view(-1, self.dim), skips


class Decoder(nn.Module):
    def __init__(self, arch: str, width: int, dim: int, nc: int = 1):
        super().__init__()
        self.arch, self.dim = arch, dim
        top = 8 * NF if arch == "dcgan" else 512
        self.upc1 = nn.Sequential(nn.ConvTranspose2d(dim, top, 4, 1, 0), nn.BatchNorm2d(top), nn.LeakyReLU(0.2, inplace=True))
        if arch == "dcgan":
            chans = list(reversed(_DCGAN[width]))            # e.g. [512, 256, 128, 64]
            self.n_up = len(chans)
            for k in range(1, self.n_up):
                setattr(self, f"upc{k + 1}", _upconv(chans[k - 1] * 2, chans[k]))
            final_act = nn.Tanh() if width == 64 else nn.Sigmoid()   # dcgan_64.py:75-79 / dcgan_128.py:80-84
            setattr(self, f"upc{self.n_up + 1}", nn.Sequential(nn.ConvTranspose2d(chans[-1] * 2, nc, 4, 2, 1), final_act))
        else:
            spec = _VGG_DEC[width]
            self.n_up = len(spec) + 1
            for k, st in enumerate(spec):
                setattr(self, f"upc{k + 2}", _vgg(st[0] * 2, *st[1:]))
            setattr(self, f"upc{self.n_up + 1}",
                    nn.Sequential(_Block(nn.Conv2d(64 * 2, 64, 3, 1, 1)), nn.ConvTranspose2d(64, nc, 3, 1, 1), nn.Sigmoid()))
            self.up = nn.UpsamplingNearest2d(scale_factor=2)

    def forward(self, inp):
        vec, skip = inp
        d = self.upc1(vec.view(-1, self.dim, 1, 1))
        n = len(skip)
        for k in range(n):
            if self.arch == "vgg":
                d = self.up(d)
            d = getattr(self, f"upc{k + 2}")(torch.cat([d, skip[n - 1 - k]], 1))
        return d


def make_codec(model: str, dim: int, nc: int = 1):
    """model in {dcgan_64, dcgan_128, vgg_64, vgg_128} -> (encoder, decoder)."""
    arch, width = model.split("_")
    return Encoder(arch, int(width), dim, nc), Decoder(arch, int(width), dim, nc)
