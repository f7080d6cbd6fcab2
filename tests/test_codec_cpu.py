"""Sample-batched conv execution (dvg_b200/codec.py) computes the same function as the plain
eval-mode nets: BatchNorm folding, the shared-skip split of the decoder convs, row chunking.  CPU, fp32."""
import importlib
import os
import sys

import pytest
import torch

from dvg_b200.codec import BatchedCodec, SharedSkipDecoder, fold_batchnorm
from dvg_b200.convnets import make_codec

REF = "/root/reference"
CASES = [("dcgan_64", 1, 64), ("vgg_64", 1, 64), ("dcgan_128", 3, 128), ("vgg_128", 3, 128)]


def _randomise_bn(net, seed):
    g = torch.Generator().manual_seed(seed)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    return net.eval()


def _close(a, b, tol=2e-4):
    err = (a - b).abs().max().item()
    ref = b.abs().max().item()
    assert err <= tol * max(ref, 1e-3), (err, ref)


def _nets(model, nc, dim=90):
    torch.manual_seed(3)
    enc, dec = make_codec(model, dim, nc)
    return _randomise_bn(enc, 1), _randomise_bn(dec, 2)


@pytest.mark.parametrize("model,nc,width", CASES)
def test_fold_batchnorm_matches_eval_nets(model, nc, width):
    enc, dec = _nets(model, nc)
    fe, fd = fold_batchnorm(enc), fold_batchnorm(dec)
    assert not any(isinstance(m, torch.nn.BatchNorm2d) for m in list(fe.modules()) + list(fd.modules()))
    x = torch.rand(2, nc, width, width)
    with torch.no_grad():
        h, sk = enc(x)
        fh, fsk = fe(x)
        _close(fh, h)
        for a, b in zip(fsk, sk):
            _close(a, b)
        _close(fd([h, sk]), dec([h, sk]))


@pytest.mark.parametrize("model,nc,width", CASES)
def test_shared_skip_decoder_matches_replicated_skips(model, nc, width):
    enc, dec = _nets(model, nc)
    S, B = 3, 2
    x = torch.rand(B, nc, width, width)
    vec = torch.tanh(torch.randn(S * B, 90))
    with torch.no_grad():
        _, sk = enc(x)
        want = dec([vec, [s.repeat(S, 1, 1, 1) for s in sk]])
        sd = SharedSkipDecoder(dec)
        sd.set_skips(sk)
        _close(sd(vec), want)
        sd = sd.to(memory_format=torch.channels_last)
        sd.set_skips([s.contiguous(memory_format=torch.channels_last) for s in sk])
        _close(sd(vec), want)


@pytest.mark.parametrize("chunk", [None, 2, 4])
def test_batched_codec_chunks(chunk):
    enc, dec = _nets("dcgan_64", 1)
    S, B = 3, 2
    codec = BatchedCodec(enc, dec, n_points=B, chunk_rows=chunk)
    x = torch.rand(S * B, 1, 64, 64)
    vec = torch.tanh(torch.randn(S * B, 90))
    with torch.no_grad():
        h, sk = enc(x)
        ch, csk = codec.encode(x)
        assert ch.is_contiguous() and ch.dtype == torch.float32
        _close(ch, h)
        for a, b in zip(csk, sk):
            _close(a, b)
        ch2, none = codec.encode(x, want_skips=False)
        assert none is None and torch.equal(ch2, ch)
        _close(codec.decode(vec, sk), dec([vec, sk]))
        codec.set_shared_skips([s[:B] for s in sk])
        want = dec([vec, [s[:B].repeat(S, 1, 1, 1) for s in sk]])
        got = codec.decode_shared(vec)
        assert got.is_contiguous()
        _close(got, want)
        buf = torch.empty_like(want)
        assert codec.decode_shared(vec, out=buf) is buf
        _close(buf, want)


def test_shared_skips_need_context_rows():
    enc, dec = _nets("dcgan_64", 1)
    codec = BatchedCodec(enc, dec, n_points=2)
    with torch.no_grad():
        _, sk = enc(torch.rand(4, 1, 64, 64))
    with pytest.raises(AssertionError):
        codec.set_shared_skips(sk)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
@pytest.mark.parametrize("model,nc,width", [("dcgan_64", 1, 64), ("vgg_64", 3, 64)])
def test_codec_drives_reference_model_classes(model, nc, width):
    """The codec walks sub-module names, so it must also work on the reference's own classes (unpickled checkpoints)."""
    sys.path.insert(0, REF)
    try:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        ref = importlib.import_module(f"models.{model}")
    finally:
        sys.path.remove(REF)
    torch.manual_seed(0)
    enc, dec = _randomise_bn(ref.encoder(90, nc), 1), _randomise_bn(ref.decoder(90, nc), 2)
    S, B = 2, 2
    codec = BatchedCodec(enc, dec, n_points=B)
    x = torch.rand(B, nc, width, width)
    vec = torch.tanh(torch.randn(S * B, 90))
    with torch.no_grad():
        h, sk = enc(x)
        ch, csk = codec.encode(x)
        _close(ch, h)
        codec.set_shared_skips(csk)
        _close(codec.decode_shared(vec), dec([vec, [s.repeat(S, 1, 1, 1) for s in sk]]))
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
