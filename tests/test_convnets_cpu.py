"""The harness conv nets interchange state_dicts with the reference modules and compute the same function.
Needs /root/reference (build container only); skipped on the GPU box."""
import importlib
import os
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")


@pytest.mark.parametrize("model,nc,width", [("dcgan_64", 1, 64), ("vgg_64", 1, 64), ("dcgan_128", 3, 128), ("vgg_128", 3, 128)])
def test_state_dict_and_outputs_match_reference(model, nc, width):
    from dvg_b200.convnets import make_codec
    sys.path.insert(0, REF)
    try:
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        ref = importlib.import_module(f"models.{model}")
    finally:
        sys.path.remove(REF)
    torch.manual_seed(0)
    r_enc, r_dec = ref.encoder(90, nc).eval(), ref.decoder(90, nc).eval()
    enc, dec = make_codec(model, 90, nc)
    assert list(enc.state_dict().keys()) == list(r_enc.state_dict().keys())
    assert list(dec.state_dict().keys()) == list(r_dec.state_dict().keys())
    enc.load_state_dict(r_enc.state_dict())
    dec.load_state_dict(r_dec.state_dict())
    enc.eval(); dec.eval()
    x = torch.rand(2, nc, width, width)
    with torch.no_grad():
        h, skips = enc(x)
        rh, rskips = r_enc(x)
        assert torch.equal(h, rh) and all(torch.equal(a, b) for a, b in zip(skips, rskips))
        assert torch.equal(dec([h, skips]), r_dec([rh, rskips]))
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
        del sys.modules[k]
