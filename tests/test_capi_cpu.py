"""CPU-side checks of the boundary: the C-ABI library builds/loads and exports every symbol
include/dvg_b200.h declares; the drop-in classes keep the reference's names and pickle layout;
no compute is attempted without a GPU."""
import io
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from dvg_b200 import _capi
    lib = _capi.load()
    header = open(os.path.join(ROOT, "include", "dvg_b200.h")).read()
    declared = sorted(set(re.findall(r"DVG_API[^;]*?\b(dvg_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 19
    assert sorted(_capi.EXPORTS) == declared
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dvg_version() >= 100


def test_state_dict_names_match_reference_layout():
    from dvg_b200.models.gp_models import GaussianLikelihood, GPRegressionLayer1
    from dvg_b200.models.lstm import gaussian_lstm, lstm
    m = lstm(90, 90, 256, 2, 4)
    assert list(m.state_dict()) == [
        "embed.weight", "embed.bias", "lstm.0.weight_ih", "lstm.0.weight_hh", "lstm.0.bias_ih", "lstm.0.bias_hh",
        "lstm.1.weight_ih", "lstm.1.weight_hh", "lstm.1.bias_ih", "lstm.1.bias_hh", "output.0.weight",
        "output.0.bias"]
    g = gaussian_lstm(90, 10, 256, 1, 4)
    assert list(g.state_dict())[-4:] == ["mu_net.weight", "mu_net.bias", "logvar_net.weight", "logvar_net.bias"]
    gp = GPRegressionLayer1(90, 40)
    sd = gp.state_dict()
    assert sd["variational_strategy.inducing_points"].shape == (90, 40, 1)
    assert sd["variational_strategy.variational_distribution.variational_mean"].shape == (90, 40)
    assert sd["variational_strategy.variational_distribution.chol_variational_covar"].shape == (90, 40, 40)
    assert sd["variational_strategy.variational_params_initialized"].dim() == 0
    assert sd["mean_module.constant"].shape == (90, 1)
    assert sd["covar_module.raw_outputscale"].shape == (90,)
    assert sd["covar_module.base_kernel.raw_lengthscale"].shape == (90, 1, 1)
    assert GaussianLikelihood(90).state_dict()["noise_covar.raw_noise"].shape == (90, 1)


def test_golden_state_dicts_load_and_module_pickles(tmp_path):
    """Checkpoints pickle the whole frame_predictor (train.py:380-383): round-trip must work and must not
    drag runtime handles along; the pickled __dict__ carries the reference's fields."""
    from dvg_b200.models.lstm import lstm
    g = torch.load(os.path.join(ROOT, "tests", "golden", "lstm_tiny.pt"), weights_only=False)
    gi, go, H, L, B = g["dims"]
    m = lstm(gi, go, H, L, B)
    m.load_state_dict(g["state_dict"])
    buf = io.BytesIO()
    torch.save({"frame_predictor": m}, buf)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)["frame_predictor"]
    assert isinstance(m2, lstm)
    for k in ("input_size", "output_size", "hidden_size", "batch_size", "n_layers", "hidden"):
        assert k in m2.__dict__
    assert "_dvg_rt" not in m2.__dict__
    for k, v in m.state_dict().items():
        assert torch.equal(v, m2.state_dict()[k])


def test_install_dropin_aliases_reference_import_paths():
    import sys
    saved = {k: sys.modules.get(k) for k in ("models", "models.lstm", "models.gp_models", "gp_models")}
    try:
        for k in saved:
            sys.modules.pop(k, None)
        import dvg_b200
        dvg_b200.install_dropin()
        import models.lstm as ml
        from dvg_b200.models import lstm as ours
        assert ml.lstm is ours.lstm and ml.gaussian_lstm is ours.gaussian_lstm
        import models.gp_models as mg
        assert hasattr(mg, "GPRegressionLayer1")
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from dvg_b200._capi import DvgError
    from dvg_b200.models.lstm import lstm
    m = lstm(12, 12, 32, 1, 2).eval()
    with torch.no_grad(), pytest.raises(DvgError):
        m(torch.zeros(2, 12))
