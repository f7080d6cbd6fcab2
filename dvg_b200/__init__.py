"""dvg_b200 -- B200-native (sm_100a) implementation of DVG's per-timestep stochastic rollout.

Only what the hot path needs lives here:
  csrc/            hand-written CUDA kernels + the C ABI (include/dvg_b200.h)
  _capi.py         ctypes binding of the C ABI (no torch types cross it)
  models/          drop-in mirrors of the reference's models/lstm.py and models/gp_models.py
  rollout.py       sample-batched rollout drivers (make_gifs / GPtrigger_gen / plot bookkeeping)
  shard.py         one-process-per-GPU sharding of the diverse samples + the final NCCL gather
  convnets.py      the reference's encoder/decoder conv stacks on the stock PyTorch path (harness only)
"""
from . import _capi  # noqa: F401

__version__ = "0.1.0"


def install_dropin():
    """Make ``import models.lstm`` / ``import models.gp_models`` (the paths the reference scripts and
    its pickled checkpoints use, train.py:76,380-383; generate_frames.py:14,59-61) resolve to the
    B200-native classes.  Call before ``torch.load`` of a reference checkpoint."""
    import sys
    import types

    from .models import gp_models, lstm
    pkg = sys.modules.get("models")
    if pkg is None:
        pkg = types.ModuleType("models")
        pkg.__path__ = []
        sys.modules["models"] = pkg
    pkg.lstm = lstm
    pkg.gp_models = gp_models
    sys.modules["models.lstm"] = lstm
    sys.modules["models.gp_models"] = gp_models
    sys.modules.setdefault("gp_models", gp_models)   # generate_frames.py:14 imports it without the prefix
    # train.py / generate_frames.py also ``import gpytorch`` (likelihoods.GaussianLikelihood, mlls.VariationalELBO,
    # settings.*): where the real library is absent a shim with exactly those names takes its place
    try:
        import gpytorch  # noqa: F401
    except ImportError:
        from .models import gp_train
        shim = gp_train.gpytorch_shim()
        sys.modules["gpytorch"] = shim
        for sub in ("likelihoods", "mlls", "settings"):
            sys.modules["gpytorch." + sub] = getattr(shim, sub)
    return pkg
