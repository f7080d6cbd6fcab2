"""Host-side mirror of the reference's ``models`` package for the hot path
(``models.lstm`` -> lstm / gaussian_lstm, ``models.gp_models`` -> GPRegressionLayer1)."""
from . import gp_models, lstm  # noqa: F401
