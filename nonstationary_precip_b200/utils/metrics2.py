"""Variant of utils.metrics used by the reference's DGP scripts (utils/metrics2.py:36-47): `rmse` is NOT rescaled by
Y_std there; everything else is shared."""
from __future__ import annotations

import torch

from .metrics import get_trainable_param_names, nlpd, print_trainable_param_names  # noqa: F401


def rmse(Y_pred_mean, Y_test, Y_std):
    return torch.sqrt(torch.mean((Y_pred_mean - Y_test) ** 2)).detach()
