"""Evaluation metrics with the names and argument meaning of the reference's utils/metrics.py:11-54 (host-side helpers;
nothing here touches the device kernels).  `Y_std` is the standard deviation the targets were divided by, so that RMSE
and NLPD are reported in the units of the raw targets."""
from __future__ import annotations

import math

import torch


def get_trainable_param_names(model):
    """Names of the parameters an optimiser would update (utils/metrics.py:28-35)."""
    return [name for name, p in model.named_parameters() if p.requires_grad]


def print_trainable_param_names(model):
    """Two-column listing of trainable parameters and their sizes (utils/metrics.py:11-25; plain text instead of
    PrettyTable, which is not a dependency here)."""
    rows = [(name, p.numel()) for name, p in model.named_parameters() if p.requires_grad]
    width = max([len("Modules")] + [len(r[0]) for r in rows])
    print("%-*s | %s" % (width, "Modules", "Parameters"))
    for name, cnt in rows:
        print("%-*s | %d" % (width, name, cnt))
    print(f"Total Trainable Params: {sum(c for _, c in rows)}")


def rmse(Y_pred_mean, Y_test, Y_std):
    """Root-mean-square error rescaled by Y_std (utils/metrics.py:37-39)."""
    err = torch.sqrt(torch.mean((Y_pred_mean - Y_test) ** 2)).detach()
    return float(Y_std) * err


def nlpd(Y_test_pred, Y_test, Y_std):
    """Negative JOINT log predictive density per test point, corrected for the target rescaling
    (utils/metrics.py:41-46).  `Y_test_pred` exposes `.log_prob` (our MultivariateNormal or a torch distribution)."""
    lpd = Y_test_pred.log_prob(Y_test).detach()
    return -(lpd / len(Y_test) - torch.log(torch.as_tensor(Y_std, dtype=lpd.dtype, device=lpd.device)))


def negative_log_predictive_density(test_y, predicted_mean, predicted_var):
    """Mean negative marginal Gaussian log density (utils/metrics.py:49-54)."""
    z2 = (test_y - predicted_mean) ** 2 / predicted_var
    return torch.mean(0.5 * (z2 + torch.log(predicted_var) + math.log(2.0 * math.pi)))
