"""Data preparation with the function names of the reference's utils/dataprep.py:9-52 (CSV table whose LAST column is the
target), plus `kmeans_inducing_points`, the replacement for `pymc3.gp.util.kmeans_inducing_points` that the experiment
scripts use to place inducing points (experiments/spatial_exp.py:153)."""
from __future__ import annotations

import math

import numpy as np
import torch


def download_data(filepath):
    """CSV -> float32 tensor of all columns (utils/dataprep.py:9-12: `torch.Tensor(df.values)` is float32)."""
    import pandas as pd
    return torch.tensor(pd.read_csv(filepath).values, dtype=torch.float32)


def prep_inputs(data):
    """z-score every input column (all but the last) with the unbiased std (utils/dataprep.py:14-22)."""
    x = data[:, :-1]
    std, mean = torch.std_mean(x, dim=-2)
    return (x - mean) / std


def prep_outputs(data):
    """Box-Cox transform of the target column, lambda by maximum likelihood (utils/dataprep.py:24-29)."""
    from scipy import stats
    y_tr, lam = stats.boxcox(np.asarray(data[:, -1]))
    return y_tr, lam


def box_cox_transform(data):
    return prep_inputs(data), prep_outputs(data)


def whitening_transform(data):
    """z-score inputs and target; also returns the statistics (utils/dataprep.py:35-43)."""
    x, y = data[:, :-1], data[:, -1]
    stdx, meanx = torch.std_mean(x, dim=-2)
    stdy, meany = torch.std_mean(y)
    return (x - meanx) / stdx, (y - meany) / stdy, meanx, stdx, meany, stdy


def train_test_split(X, y, train_prop):
    """First floor(train_prop * n) rows train, the rest test (utils/dataprep.py:45-52; no shuffling)."""
    k = int(math.floor(train_prop * len(X)))
    return X[:k, :].contiguous(), y[:k].contiguous(), X[k:, :].contiguous(), y[k:].contiguous()


def kmeans_inducing_points(n_inducing, X, iters: int = 50, seed: int = 0):
    """Inducing locations = k-means centroids of X, computed on per-column std-scaled inputs and scaled back (what
    pymc3.gp.util.kmeans_inducing_points does with scipy's kmeans, recalled).  Lloyd iterations in torch on X's device;
    columns with zero spread are left unscaled; empty clusters are re-seeded on the farthest points.  Deterministic
    for a given seed.  Returns (n_inducing, d) in X's dtype."""
    Xt = torch.as_tensor(X)
    if Xt.dim() != 2:
        raise ValueError("X must be (n, d)")
    n = Xt.shape[0]
    if n_inducing > n:
        raise ValueError("more inducing points than data rows")
    Xd = Xt.to(torch.float64)
    scale = Xd.std(0)
    scale = torch.where(scale > 0, scale, torch.ones_like(scale))
    Xw = Xd / scale
    g = torch.Generator().manual_seed(seed)
    C = Xw[torch.randperm(n, generator=g)[:n_inducing].to(Xw.device)].clone()
    for _ in range(iters):
        d2 = (Xw * Xw).sum(1, keepdim=True) - 2.0 * Xw @ C.T + (C * C).sum(1)
        mind, lab = d2.min(1)
        sums = torch.zeros_like(C).index_add_(0, lab, Xw)
        cnt = torch.zeros(n_inducing, dtype=torch.float64, device=Xw.device).index_add_(0, lab, torch.ones_like(mind))
        newC = torch.where(cnt[:, None] > 0, sums / cnt.clamp_min(1.0)[:, None], C)
        empty = (cnt == 0).nonzero().reshape(-1)
        if empty.numel():
            newC[empty] = Xw[torch.topk(mind, empty.numel()).indices]
        shift = (newC - C).abs().max()
        C = newC
        if float(shift) < 1e-12:
            break
    return (C * scale).to(Xt.dtype)
