#!/usr/bin/env python3
"""Timings for the other BASELINE.json configurations (the headline config 2 is bench.py):

  c1  exact-GP MAP step, diagonal Gibbs kernel, n = 316, D = 2 (shape of experiments/spatial_exp.py; synthetic data)
  c4  2-layer DGP with DSVI (models/dgps.py), N = 2^20, M = 512 per layer, S = 32 samples, minibatch B (default 65536)
  c5  batched prediction on a 4096 x 4096 lat/lon grid (diagonal Gibbs SVGP posterior, M = 1024), rows of this rank

Each prints one JSON object.  Run on the GPU box:  python tools/bench_configs.py c5 c4 c1"""
import json
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

f64 = dict(dtype=torch.float64, device="cuda")


def ev_time(fn, iters=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


def c5():
    from nonstationary_precip_b200.svgp import SVGPGibbs
    G = int(os.environ.get("GRID", 4096))
    world, rank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0))
    M, d = 1024, 2
    kw = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in bench.make_params("diag", M, d).items()}
    g = torch.Generator().manual_seed(5)
    Z = (torch.rand(M, d, generator=g, dtype=torch.float64) * 2 - 1).cuda()
    model = SVGPGibbs("diag", Z, 1 << 20, **kw)
    lin = torch.linspace(-1, 1, G, **f64)
    rows = G // world  # rows of the grid owned by this rank (row-sharded, no collective)
    lat = lin[rank * rows:(rank + 1) * rows]
    xs = torch.stack(torch.meshgrid(lat, lin, indexing="ij"), -1).reshape(-1, 2)
    best, med = ev_time(lambda: model.predict(xs, chunk=1 << 18), iters=2, warm=1)
    n = xs.shape[0]
    mean, var = model.predict(xs[:4096].contiguous())
    print(json.dumps({"config": "c5 grid prediction %dx%d, M=%d, diag Gibbs, rows on this rank %d" % (G, G, M, n),
                      "ms": best, "rows_per_s": n / best * 1e3, "tflops_rowquad_equiv": 2.0 * n * M * M / best / 1e9,
                      "finite": bool(torch.isfinite(mean).all() and torch.isfinite(var).all())}), flush=True)


def c4():
    from nonstationary_precip_b200.models import dgps
    B = int(os.environ.get("B", 65536))
    S = int(os.environ.get("S", 32))
    M = int(os.environ.get("M", 512))
    torch.manual_seed(4)
    g = torch.Generator().manual_seed(4)
    x = (torch.rand(B, 3, generator=g, dtype=torch.float64) * 2 - 1).cuda()
    y = (torch.sin(3 * x[:, 0]) + 0.5 * torch.cos(5 * x[:, 1] * x[:, 2])).contiguous()
    model = dgps.DeepGP(1, x.shape, num_inducing=M).cuda().double()
    mll = dgps.DeepApproximateMLL(dgps.VariationalELBO(model.likelihood, model, 1 << 20))
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    step_no = [0]

    def step():
        opt.zero_grad(set_to_none=True)
        with dgps.num_likelihood_samples(S):
            out = model(x, seed=1000 + step_no[0])
            loss = -mll(out, y)
        loss.backward()
        opt.step()
        step_no[0] += 1
        return loss

    best, med = ev_time(step, iters=3, warm=2)
    loss = step().item()
    flop = 6.0 * (S + 2) * B * M * M
    print(json.dumps({"config": "c4 2-layer DGP DSVI, B=%d, M=%d/layer, S=%d, fp64, eager autograd over npgp kernels" % (
        B, M, S), "ms_per_step": best, "steps_per_s": 1e3 / best, "tflops_vs_6(S+2)BM2": flop / best / 1e9,
        "loss": loss, "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}), flush=True)


def c3():
    """Spatial (Gibbs) part of config 3 at full scale: streamed SGPR objective + gradients, N = 4 194 304 rows shaped like
    uib_spatio_temporal.csv (1024 time steps x 64x64 cells; here the lon/lat columns, z-scored), M = 2048."""
    from nonstationary_precip_b200.sgpr import SGPRGibbsStream
    N = int(os.environ.get("N", 1 << 22))
    M = int(os.environ.get("M", 2048))
    world, rank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0))
    g = torch.Generator().manual_seed(3)
    cells = torch.stack(torch.meshgrid(torch.arange(64, dtype=torch.float64), torch.arange(64, dtype=torch.float64),
                                       indexing="ij"), -1).reshape(-1, 2)
    cells = (cells - cells.mean(0)) / cells.std(0)
    idx = torch.arange(N) % 4096
    x = (cells[idx] + 0.01 * torch.randn(N, 2, generator=g, dtype=torch.float64)).cuda()
    y = (torch.exp(-(x ** 2).sum(-1)) * torch.sin(2 * math.pi * torch.arange(N, device="cuda") / (4096.0 * 12.0))).contiguous()
    Z = x[torch.randperm(N, generator=g)[:M].cuda()].clone()
    D = 2
    model = SGPRGibbsStream(Z, torch.full((D, M), math.log(0.3), **f64), torch.full((D,), math.log(0.3), **f64),
                            torch.ones(D, **f64), torch.full((D, D), 1.3, **f64), outputscale=0.644, noise=0.05)
    n_loc = N // world
    xs, ys = x[rank * n_loc:(rank + 1) * n_loc], y[rank * n_loc:(rank + 1) * n_loc]
    best, med = ev_time(lambda: model.neg_objective_and_grad(xs, ys, chunk=65536, n_total=N), iters=2, warm=1)
    loss = model.neg_objective_and_grad(xs, ys, chunk=65536, n_total=N).item()
    flop = 22.0 * n_loc * M * M  # whitening 2+2, SYRK 4, dG 8, dK 2, dP 4 (x n M^2); FP64-equivalent (int8 GEMMs inside)
    print(json.dumps({"config": "c3 (spatial Gibbs part) streamed SGPR objective+grad, N=%d, M=%d, rows on this rank %d" % (
        N, M, n_loc), "ms_per_eval": best, "evals_per_s": 1e3 / best, "fp64_equiv_tflops_22NM2": flop / best / 1e9, "loss": loss,
        "finite_grads": bool(all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)),
        "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}), flush=True)


def c3st():
    """Config 3 in full: N = 4 194 304 rows = 1024 monthly time steps x 64x64 cells (time, lon, lat; z-scored columns,
    time-major as uib_spatio_temporal.csv), M = 2048, kernel (>=7)-scaled RBF x Periodic (time) + scaled Gibbs (lon, lat),
    streamed collapsed bound + all gradients (rank-2M root never formed)."""
    from nonstationary_precip_b200.sgpr import SGPRSpatioTemporalStream
    T_, Gs = int(os.environ.get("T", 1024)), int(os.environ.get("GS", 64))
    M = int(os.environ.get("M", 2048))
    chunk = int(os.environ.get("CHUNK", 32768))
    world, rank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0))
    N = T_ * Gs * Gs
    g = torch.Generator().manual_seed(3)
    tt = 2000.0 + (torch.arange(T_, dtype=torch.float64) + 0.5) / 12.0
    lon = 72.25 + 0.25 * torch.arange(Gs, dtype=torch.float64)
    lat = 34.0 + 0.25 * torch.arange(Gs, dtype=torch.float64)
    grid = torch.stack(torch.meshgrid(tt, lon, lat, indexing="ij"), -1).reshape(-1, 3)
    raw_t = grid[:, 0].clone()
    grid = (grid - grid.mean(0)) / grid.std(0)
    # break exact ties (inducing points are a subset of the rows; duplicated times / cells make Kzz singular)
    x = (grid + 1e-3 * torch.randn(N, 3, generator=g, dtype=torch.float64)).cuda()
    y = (torch.sin(2 * math.pi * raw_t).cuda() * torch.exp(-(x[:, 1:] ** 2).sum(-1))
         + 0.1 * torch.randn(N, generator=g, dtype=torch.float64).cuda()).contiguous()
    Z = x[torch.randperm(N, generator=g)[:M].cuda()].clone()
    period = 1.0 / float(raw_t.std())  # one year in z-scored time units
    model = SGPRSpatioTemporalStream(Z, torch.full((2, M), math.log(0.3), **f64), torch.full((2,), math.log(0.3), **f64),
                                     torch.ones(2, **f64), torch.full((2, 2), 1.3, **f64),
                                     hyp_t=(1.0, 1.0, period, 7.7), outputscale_s=0.644, noise=0.05)
    n_loc = N // world
    xs, ys = x[rank * n_loc:(rank + 1) * n_loc], y[rank * n_loc:(rank + 1) * n_loc]
    best, med = ev_time(lambda: model.neg_objective_and_grad(xs, ys, chunk=chunk, n_total=N), iters=1, warm=1)
    loss = model.neg_objective_and_grad(xs, ys, chunk=chunk, n_total=N).item()
    flop = 22.0 * n_loc * M * M  # whitening 2+2, SYRK 4, dG 8, dF 2, dP 4 (x n M^2), triangular factors exploited
    print(json.dumps({"config": "c3 spatio-temporal (temporal + Gibbs Nystrom sum) streamed SGPR objective+grad, N=%d, M=%d "
                                "(rank-2M), rows on this rank %d" % (N, M, n_loc), "ms_per_eval": best,
                      "evals_per_s": 1e3 / best, "tflops_22NM2": flop / best / 1e9, "loss": loss,
                      "finite_grads": bool(all(torch.isfinite(p.grad).all() for p in model.parameters()
                                               if p.grad is not None)),
                      "max_mem_GB": torch.cuda.max_memory_allocated() / 1e9}), flush=True)


def c1():
    from nonstationary_precip_b200.gp_base import ExactMarginalLogLikelihood, GaussianLikelihood
    from nonstationary_precip_b200.models.gibbs_kernels import LogNormalPriorProcess
    from nonstationary_precip_b200.models.nonstationary_models import DiagonalExactGP
    g = torch.Generator().manual_seed(1)
    n, D = 316, 2
    x = torch.randn(n, D, generator=g, dtype=torch.float64).cuda()
    y = (torch.sin(2 * x[:, 0]) * torch.cos(x[:, 1])).contiguous()
    prior = LogNormalPriorProcess(input_dim=D).cuda().double()
    prior.covar_module.base_kernel.lengthscale = 1.3 * torch.ones_like(prior.covar_module.base_kernel.lengthscale)
    prior.mean_module.constant = torch.nn.Parameter(math.log(0.3) * torch.ones_like(prior.mean_module.constant))
    for p in prior.parameters():
        p.requires_grad = False
    lik = GaussianLikelihood().cuda().double()
    model = DiagonalExactGP(x, y, lik, prior, num_dim=D).cuda().double()
    model.likelihood.noise = 0.011
    model.covar_module.outputscale = 0.644
    for p in list(lik.parameters()) + [model.covar_module.raw_outputscale]:
        p.requires_grad = False
    opt = torch.optim.Adam([model.log_ell_train_x], lr=0.01)
    mll = ExactMarginalLogLikelihood(lik, model)
    model.train()

    def step():
        opt.zero_grad()
        loss = -mll(model(x), y)
        loss.backward()
        opt.step()
        return loss

    l0 = step().item()
    t0 = time.perf_counter()
    for _ in range(50):
        loss = step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 50
    print(json.dumps({"config": "c1 exact-GP MAP step n=316 D=2 (spatial_exp.py shape, synthetic data), eager",
                      "ms_per_step": dt * 1e3, "steps_per_s": 1 / dt, "loss_first": l0, "loss_last": loss.item()}),
          flush=True)


if __name__ == "__main__":
    for name in (sys.argv[1:] or ["c5", "c4", "c1"]):
        globals()[name]()
