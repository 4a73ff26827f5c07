"""Short legs for the other shardable paths of BASELINE.json, run by bench.py after the headline measurement and reported
under the JSON line's "configs" key (the headline `value` is untouched by them):

  c5  batched prediction: Gibbs posterior mean / variance on a 4096 x 4096 lat/lon grid, test rows sharded over the ranks,
      no collective (reference models/nonstationary_models.py:91-153, SURVEY.md 8(e) "Prediction")
  c4  2-layer DGP with DSVI, M = 512 per layer, S = 32 Monte-Carlo samples sharded over the ranks, one all-reduce of the
      gradients (reference models/dgps.py:72-98, experiments/deepgp_spatial_bench.py:84)
  c3  streamed SGPR bound + gradients of the Gibbs (lon, lat) part of the spatio-temporal model on a 2^20-row slice shaped
      like uib_spatio_temporal.csv, M = 2048, rows sharded, two all-reduces (reference models/gibbs_kernels.py:187-261)

Each leg carries its own sampled parity against the CPU oracle (rank 0; the oracle is the checker, never the thing timed).
Timing: CUDA events, max over ranks."""
from __future__ import annotations

import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
F64 = torch.float64


def _rel(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def _timed(fn, warm=1, iters=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        out = fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters, out


def leg_c5(rank, world, dev, allmax, make_params):
    from nonstationary_precip_b200.svgp import SVGPGibbs, _inv_softplus
    G, M, d = 4096, 1024, 2
    g = torch.Generator().manual_seed(5)
    kw = make_params("diag", M, d)
    kw["m"] = 0.3 * torch.randn(M, generator=g, dtype=F64)  # a trained-looking variational state (S != I)
    kw["Ls"] = 0.8 * torch.eye(M, dtype=F64) + 0.02 * torch.tril(torch.randn(M, M, generator=g, dtype=F64))
    Z = torch.rand(M, d, generator=g, dtype=F64) * 2 - 1
    model = SVGPGibbs("diag", Z.to(dev), 1 << 20, **{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in kw.items()})
    lin = torch.linspace(-1, 1, G, dtype=F64, device=dev)
    rows = G // world  # grid rows (latitudes) owned by this rank: G * rows test points, no collective
    xs = torch.stack(torch.meshgrid(lin[rank * rows:(rank + 1) * rows], lin, indexing="ij"), -1).reshape(-1, 2)
    model.predict(xs[:1 << 18].contiguous(), chunk=1 << 18)  # warm-up: buffers, lazy attributes
    ms, (mean, var) = _timed(lambda: model.predict(xs, chunk=1 << 18), warm=0, iters=1)
    ms = allmax(ms)
    out = {"what": "c5: %dx%d grid prediction, diagonal Gibbs SVGP, M=%d, %d test rows per rank, no collective" % (G, G, M, xs.shape[0]),
           "rows_per_s": G * G / ms * 1e3, "ms": ms, "finite": bool(torch.isfinite(mean).all() and torch.isfinite(var).all())}
    if rank == 0:
        from oracle import gibbs_oracle as o
        idx = torch.randperm(xs.shape[0], generator=torch.Generator().manual_seed(1))[:4096].to(dev)
        xc = xs[idx].cpu()
        mu_w, var_w = o.svgp_gibbs_predict(xc, Z, kw["m"], kw["Ls"], torch.tensor(_inv_softplus(kw["outputscale"]), dtype=F64),
                                           "diag", log_ell_z=kw["log_ell_z"], prior_c=kw["prior_c"], prior_os=kw["prior_os"],
                                           prior_lam=kw["prior_lam"])
        out["parity"] = {"rows": 4096, "mean_rel": _rel(mean[idx], mu_w), "var_rel": _rel(var[idx], var_w), "tol": 1e-6,
                         "against": "oracle.svgp_gibbs_predict on 4096 random grid rows"}
        out["parity"]["ok"] = bool(out["parity"]["mean_rel"] <= 1e-6 and out["parity"]["var_rel"] <= 1e-6)
    return out


def _dgp_layer_dict(layer):
    vs, vd = layer.variational_strategy, layer.variational_strategy._variational_distribution
    c = lambda t: t.detach().cpu().clone()
    d = dict(Z=c(vs.inducing_points), m=c(vd.variational_mean), Ls=c(vd.chol_variational_covar),
             raw_os=c(layer.covar_module.raw_outputscale), raw_ls=c(layer.covar_module.base_kernel.raw_lengthscale.squeeze(-2)))
    if hasattr(layer.mean_module, "weights"):
        d["W"], d["b"] = c(layer.mean_module.weights), c(layer.mean_module.bias)
    else:
        d["c"] = c(layer.mean_module.constant.reshape(()))
    return d


def leg_c4(rank, world, dev, allmax, all_reduce):
    from nonstationary_precip_b200.models import dgps
    B, S, M = 65536, 32, 512
    torch.manual_seed(4)  # identical initial parameters on every rank
    g = torch.Generator().manual_seed(4)
    x = (torch.rand(B, 3, generator=g, dtype=F64) * 2 - 1).to(dev)
    y = (torch.sin(3 * x[:, 0]) + 0.5 * torch.cos(5 * x[:, 1] * x[:, 2])).contiguous()
    model = dgps.DeepGP(1, x.shape, num_inducing=M).to(dev).double()
    mll = dgps.DeepApproximateMLL(dgps.VariationalELBO(model.likelihood, model, 1 << 20))
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    reduce_ = all_reduce if all_reduce is not None else (lambda t: t)
    k = [0]

    def step():
        opt.zero_grad(set_to_none=True)
        with dgps.num_likelihood_samples(S), dgps.sample_shard(rank, world):
            loss = -mll(model(x, seed=1000 + k[0]), y)
        loss.backward()
        dgps.allreduce_gradients(model, reduce_)
        opt.step()
        k[0] += 1
        return loss.detach()

    ms, loss = _timed(step, warm=2, iters=3)  # (the first two steps allocate the 7.5 GB digit workspace of the int8 path)
    ms = allmax(ms)
    # work actually executed per step (T = K C is cached between forward and backward): 4 (S + 2) B M^2
    out = {"what": "c4: 2-layer DGP, DSVI, B=%d, M=%d per layer, S=%d samples sharded over %d rank(s), fp64" % (B, M, S, world),
           "steps_per_s": 1e3 / ms, "ms_per_step": ms, "fp64_tflops_4(S+2)BM2": 4.0 * (S + 2) * B * M * M / ms / 1e9,
           "finite": bool(torch.isfinite(loss))}
    if rank == 0:
        from oracle import gibbs_oracle as o
        Bp, Sp = 1024, 8
        gp = torch.Generator().manual_seed(6)
        xp = torch.rand(Bp, 3, generator=gp, dtype=F64) * 2 - 1
        yp = torch.sin(3 * xp[:, 0]) + 0.1 * torch.randn(Bp, generator=gp, dtype=F64)
        eps = [torch.randn(Sp, Bp, 2, generator=gp, dtype=F64)]
        with torch.no_grad(), dgps.num_likelihood_samples(Sp):
            got = mll(model(xp.to(dev), eps=[e.to(dev) for e in eps]), yp.to(dev))
        raw_noise = model.likelihood.noise_covar.raw_noise.detach().cpu().clone()
        with torch.no_grad():
            want = o.dgp_elbo(xp, yp, 1 << 20, [_dgp_layer_dict(model.layers[0])], _dgp_layer_dict(model.last_layer), raw_noise[0], eps)
        out["parity"] = {"rows": Bp, "M": M, "S": Sp, "elbo_rel": abs(got.item() - want.item()) / abs(want.item()), "tol": 1e-6,
                         "against": "oracle.dgp_elbo, same parameters (after the timed steps) and the same N(0,1) draws"}
        out["parity"]["ok"] = bool(out["parity"]["elbo_rel"] <= 1e-6)
    return out


def leg_c3(rank, world, dev, allmax, all_reduce):
    from nonstationary_precip_b200.sgpr import SGPRGibbsStream
    N, M, D = 1 << 20, 2048, 2
    g = torch.Generator().manual_seed(3)
    cells = torch.stack(torch.meshgrid(torch.arange(64, dtype=F64), torch.arange(64, dtype=F64), indexing="ij"), -1).reshape(-1, 2)
    cells = (cells - cells.mean(0)) / cells.std(0)  # 64 x 64 cells at 0.25 degrees, z-scored (spatio_temporal_exp.py:45-49)
    x = cells[torch.arange(N) % 4096] + 0.01 * torch.randn(N, 2, generator=g, dtype=F64)  # time-major rows, as the CSV
    y = torch.exp(-(x ** 2).sum(-1)) * torch.sin(2 * math.pi * torch.arange(N, dtype=F64) / (4096.0 * 12.0))
    y = y + 0.05 * torch.randn(N, generator=g, dtype=F64)
    Z = x[torch.randperm(N, generator=g)[:M]].clone()
    hyp = dict(log_ell=torch.full((D, M), math.log(0.3), dtype=F64), c=torch.full((D,), math.log(0.3), dtype=F64),
               os=torch.ones(D, dtype=F64), lam=torch.full((D, D), 1.3, dtype=F64))
    mk = lambda: SGPRGibbsStream(Z.to(dev), hyp["log_ell"].to(dev), hyp["c"].to(dev), hyp["os"].to(dev), hyp["lam"].to(dev),
                                 outputscale=0.644, noise=0.05)
    model = mk()
    n_loc = N // world
    xs, ys = x[rank * n_loc:(rank + 1) * n_loc].to(dev), y[rank * n_loc:(rank + 1) * n_loc].to(dev)
    ms, loss = _timed(lambda: model.neg_objective_and_grad(xs, ys, chunk=65536, n_total=N, all_reduce=all_reduce), warm=1, iters=1)
    ms = allmax(ms)
    out = {"what": "c3 (Gibbs lon/lat part): streamed SGPR bound + gradients, %d rows (256 time steps x 64x64 cells), M=%d, rows "
                   "sharded over %d rank(s), two all-reduces" % (N, M, world),
           "rows_per_s": N / ms * 1e3, "ms_per_eval": ms, "fp64_equiv_tflops_22NM2": 22.0 * N * M * M / ms / 1e9,
           "finite": bool(torch.isfinite(loss))}
    if rank == 0:
        from oracle import gibbs_oracle as o
        ns = 16384
        sub = torch.randperm(N, generator=torch.Generator().manual_seed(2))[:ns]
        small = mk()
        got = small.neg_objective_and_grad(x[sub].to(dev), y[sub].to(dev), chunk=8192)
        Zc, lec = Z.clone().requires_grad_(True), hyp["log_ell"].clone().requires_grad_(True)
        t0 = time.perf_counter()
        want = -o.sgpr_gibbs_objective(x[sub], y[sub], Zc, lec, torch.tensor(0.644, dtype=F64), torch.tensor(0.05, dtype=F64),
                                       hyp["c"], hyp["os"], hyp["lam"])
        want.backward()
        out["parity"] = {"rows": ns, "M": M, "objective_rel": abs(got.item() - want.item()) / abs(want.item()),
                         "grad_Z_rel": _rel(small.Z.grad, Zc.grad), "grad_log_ell_rel": _rel(small.log_ell_z.grad, lec.grad),
                         "tol": 1e-6, "tol_grad": 1e-5, "oracle_seconds": round(time.perf_counter() - t0, 2),
                         "against": "oracle.sgpr_gibbs_objective + autograd on a %d-row random subset, same Z and parameters" % ns}
        # objective at the north-star tolerance; gradients at the 1e-5 the streamed-SGPR tests use (M = 2048: both sides lose
        # ~cond(Kzz) eps in dL/dZ)
        out["parity"]["ok"] = bool(out["parity"]["objective_rel"] <= 1e-6 and
                                   max(out["parity"]["grad_Z_rel"], out["parity"]["grad_log_ell_rel"]) <= 1e-5)
    return out


def run_all(rank, world, dev, allmax, all_reduce, make_params, which=("c5", "c4", "c3")):
    if os.path.join(ROOT, "tests") not in sys.path:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
    legs = {"c5": lambda: leg_c5(rank, world, dev, allmax, make_params), "c4": lambda: leg_c4(rank, world, dev, allmax, all_reduce),
            "c3": lambda: leg_c3(rank, world, dev, allmax, all_reduce)}
    out = {}
    for name in which:
        try:
            out[name] = legs[name]()
        except Exception as e:  # a broken leg must not take the headline line down; it is reported as such
            out[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        torch.cuda.empty_cache()
    return out
