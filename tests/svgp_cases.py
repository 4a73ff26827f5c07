"""Shared seeded problem builders for the SVGP-Gibbs tests (CPU emulator tier and GPU tier) and bench.py."""
import math

import torch


def make_problem(variant, B=96, M=24, d=3, seed=0, device="cpu", N_total=None):
    g = torch.Generator().manual_seed(173 + seed)
    f64 = torch.float64
    N_total = N_total or 4 * B
    x = torch.rand(B, d, generator=g, dtype=f64) * 2 - 1
    y = torch.sin(3 * x[:, 0]) + 0.5 * torch.cos(5 * x[:, 1] * x[:, -1]) + 0.1 * torch.randn(B, generator=g, dtype=f64)
    Z = torch.rand(M, d, generator=g, dtype=f64) * 2 - 1
    m = 0.3 * torch.randn(M, generator=g, dtype=f64)
    Ls = torch.eye(M, dtype=f64) * 0.8 + 0.05 * torch.tril(torch.randn(M, M, generator=g, dtype=f64))
    kw = dict(m=m, Ls=Ls, outputscale=0.644, noise=0.011)
    if variant == "diag":
        kw.update(log_ell_z=math.log(0.3) + 0.1 * torch.randn(d, M, generator=g, dtype=f64),
                  prior_c=torch.full((d,), math.log(0.3), dtype=f64), prior_os=torch.ones(d, dtype=f64),
                  prior_lam=torch.full((d, d), 1.3, dtype=f64))
    else:
        Dd = torch.randn(d, generator=g, dtype=f64)
        Dd = torch.sign(Dd) * Dd.abs().clamp_min(0.7)  # 3-D Sigma(h) is only PD for |D_kk| >~ 0.6 (softplus part has min eig ~ -0.35)
        kw.update(H=torch.randn(M, d, generator=g, dtype=f64), Dm=torch.diag(Dd), row_os=1.0,
                  row_lam=torch.ones(d, dtype=f64))
    mv = lambda t: t.to(device) if torch.is_tensor(t) else t
    return mv(x), mv(y), mv(Z), {k: mv(v) for k, v in kw.items()}, N_total


def oracle_loss_and_grads(variant, x, y, Z, kw, N_total, include_prior=True, learn_z=True):
    """-ELBO and its gradients from the CPU oracle with autograd (all tensors moved to CPU)."""
    from oracle import gibbs_oracle as o
    from nonstationary_precip_b200.svgp import _inv_softplus
    c = lambda t: t.detach().cpu().clone()
    P = dict(Z=c(Z).requires_grad_(True), m=c(kw["m"]).requires_grad_(True), Ls=c(kw["Ls"]).requires_grad_(True),
             raw_outputscale=torch.tensor([_inv_softplus(kw["outputscale"])], dtype=torch.float64, requires_grad=True),
             raw_noise=torch.tensor([_inv_softplus(kw["noise"] - 1e-4)], dtype=torch.float64, requires_grad=True))
    extra = {}
    if variant == "diag":
        P["log_ell_z"] = c(kw["log_ell_z"]).requires_grad_(True)
        extra = dict(log_ell_z=P["log_ell_z"], prior_c=c(kw["prior_c"]), prior_os=c(kw["prior_os"]),
                     prior_lam=c(kw["prior_lam"]))
    else:
        P["H"] = c(kw["H"]).requires_grad_(True)
        P["D"] = c(kw["Dm"]).requires_grad_(True)
        extra = dict(H=P["H"], Dm=P["D"], row_os=torch.tensor(float(kw["row_os"]), dtype=torch.float64),
                     row_lam=c(kw["row_lam"]))
    elbo = o.svgp_gibbs_elbo(c(x), c(y), N_total, P["Z"], P["m"], P["Ls"], P["raw_outputscale"][0], P["raw_noise"][0],
                             variant, include_prior=include_prior, **extra)
    loss = -elbo
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in P.items()}
    grads["Ls"] = torch.tril(grads["Ls"])
    if not learn_z:
        grads["Z"] = torch.zeros_like(grads["Z"])
    return loss.detach(), grads
