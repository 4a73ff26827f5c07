"""The mirrored reference classes (nonstationary_precip_b200/models) on the CUDA kernels against the CPU oracle and the
golden fixtures produced from the reference's own source."""
import math

import pytest
import torch

from oracle import gibbs_oracle as o

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)


def rel(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def make_prior(D, os=1.0, lam=1.3, mean=0.3, device="cuda", active_dims=None):
    from nonstationary_precip_b200.models.gibbs_kernels import LogNormalPriorProcess
    prior = LogNormalPriorProcess(input_dim=D, active_dims=active_dims).to(device).double()
    # the reference's way of setting the hyper-parameters (experiments/spatial_exp.py:159-167)
    prior.covar_module.outputscale = os * torch.ones_like(prior.covar_module.outputscale)
    prior.covar_module.base_kernel.lengthscale = lam * torch.ones_like(prior.covar_module.base_kernel.lengthscale)
    prior.mean_module.constant = torch.nn.Parameter(math.log(mean) * torch.ones_like(prior.mean_module.constant))
    for p in prior.parameters():
        p.requires_grad = False
    return prior


@pytest.mark.parametrize("tag", ["d2", "d3"])
def test_lognormal_prior_process_golden(golden, tag):
    from nonstationary_precip_b200.models.gibbs_kernels import GibbsKernel, LogNormalPriorProcess
    g = golden("lognormal_field_" + tag)
    D = g["c"].shape[0]
    prior = LogNormalPriorProcess(input_dim=D).cuda().double()
    prior.covar_module.outputscale = g["os"].cuda()
    prior.covar_module.base_kernel.lengthscale = g["lam"].cuda().unsqueeze(1)
    prior.mean_module.constant.data = g["c"].cuda().unsqueeze(-1)
    x, xg, ell_g = g["x"].cuda(), g["xg"].cuda(), g["ell_g"].cuda()
    with torch.no_grad():
        assert rel(prior.conditional_sample(x, given=(xg, ell_g)), g["ell_x"]) < 1e-10
        assert rel(prior.log_prob((xg, torch.log(ell_g))), g["log_prob"]) < 1e-9
        K = GibbsKernel(lengthscale_prior=prior).forward(xg, x, ell1=ell_g)  # conditional branch (ell2=None)
        assert rel(K, g["K_cond"]) < 1e-10


def test_lognormal_prior_active_dims_golden(golden):
    """LogNormalPriorProcess(input_dim=2, active_dims=(0,1)) as the spatio-temporal experiment builds it: log_prob on the
    full (M,3) inducing points uses columns (time, lon) like gpytorch.Kernel.__call__; fixture from the reference's lines."""
    from nonstationary_precip_b200.models.gibbs_kernels import LogNormalPriorProcess
    g = golden("lognormal_prior_active_dims")
    prior = LogNormalPriorProcess(input_dim=2, active_dims=(0, 1)).cuda().double()
    prior.covar_module.outputscale = g["os"].cuda()
    prior.covar_module.base_kernel.lengthscale = g["lam"].cuda().unsqueeze(1)
    prior.mean_module.constant.data = g["c"].cuda().unsqueeze(-1)
    Z3, x2, log_ell = g["Z3"].cuda(), g["x2"].cuda(), g["log_ell"].cuda()
    with torch.no_grad():
        assert rel(prior.log_prob((Z3, log_ell)), g["log_prob_full_Z"]) < 1e-9
        assert rel(prior.log_prob((Z3[:, 1:3].contiguous(), log_ell)), g["log_prob_lonlat"]) < 1e-9
        assert rel(prior.conditional_sample(x2, given=(Z3[:, 1:3].contiguous(), torch.exp(log_ell))), g["ell_x"]) < 1e-10


def test_exact_gp_map_objective_gradient_and_predict():
    """DiagonalExactGP + ExactMarginalLogLikelihood (config 1 shapes: n = 316, D = 2) vs the oracle."""
    from nonstationary_precip_b200.gp_base import ExactMarginalLogLikelihood, GaussianLikelihood
    from nonstationary_precip_b200.models.nonstationary_models import DiagonalExactGP
    g = torch.Generator().manual_seed(3)
    n, ns, D = 316, 78, 2
    x = torch.randn(n, D, generator=g)
    y = torch.sin(2 * x[:, 0]) * torch.cos(x[:, 1]) + 0.1 * torch.randn(n, generator=g)
    xs = torch.randn(ns, D, generator=g)
    prior = make_prior(D)
    lik = GaussianLikelihood().cuda().double()
    model = DiagonalExactGP(x.cuda(), y.cuda(), lik, prior, num_dim=D).cuda().double()
    model.likelihood.noise = 0.011
    model.covar_module.outputscale = 0.644
    with torch.no_grad():
        model.log_ell_train_x += 0.2 * torch.randn(D, n, generator=g).cuda()
    mll = ExactMarginalLogLikelihood(lik, model)
    model.train()
    loss = -mll(model(model.train_inputs[0]), model.train_targets)
    loss.backward()
    c = torch.full((D,), math.log(0.3))
    os_, lam = torch.ones(D), torch.full((D, D), 1.3)
    le = model.log_ell_train_x.detach().cpu().clone().requires_grad_(True)
    want = -o.exact_gp_map_objective(x, y, le, torch.tensor(0.644), torch.tensor(0.011), c, os_, lam)
    want.backward()
    assert abs(loss.item() - want.item()) < 1e-9 * abs(want.item())
    assert rel(model.log_ell_train_x.grad, le.grad) < 1e-6
    model.eval()
    with torch.no_grad():
        pred = model.predict(xs.cuda())
    mu, sigma = o.exact_gp_predict(x, y, le.detach(), xs, torch.tensor(0.644), torch.tensor(0.011), c, os_, lam)
    assert rel(pred.mean, mu) < 1e-7
    assert rel(pred.covariance_matrix, sigma) < 1e-7


def test_sgpr_objective_gradient_and_predict():
    """DiagonalSparseGP (collapsed SGPR bound + trace term + prior) vs the oracle."""
    from nonstationary_precip_b200.gp_base import ExactMarginalLogLikelihood, GaussianLikelihood
    from nonstationary_precip_b200.models.nonstationary_models import DiagonalSparseGP
    g = torch.Generator().manual_seed(4)
    n, ns, M, D = 500, 60, 64, 2
    x = torch.rand(n, D, generator=g) * 2 - 1
    y = torch.sin(3 * x[:, 0]) + 0.1 * torch.randn(n, generator=g)
    xs = torch.rand(ns, D, generator=g) * 2 - 1
    z = x[torch.randperm(n, generator=g)[:M]].clone()
    prior = make_prior(D)
    lik = GaussianLikelihood().cuda().double()
    model = DiagonalSparseGP(x.cuda(), y.cuda(), lik, prior, z.cuda(), num_dim=D).cuda().double()
    model.likelihood.noise = 0.05
    model.covar_module.outputscale = 0.8
    with torch.no_grad():
        model.log_ell_z += 0.2 * torch.randn(D, M, generator=g).cuda()
    mll = ExactMarginalLogLikelihood(lik, model)
    model.train()
    loss = -mll(model(model.train_inputs[0]), model.train_targets)
    loss.backward()
    c, os_, lam = torch.full((D,), math.log(0.3)), torch.ones(D), torch.full((D, D), 1.3)
    le = model.log_ell_z.detach().cpu().clone().requires_grad_(True)
    zc = z.clone().requires_grad_(True)
    want = -o.sgpr_gibbs_objective(x, y, zc, le, torch.tensor(0.8), torch.tensor(0.05), c, os_, lam)
    want.backward()
    assert abs(loss.item() - want.item()) < 1e-8 * abs(want.item())
    assert rel(model.log_ell_z.grad, le.grad) < 1e-5
    assert rel(model.covar_module.base_kernel.inducing_points.grad, zc.grad) < 1e-5
    with pytest.raises(RuntimeError, match="x1 should equal x2 in training mode"):
        model.covar_module.base_kernel.forward(x.cuda()[:10], x.cuda()[10:30], ell=torch.exp(model.log_ell_z))
    model.eval()
    with torch.no_grad():
        pred = model.predict(xs.cuda())
    mu, var = o.sgpr_gibbs_predict(x, y, z, le.detach(), xs, torch.tensor(0.8), torch.tensor(0.05), c, os_, lam)
    assert rel(pred.mean, mu) < 1e-6
    assert rel(pred.variance, var) < 1e-6


def test_sparse_multivariate_kernel_golden(golden):
    from nonstationary_precip_b200.models.sparse_multivariate_gibbs_kernel import SparseMultivariateGibbsKernel
    g = golden("sparse_multivariate_d2")
    Z, x = g["Z"].cuda(), g["x"].cuda()
    torch.manual_seed(0)
    k = SparseMultivariateGibbsKernel(Z, 2, Z.clone())
    k.H.data = g["H"].cuda()
    k.D.data = g["Dm"].cuda()
    with torch.no_grad():
        assert rel(k.expectation_conditional_matrix_variate_dist(x), g["Hx"]) < 1e-8
        assert rel(k.forward(x, Z), g["Kxz"]) < 1e-8
        assert rel(k.forward(Z, Z), g["Kzz"]) < 1e-11
        assert rel(k.forward(x, x), g["Kxx"]) < 1e-8
        lp = k.H_matrix_prior.log_prob(k.H.data)
        assert abs(lp.item() - g["prior_H_log_prob"].item()) < 1e-7 * abs(lp.item())
    with pytest.raises(ValueError, match="Use gibbs 1d kernel"):
        SparseMultivariateGibbsKernel(Z, 1, Z.clone())


def test_multivariate_kernel_golden_and_gradients(golden):
    from nonstationary_precip_b200.models.multivariate_gibbs_kernel import MultivariateGibbsKernel
    g = golden("gibbs_full_d2_f64")
    x1, x2 = g["x1"].cuda(), g["x2"].cuda()
    torch.manual_seed(0)
    k = MultivariateGibbsKernel(x1, 2)
    k.H.data = g["H1"].cuda()
    k.D.data = g["Dm"].cuda()
    assert rel(k.forward(x1, x1), g["K11"]) < 1e-12
    k.expectation_conditional_matrix_variate_dist = lambda xs: g["H2"].cuda()
    K12 = k.forward(x1, x2)
    assert rel(K12, g["K12"]) < 1e-12
    K12.sum().backward()
    assert k.H.grad is None  # H is detached in forward, as in the reference
    Dc = g["Dm"].clone().requires_grad_(True)
    o.gibbs_full_K(g["x1"], g["x2"], o.sigma_from_H(g["H1"], Dc), o.sigma_from_H(g["H2"], Dc)).sum().backward()
    assert rel(k.D.grad, Dc.grad) < 1e-9


def test_spatio_temporal_nonstationary_objective_and_predict():
    """SparseSpatioTemporal_Nonstationary (reference models/spatio_temporal_models.py:35-126) vs the oracle: training
    objective, gradients (lengthscale field, spatial Z, temporal hyper-parameters) and both predict variants."""
    from nonstationary_precip_b200.gp_base import ExactMarginalLogLikelihood, GaussianLikelihood
    from nonstationary_precip_b200.models.spatio_temporal_models import SparseSpatioTemporal_Nonstationary
    g = torch.Generator().manual_seed(8)
    n, ns, M = 172, 43, 40
    t = torch.sort(torch.rand(n + ns, generator=g) * 4 - 2)[0]
    xy = torch.rand(n + ns, 2, generator=g) * 2 - 1
    xa = torch.cat([t[:, None], xy], 1)
    perm = torch.randperm(n + ns, generator=g)
    x, xs = xa[perm[:n]], xa[perm[n:]]
    y = torch.sin(2 * math.pi * x[:, 0]) * torch.exp(-x[:, 1] ** 2) + 0.1 * torch.randn(n, generator=g)
    z = x[torch.randperm(n, generator=g)[:M]].clone()
    prior = make_prior(2, active_dims=(0, 1))  # as the reference's experiments/spatio_temporal_exp.py:111
    lik = GaussianLikelihood().cuda().double()
    model = SparseSpatioTemporal_Nonstationary(x.cuda(), y.cuda(), lik, prior, z.cuda(), num_dim=2).cuda().double()
    model.likelihood.noise = 0.05
    model.spatial_covar_module.outputscale = 0.8
    with torch.no_grad():
        model.log_ell_z += 0.2 * torch.randn(2, M, generator=g).cuda()
    tk = model.temporal_covar_module.base_kernel
    mll = ExactMarginalLogLikelihood(lik, model)
    model.train()
    loss = -mll(model(model.train_inputs[0]), model.train_targets)
    loss.backward()

    c, os_, lam = torch.full((2,), math.log(0.3)), torch.ones(2), torch.full((2, 2), 1.3)
    le = model.log_ell_z.detach().cpu().clone().requires_grad_(True)
    zc = z.clone().requires_grad_(True)
    hyp = tk.base_kernel.hyper(tk.outputscale).detach().cpu().clone().requires_grad_(True)
    want = -o.st_sgpr_objective(x, y, zc, le, hyp, torch.tensor(0.8), torch.tensor(0.05), c, os_, lam)
    want.backward()
    assert abs(loss.item() - want.item()) < 1e-8 * abs(want.item())
    assert rel(model.log_ell_z.grad, le.grad) < 1e-5
    # spatial columns of Z are trainable; the temporal alias is frozen (spatio_temporal_models.py:44)
    gz = model.spatial_covar_module.base_kernel.inducing_points.grad
    assert rel(gz[:, 1:], zc.grad[:, 1:]) < 1e-5
    assert model.temporal_covar_module.inducing_points.grad is None
    # chain rule through softplus for the temporal hyper-parameters
    rbf, per = tk.base_kernel.kernels
    assert rel(rbf.raw_lengthscale.grad.reshape(()), hyp.grad[0] * torch.sigmoid(rbf.raw_lengthscale.detach().cpu()).reshape(())) < 1e-6
    assert rel(per.raw_period_length.grad.reshape(()), hyp.grad[2] * torch.sigmoid(per.raw_period_length.detach().cpu()).reshape(())) < 1e-6
    assert rel(tk.raw_outputscale.grad.reshape(()), hyp.grad[3] * torch.sigmoid(tk.raw_outputscale.detach().cpu()).reshape(())) < 1e-6

    model.eval()
    for literal in (True, False):
        with torch.no_grad():
            pred = model.predict(xs.cuda(), literal=literal)
        mu, cov = o.st_predict(x, y, z, le.detach(), xs, hyp.detach(), torch.tensor(0.8), torch.tensor(0.05), c, os_,
                               lam, literal=literal)
        assert rel(pred.mean, mu) < 1e-6, literal
        assert rel(pred.covariance_matrix, cov) < 1e-6, literal
