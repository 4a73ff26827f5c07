"""Mirror of the reference's models/nonstationary_models.py (GP models using Gibbs kernels) on the npgp kernels.

  DiagonalExactGP(train_x, train_y, likelihood, prior, num_dim=1)      reference models/nonstationary_models.py:22-62
  DiagonalSparseGP(train_x, train_y, likelihood, prior, z, num_dim=1)  reference models/nonstationary_models.py:64-153

`forward(x)` returns the prior MultivariateNormal at the training inputs (to be scored by ExactMarginalLogLikelihood),
`predict(x_new)` the predictive distribution, exactly as in the reference."""
from __future__ import annotations

import torch

from .. import functional as F
from ..gp_base import ExactGP, LowRankRootCovar, MultivariateNormal, ZeroMean
from .gibbs_kernels import GibbsKernel, GibbsSafeScaleKernel, InducingGibbsKernel


class DiagonalExactGP(ExactGP):
    """MAP inference of a diagonal-Gibbs GP over the per-training-point log-lengthscales `log_ell_train_x` (D, n)."""

    def __init__(self, train_x, train_y, likelihood, prior, num_dim=1):
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = ZeroMean()
        self.covar_module = GibbsSafeScaleKernel(GibbsKernel(lengthscale_prior=prior, ard_num_dims=num_dim))
        prior_mean = self.covar_module.base_kernel.lengthscale_prior.mean_module(self.train_inputs[0])
        self.register_parameter("log_ell_train_x", torch.nn.Parameter(prior_mean.detach().clone()))
        self.register_prior("ell_train_prior", self.covar_module.base_kernel.lengthscale_prior,
                            lambda module: (module.train_inputs[0], module.log_ell_train_x))

    def forward(self, x):
        mean = self.mean_module(x)
        covar = self.covar_module(x, ell1=torch.exp(self.log_ell_train_x))
        return MultivariateNormal(mean, covar)

    def predict(self, x_new):
        """Predictive at x_new given the lengthscales at the training points (reference :45-62): full covariance
        K_ss - K_sx (K_xx + noise I)^-1 K_xs + 1e-4 I."""
        x = self.train_inputs[0]
        ell = torch.exp(self.log_ell_train_x)
        noise = self.likelihood.noise.reshape(())
        K_xx = self.covar_module(x, ell1=ell)
        ell2 = self.covar_module.base_kernel.lengthscale_prior.conditional_sample(x_new, given=(x, ell))
        K_ss = self.covar_module(x_new, ell1=ell2)
        K_sx = self.covar_module(x_new, x, ell1=ell2, ell2=ell)
        n = K_xx.shape[-1]
        _, P = F.psd_safe_chol_inv(K_xx + noise * torch.eye(n, dtype=x.dtype, device=x.device))
        A = F.matmul(K_sx, P.T)  # K_sx L^-T
        mu = F.matmul(A, F.matmul(P, self.train_targets))
        sigma = K_ss - F.matmul(A, A.T)
        ns = K_ss.shape[-1]
        return MultivariateNormal(mu, sigma + 1e-4 * torch.eye(ns, dtype=x.dtype, device=x.device))


class DiagonalSparseGP(ExactGP):
    """MAP inference of the sparse (SGPR) Gibbs GP over the log-lengthscales at the inducing points `log_ell_z` (D, M)."""

    def __init__(self, train_x, train_y, likelihood, prior, z, num_dim=1):
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = ZeroMean()
        self.covar_module = GibbsSafeScaleKernel(
            InducingGibbsKernel(GibbsKernel(lengthscale_prior=prior, ard_num_dims=num_dim), z, likelihood))
        zz = z.unsqueeze(-1) if z.dim() == 1 else z
        prior_mean = self.covar_module.base_kernel.base_kernel.lengthscale_prior.mean_module(zz)
        self.register_parameter("log_ell_z", torch.nn.Parameter(prior_mean.detach().clone()))
        self.register_prior("ell_z_prior", self.covar_module.base_kernel.base_kernel.lengthscale_prior,
                            lambda module: (module.covar_module.base_kernel.inducing_points, module.log_ell_z))

    def forward(self, x, ell=None):
        mean = self.mean_module(x)
        covar = self.covar_module(x, ell=torch.exp(self.log_ell_z))
        return MultivariateNormal(mean, covar)

    def predict(self, x_new):
        """Predictive at x_new (reference :91-153): joint low-rank root over [train; test], B = I + A A^T,
        mean = L B^-1 A y / sigma, covariance = K_** - L (I - B^-1) L^T.  As the reference warns, only the marginals
        are meaningful."""
        x = self.train_inputs[0]
        if x_new.dim() == 1:
            x_new = x_new.unsqueeze(-1)
        n = x.shape[-2]
        full_output = self.forward(torch.cat([x, x_new], dim=-2))
        full_covar = full_output.lazy_covariance_matrix
        assert isinstance(full_covar, LowRankRootCovar)
        root = full_covar.root
        noise = self.likelihood.noise.reshape(())
        L = root[n:]
        At = root[:n] / torch.sqrt(noise)
        M = root.shape[-1]
        eye = torch.eye(M, dtype=x.dtype, device=x.device)
        B = eye + F.matmul(At.T, At)
        LB, PB = F.psd_safe_chol_inv(B)
        Binv = F.matmul(PB.T, PB)
        mean = F.matmul(L, F.matmul(Binv, F.matmul(At.T, self.train_targets))) / torch.sqrt(noise) \
            + full_output.loc[n:]
        test_test = F.matmul(L, L.T)
        if full_covar.added_diag is not None:
            test_test = test_test + torch.diag(full_covar.added_diag[n:])
        covar = test_test - F.matmul(L, F.matmul(eye - Binv, L.T))
        return MultivariateNormal(mean, covar)
