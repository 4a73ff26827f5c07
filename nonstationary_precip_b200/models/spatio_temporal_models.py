"""Mirror of the reference's models/spatio_temporal_models.py:35-126 (SparseSpatioTemporal_Nonstationary) on the npgp
kernels: a Nystrom (RBF x Periodic) kernel on time (column 0) plus a Nystrom Gibbs kernel on (lon, lat) (columns 1, 2),
both on ONE set of inducing points Z (M, 3); the sum is a rank-2M low-rank root scored by the Woodbury identity.

`predict(x_new)` follows the reference's arithmetic literally by default (`literal=True`): because the summed covariance
is not a LowRankRootLazyTensor, the reference takes the `else` branches (:101-110) and uses ROWS OF THE DENSE COVARIANCE as
the factors `L` and `A^T`.  `literal=False` uses the concatenated low-rank root, i.e. what the comments in the reference
describe (`L = K_*z K_zz^{-1/2}`)."""
from __future__ import annotations

import torch

from .. import functional as F
from ..gp_base import (ExactGP, GreaterThan, InducingPointKernel, LowRankRootCovar, MultivariateNormal, PeriodicKernel,
                       RBFKernel, ScaleKernel, ZeroMean, sum_covariances)
from .gibbs_kernels import GibbsKernel, GibbsSafeScaleKernel, InducingGibbsKernelST


class _ScaledTemporalKernel(ScaleKernel):
    """ScaleKernel(RBF * Periodic, outputscale >= 7): outputscale fused into the temporal tile kernel."""

    def forward(self, x1, x2, diag=False, **params):
        if diag and torch.equal(x1, x2):  # stationary: k(t,t) = outputscale (no n x n matrix for a diagonal)
            return self.outputscale.reshape(1).expand(x1.shape[0])
        return self.base_kernel.forward(x1, x2, outputscale=self.outputscale)


class SpatioTemporal_Stationary(ExactGP):
    """The reference's stationary baseline (models/spatio_temporal_models.py:17-33): wiring only, so that the driver's import
    line resolves (experiments/spatio_temporal_exp.py:20).  Not part of the accelerated path (out of scope)."""

    def __init__(self, train_x, train_y, likelihood, z=None):
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = ZeroMean()
        self.temporal_covar_module = ScaleKernel(RBFKernel(active_dims=(0,)) * PeriodicKernel(active_dims=(0,)),
                                                 outputscale_constraint=GreaterThan(7), active_dims=0)
        self.spatial_covar_module = ScaleKernel(RBFKernel(active_dims=(1, 2)), active_dims=(1, 2))
        base = self.temporal_covar_module + self.spatial_covar_module
        self.covar_module = base if z is None else InducingPointKernel(base_kernel=base, inducing_points=z,
                                                                       likelihood=likelihood)

    def forward(self, x):
        return MultivariateNormal(self.mean_module(x), self.covar_module(x))


class SparseSpatioTemporal_Nonstationary(ExactGP):
    def __init__(self, train_x, train_y, likelihood, prior, z, num_dim=1):
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = ZeroMean()
        self.spatial_covar_module = GibbsSafeScaleKernel(
            InducingGibbsKernelST(GibbsKernel(lengthscale_prior=prior, active_dims=(0, 1)), inducing_points=z,
                                  likelihood=likelihood, active_dims=(1, 2)), active_dims=(1, 2))
        self.temporal_covar_module = InducingPointKernel(
            _ScaledTemporalKernel(RBFKernel(active_dims=(0,)) * PeriodicKernel(active_dims=(0,)),
                                  outputscale_constraint=7.0, active_dims=0),
            inducing_points=self.spatial_covar_module.base_kernel.inducing_points, likelihood=likelihood,
            active_dims=(0,))
        self.temporal_covar_module.inducing_points.requires_grad = False
        z_sp = self.spatial_covar_module.base_kernel._z()
        prior_mean = self.spatial_covar_module.base_kernel.base_kernel.lengthscale_prior.mean_module(z_sp)
        self.register_parameter("log_ell_z", torch.nn.Parameter(prior_mean.detach().clone()))
        # reference :52-55 hands the prior the FULL (M,3) inducing points; the prior's own active_dims=(0,1)
        # (experiments/spatio_temporal_exp.py:111) then select columns 0,1 = (time, lon) for the log-prior term, while the
        # field interpolation sees the (lon, lat) slice (gibbs_kernels.py:310-316).  Kept as the reference has it.
        self.register_prior("ell_z_prior", self.spatial_covar_module.base_kernel.base_kernel.lengthscale_prior,
                            lambda module: (module.spatial_covar_module.base_kernel.inducing_points, module.log_ell_z))

    def _covar(self, x):
        kt = self.temporal_covar_module(x)
        ks = self.spatial_covar_module(x, ell=torch.exp(self.log_ell_z))
        return sum_covariances([kt, ks])

    def forward(self, x, ell=None):
        return MultivariateNormal(self.mean_module(x), self._covar(x))

    def predict(self, x_new, literal=True):
        x = self.train_inputs[0]
        if x_new.dim() == 1:
            x_new = x_new.unsqueeze(-1)
        n = x.shape[-2]
        full_output = self.forward(torch.cat([x, x_new], dim=-2))
        full_covar = full_output.lazy_covariance_matrix
        noise = self.likelihood.noise.reshape(())
        dense = full_covar.evaluate() if isinstance(full_covar, LowRankRootCovar) else full_covar
        test_test = dense[n:, n:]
        if literal:
            L, At = dense[n:, :], dense[:n, :] / torch.sqrt(noise)  # reference :104-110
        else:
            L, At = full_covar.root[n:], full_covar.root[:n] / torch.sqrt(noise)
        k = At.shape[-1]
        eye = torch.eye(k, dtype=x.dtype, device=x.device)
        B = eye + F.matmul(At.T, At)
        _, PB = F.psd_safe_chol_inv(B)
        Binv = F.matmul(PB.T, PB)
        mean = F.matmul(L, F.matmul(Binv, F.matmul(At.T, self.train_targets))) / torch.sqrt(noise) + full_output.loc[n:]
        covar = test_test - F.matmul(L, F.matmul(eye - Binv, L.T))
        return MultivariateNormal(mean, covar)
