"""Mirror of the reference's models/sparse_multivariate_gibbs_kernel.py:20-154 (SparseMultivariateGibbsKernel): the
latent matrix H (M x d) lives at the inducing locations Z; every input whose row count differs from M gets
H(x) = K_row(x, Z) (K_row(Z,Z) + 1e-5 I)^-1 H (reference :67-80), evaluated matrix free.  As in the reference the row
covariance at Z used in that formula is the one computed at construction time (:51) and the matrix-normal prior uses
the static K_row(Z_init) (:53-58).  (The reference module is not importable as shipped: it imports the non-existent
package `kernels`, :11.)"""
from __future__ import annotations

import torch

from .. import functional as F
from ..gp_base import RBFKernel, ScaleKernel
from .latent_priors import MatrixVariateNormalPrior
from .multivariate_gibbs_kernel import _MultivariateGibbsBase

jitter = 1e-5


class SparseMultivariateGibbsKernel(_MultivariateGibbsBase):
    def __init__(self, Z, input_dim, Z_init, **kwargs):
        super().__init__(**kwargs)
        if input_dim == 1:
            raise ValueError("Use gibbs 1d kernel for dim 1")
        self.inducing_locations = Z
        self.d = input_dim
        self.m = Z.shape[0]
        dev = Z.device
        # ScaleKernel(RBFKernel(ard_num_dims=d, lengthscale=[1.3, 1.1])): the lengthscale kwarg is swallowed upstream
        self.row_covar_kernel = ScaleKernel(RBFKernel(ard_num_dims=self.d)).to(dev, torch.float64)
        self.row_covar_kernel.requires_grad_(False)
        self.loc = torch.zeros(self.m, self.d, dtype=torch.float64, device=dev)
        with torch.no_grad():
            self.row_covar = self.row_covar_kernel(Z.detach())
            self.static_row_covar = self.row_covar_kernel(Z_init.detach())
            eye = torch.eye(self.m, dtype=torch.float64, device=dev)
            _, self._P_row = F.psd_safe_chol_inv(self.row_covar + eye * jitter)
        Z_init.requires_grad_(False)
        self.col_covar = torch.eye(self.d, dtype=torch.float64, device=dev)
        self.H_matrix_prior = MatrixVariateNormalPrior(self.loc, self.static_row_covar, self.col_covar)
        self.register_parameter("H", torch.nn.Parameter(self.H_matrix_prior.sample_n(1)))
        self.register_prior("prior_H", self.H_matrix_prior, "H")
        self.register_parameter("D", torch.nn.Parameter(torch.diag(torch.randn(self.d, dtype=torch.float64, device=dev))))

    def _row_hypers(self):
        lam = self.row_covar_kernel.base_kernel.lengthscale.reshape(-1).expand(self.d).contiguous()
        return lam, self.row_covar_kernel.outputscale.reshape(1)

    def _anchor(self):
        return self.inducing_locations.detach(), self._P_row
