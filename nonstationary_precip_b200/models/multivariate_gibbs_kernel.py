"""Mirror of the reference's models/multivariate_gibbs_kernel.py:20-150 (MultivariateGibbsKernel) on the npgp kernels.

Paciorek-Schervish kernel with per-point matrices Sigma(x) = softplus((h h^T) o (h h^T)) + D o D built from a latent
N x d matrix H with a matrix-normal prior.  Differences from the reference, all deliberate and documented:
  * everything is fp64 (the reference silently builds Sigma in float32, :98) and d may be 2 or 3 (reference: 2 only);
  * no (N1,N2,d,d) temporaries, no per-row Python loop, no batched LU: one fused tile kernel;
  * the O(N1 N2 d^2) debugging attributes (sigma_matrix_i/j, sig_inv, diff) are not kept.
As in the reference, H enters `forward` detached (:85,:93): the data term sends gradients to D and to the inputs only."""
from __future__ import annotations

import torch

from .. import functional as F
from ..gp_base import Kernel, RBFKernel
from .latent_priors import MatrixVariateNormalPrior

jitter = 1e-5


class _MultivariateGibbsBase(Kernel):
    is_stationary = False

    def _row_hypers(self):
        raise NotImplementedError

    def _anchor(self):
        """(locations, static (K_row + 1e-5 I)^-1 factor P) at which H lives."""
        raise NotImplementedError

    def expectation_conditional_matrix_variate_dist(self, x_star):
        """E[H(x*) | H] = K_row(x*, X) (K_row(X,X) + 1e-5 I)^-1 H  (reference :65-75; the column covariance cancels),
        evaluated matrix free."""
        X, P = self._anchor()
        lam, os = self._row_hypers()
        W = F.spd_solve(P, self.H)
        out = F.rbf_matvec(x_star.contiguous(), X.contiguous(), lam.reshape(1, -1), os, W.unsqueeze(0))
        return out[0]

    def _sigma(self, Hx):
        return F.sigma_from_h(Hx, self.D)

    def forward(self, x1, x2, diag=False, **params):
        nH = self.H.shape[0]
        if torch.equal(x1, x2):
            if x1.shape[0] == nH:
                Hx = self.H.detach()
            else:  # K_** on new inputs (reference :88-94)
                Hx = self.expectation_conditional_matrix_variate_dist(x1).detach()
            S = self._sigma(Hx)
            return F.gibbs_full(x1, S, x1, S, None, jitter)
        if x1.shape[0] == nH:
            Hx1, Hx2 = self.H.detach(), self.expectation_conditional_matrix_variate_dist(x2).detach()
        elif x2.shape[0] == nH:
            Hx2, Hx1 = self.H.detach(), self.expectation_conditional_matrix_variate_dist(x1).detach()
        else:
            # the reference leaves Hx1/Hx2 unbound here (multivariate_gibbs_kernel.py:113-123)
            raise ValueError("one of x1, x2 must have as many rows as the latent matrix H (%d)" % nH)
        return F.gibbs_full(x1, self._sigma(Hx1), x2, self._sigma(Hx2), None, jitter)


class MultivariateGibbsKernel(_MultivariateGibbsBase):
    """MultivariateGibbsKernel(x, input_dim): H lives at the training inputs x (reference :28-63)."""

    def __init__(self, x, input_dim, **kwargs):
        super().__init__(**kwargs)
        if input_dim == 1:
            raise ValueError("Use gibbs 1d kernel for dim 1")
        self.x, self.n, self.d = x, len(x), input_dim
        # RBFKernel(ard_num_dims=d, lengthscale=[0.2, 0.2]): the kwarg is swallowed upstream -> lengthscale softplus(0)
        self.row_covar_kernel = RBFKernel(ard_num_dims=self.d).to(x.device, torch.float64)
        self.row_covar_kernel.requires_grad_(False)
        self.loc = torch.zeros(self.n, self.d, dtype=torch.float64, device=x.device)
        with torch.no_grad():
            self.row_covar = self.row_covar_kernel(self.x)
        self.col_covar = 5.0 * torch.eye(self.d, dtype=torch.float64, device=x.device)
        self.H_matrix_prior = MatrixVariateNormalPrior(self.loc, self.row_covar, self.col_covar)
        self.register_parameter("H", torch.nn.Parameter(self.H_matrix_prior.sample_n(1)))
        self.register_prior("prior_H", self.H_matrix_prior, "H")
        self.register_parameter("D", torch.nn.Parameter(torch.diag(torch.randn(self.d, dtype=torch.float64,
                                                                               device=x.device))))

    def _row_hypers(self):
        return self.row_covar_kernel.lengthscale.reshape(-1).expand(self.d).contiguous(), None

    def _anchor(self):
        return self.x, self.H_matrix_prior._PR
