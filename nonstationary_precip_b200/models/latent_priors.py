"""Mirror of the reference's models/latent_priors.py:27-64: the matrix-normal prior over the latent N x D matrix H.

The reference materialises the (N D) x (N D) Kronecker covariance and its inverse; here the density is evaluated through
the two factors (an N x N Cholesky on the npgp kernels and a D x D one), which is what makes it usable beyond N ~ 1e3.
The vec-ordering quirk of the reference is preserved: the covariance is kron(row + 1e-5 I, col) (row-major vec, :45) but
the density is evaluated at x.T.flatten() (column-major vec, :64)."""
from __future__ import annotations

import math

import torch

from .. import functional as F

jitter = 1e-5
LOG2PI = math.log(2.0 * math.pi)


class MatrixVariateNormalPrior:
    def __init__(self, loc, row_covariance_matrix, column_covariance_matrix):
        self.n = row_covariance_matrix.shape[0]
        self.d = column_covariance_matrix.shape[0]
        self.loc = loc
        self.vec_loc = loc.flatten()
        self.row_covariance_matrix = row_covariance_matrix
        self.col_covariance_matrix = column_covariance_matrix
        dev, dt = row_covariance_matrix.device, torch.float64
        R = row_covariance_matrix.to(dt) + torch.eye(self.n, dtype=dt, device=dev) * jitter
        self._LR, self._PR = F.psd_safe_chol_inv(R)
        C = column_covariance_matrix.to(dt)
        self._LC = torch.linalg.cholesky(C.cpu()).to(dev)  # D x D (D = 2 or 3): host-side scalar work
        self._PC = torch.linalg.inv(self._LC.cpu()).to(dev)

    @property
    def kron_cov_inv(self):
        """kron(col^-1, (row + 1e-5 I)^-1) as the reference stores it (:46); dense, only for small N."""
        Rinv = F.matmul(self._PR.T, self._PR)
        Cinv = self._PC.T @ self._PC
        return torch.kron(Cinv, Rinv)

    def sample_n(self, num_samples):
        """One draw reshaped to (N, D) (reference :59-61; meaningful for num_samples == 1, as the reference uses it)."""
        dev = self._LR.device
        eps = torch.randn(self.n, self.d, dtype=torch.float64, device=dev)
        # vec_r(X) ~ N(0, R (x) C)  <=>  X = L_R E L_C^T
        return self.loc.to(torch.float64) + F.matmul(self._LR, eps) @ self._LC.T

    def log_prob(self, x):
        """log N(x.T.flatten(); vec(loc), kron(R, C)) evaluated through the factors.

        With v = x.T.flatten() read as a row-major (N, D) matrix V (that is the reference's ordering mismatch),
        quad = tr(C^-1 V^T R^-1 V) and log det = D log det R + N log det C."""
        n, d = self.n, self.d
        V = (x.to(torch.float64).T.flatten() - self.vec_loc.to(torch.float64)).reshape(n, d)
        W = F.matmul(self._PR, V) @ self._PC.T  # L_R^-1 V L_C^-T
        logdet = 2.0 * d * torch.log(torch.diagonal(self._LR)).sum() + 2.0 * n * torch.log(torch.diagonal(self._LC)).sum()
        return -0.5 * (W * W).sum() - 0.5 * logdet - 0.5 * n * d * LOG2PI
