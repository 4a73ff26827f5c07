"""Matrix-free (streamed) SGPR objective for the diagonal Gibbs kernel: the collapsed bound of DiagonalSparseGP
(reference models/nonstationary_models.py:64-89 driving models/gibbs_kernels.py:187-261)
evaluated WITHOUT materialising the N x M root, so that it runs at the scale of BASELINE config 3 (N = 4M rows,
M = 2048), where the reference's `k_ux1.matmul(inv_root)` would need 68 GB.

With K = Gibbs(X, Z) (unscaled, rows streamed in chunks) the bound depends on the rows only through
    A = K^T K (M x M, DMMA SYRK),   b = K^T y,   y^T y
(the root R = K U^-1 of the reference satisfies R^T R = U^-T A U^-1), so one pass over the rows accumulates just those.  Everything else is
M x M algebra on the blocked Cholesky / GEMM kernels (differentiated by the autograd Functions of `functional.py`).
The data-side gradient needs a second pass:  dObj/dK_chunk = K_chunk (dA + dA^T) + y_chunk db^T, formed by one DMMA GEMM
per chunk and consumed inside the analytic Gibbs backward kernel (never stored as a gradient matrix); the lengthscale
field is interpolated matrix free in both passes.  Rows shard over ranks: A, b, y^T y and the gradients are sums."""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import functional as F
from . import ops

LOG2PI = math.log(2.0 * math.pi)


def _softplus(x):
    return torch.nn.functional.softplus(x)


def _inv_softplus(v: float) -> float:
    return v + math.log(-math.expm1(-v))


class SGPRGibbsStream(torch.nn.Module):
    def __init__(self, Z, log_ell_z, prior_c, prior_os, prior_lam, outputscale=0.644, noise=0.011,
                 learn_inducing_locations=True, include_prior=True, jitter_zz=0.0):
        super().__init__()
        self.jitter_zz = jitter_zz
        self.Z = torch.nn.Parameter(Z.clone(), requires_grad=learn_inducing_locations)
        self.log_ell_z = torch.nn.Parameter(log_ell_z.clone())
        self.raw_outputscale = torch.nn.Parameter(torch.tensor([_inv_softplus(outputscale)], dtype=torch.float64,
                                                               device=Z.device))
        self.raw_noise = torch.nn.Parameter(torch.tensor([_inv_softplus(noise - 1e-4)], dtype=torch.float64,
                                                         device=Z.device))
        self.register_buffer("prior_c", prior_c.clone())
        self.register_buffer("prior_os", prior_os.clone())
        self.register_buffer("prior_lam", prior_lam.clone())
        self.include_prior = include_prior

    # ------------------------------------------------------------------------------------------------------------------
    def _z_side(self):
        """Autograd graph of everything that depends on Z-side parameters only."""
        M, D = self.Z.shape
        ell_z = torch.exp(self.log_ell_z)
        eye = torch.eye(M, dtype=torch.float64, device=self.Z.device)
        alphas, lp = [], self.Z.new_zeros(())
        for b in range(D):
            Kp = F.rbf_ard(self.Z, self.Z, self.prior_lam[b], self.prior_os[b]) + 1e-4 * eye
            Lb, Pb = F.psd_safe_chol_inv(Kp)
            r = self.log_ell_z[b] - self.prior_c[b]
            a = F.spd_solve(Pb, r)
            alphas.append(a)
            if self.include_prior:  # LogNormalPriorProcess.log_prob (gibbs_kernels.py:102-109): per dim, divided by M
                lp = lp + (-0.5 * (r * a).sum() - torch.log(torch.diagonal(Lb)).sum() - 0.5 * M * LOG2PI) / M
        alpha = torch.stack(alphas)
        Kzz = F.gibbs_diag(self.Z, ell_z, self.Z, ell_z)
        # psd_safe_cholesky ladder (gibbs_kernels.py:201), keeping the jittered matrix: it is reused in Sigma below
        for jit in (0.0, 1e-8, 1e-7, 1e-6):
            Kj = Kzz + (self.jitter_zz + jit) * eye
            L, P, info = F.chol_inv(Kj)
            if int(info) == 0:
                break
        else:
            raise RuntimeError("Kzz not positive definite after adding jitter up to 1e-6")
        return ell_z, alpha, (Kj, L, P), lp

    def _chunk_forward(self, xc, ell_z, alpha, K_out=None):
        ell_x = ops.rbf_matvec_fwd(xc, self.Z.detach(), self.prior_lam, self.prior_os, alpha.unsqueeze(-1),
                                   bias=self.prior_c, apply_exp=True).squeeze(-1)
        K = ops.gibbs_diag_fwd(xc, ell_x, self.Z.detach(), ell_z, out=K_out)
        return ell_x, K

    def neg_objective_and_grad(self, x, y, chunk: int = 65536, n_total: Optional[int] = None, all_reduce=None,
                               world_size: int = 1):
        """Fills .grad of the parameters with the gradient of MINUS the collapsed SGPR objective (divided by n, as
        ExactMarginalLogLikelihood does) and returns its value.  `x`, `y`: this rank's rows; `n_total`: global row count;
        `all_reduce(t)` sums a tensor over ranks (A, b, y^T y after pass 1; the flat gradient after pass 2)."""
        dev = x.device
        n_loc = x.shape[0]
        n = n_total if n_total is not None else n_loc
        M, D = self.Z.shape
        ell_z, alpha, (Kzz, L, P), lp = self._z_side()
        ell_zd, alphad = ell_z.detach().contiguous(), alpha.detach().contiguous()

        # ---- pass 1: A = K^T K, b = K^T y, y^T y
        A = torch.zeros(M, M, dtype=torch.float64, device=dev)
        b = torch.zeros(M, dtype=torch.float64, device=dev)
        yy = torch.zeros((), dtype=torch.float64, device=dev)
        Kbuf = torch.empty(min(chunk, n_loc), M, dtype=torch.float64, device=dev)
        Ac = torch.empty_like(A)
        for lo in range(0, n_loc, chunk):
            xc, yc = x[lo:lo + chunk].contiguous(), y[lo:lo + chunk].contiguous()
            _, K = self._chunk_forward(xc, ell_zd, alphad, Kbuf[:xc.shape[0]])
            A += ops.wsyrk(K, out=Ac)
            ops.colwsum(K, w=yc, out=b)
            yy += (yc * yc).sum()
        if all_reduce is not None:
            packed = torch.cat([A.reshape(-1), b, yy.reshape(1)])
            all_reduce(packed)
            A, b, yy = packed[:M * M].reshape(M, M), packed[M * M:M * M + M], packed[-1]
        A = A.clone().requires_grad_(True)
        b = b.clone().requires_grad_(True)

        # ---- M x M algebra (autograd over the Cholesky / GEMM kernels)
        s = _softplus(self.raw_outputscale).reshape(())
        noise = (1e-4 + _softplus(self.raw_noise)).reshape(())
        # Numerically stable form: with Sigma = Kzz + (s/noise) A (positive definite by construction),
        #   log det(s Q + noise I) = n log noise + log det Sigma - log det Kzz,
        #   y^T (s Q + noise I)^-1 y = y^T y / noise - (s / noise^2) b^T Sigma^-1 b,     Q = K Kzz^-1 K^T.
        # (Forming I + (s/noise) P A P^T instead loses positive definiteness for ill-conditioned Kzz.)
        Sigma = Kzz + (s / noise) * (0.5 * (A + A.T))
        LS, PS = F.psd_safe_chol_inv(Sigma)
        w = F.matmul(PS, b)
        quad = yy / noise - (s / (noise * noise)) * (w * w).sum()
        logdet = 2.0 * (torch.log(torch.diagonal(LS)).sum() - torch.log(torch.diagonal(L)).sum()) + n * torch.log(noise)
        ll = -0.5 * (quad + logdet + n * LOG2PI)
        trace = -0.5 * (n - (F.matmul(P, A) * P).sum()) / noise  # tr(Kzz^-1 A) = tr(P A P^T)
        obj = (ll + trace + lp) / n
        loss = -obj
        dA, db = torch.autograd.grad(loss, [A, b], retain_graph=True)
        dA2 = (dA + dA.T).contiguous()  # d(loss)/dK_chunk = K_chunk (dA + dA^T) + y db^T
        db = db.contiguous()

        # ---- pass 2: data-side gradients, streamed
        d_ell_z = torch.zeros_like(ell_zd)
        dZ = torch.zeros_like(self.Z)
        dalpha = torch.zeros(D, M, 1, dtype=torch.float64, device=dev)
        Tbuf = torch.empty_like(Kbuf)
        for lo in range(0, n_loc, chunk):
            xc, yc = x[lo:lo + chunk].contiguous(), y[lo:lo + chunk].contiguous()
            ell_x, K = self._chunk_forward(xc, ell_zd, alphad, Kbuf[:xc.shape[0]])
            T, _ = ops.rowquad(K, dA2, need_q=False, T=Tbuf[:xc.shape[0]])
            r = ops.gibbs_diag_bwd(xc, ell_x, self.Z.detach(), ell_zd, None, G=T, rowvec=yc, colvec=db,
                                   need_dx2=self.Z.requires_grad)
            d_ell_z += r["d_ell2"]
            if self.Z.requires_grad:
                dZ += r["d_x2"]
            dlog = (r["d_ell1"] * ell_x).unsqueeze(-1)
            da, dzf = ops.rbf_matvec_bwd(xc, self.Z.detach(), self.prior_lam, self.prior_os, alphad.unsqueeze(-1), dlog,
                                         need_dz=self.Z.requires_grad)
            dalpha += da
            if self.Z.requires_grad:
                dZ += dzf
        if all_reduce is not None:
            packed = torch.cat([d_ell_z.reshape(-1), dZ.reshape(-1), dalpha.reshape(-1)])
            all_reduce(packed)
            k1, k2 = d_ell_z.numel(), dZ.numel()
            d_ell_z, dZ, dalpha = (packed[:k1].reshape(d_ell_z.shape), packed[k1:k1 + k2].reshape(dZ.shape),
                                   packed[k1 + k2:].reshape(dalpha.shape))

        # ---- finish: Z-side graph (Kzz, prior, alpha) + the streamed contributions
        for p_ in self.parameters():
            p_.grad = None
        torch.autograd.backward([loss, alpha], [torch.ones_like(loss), dalpha.squeeze(-1)])
        with torch.no_grad():
            self.log_ell_z.grad += d_ell_z * ell_zd
            if self.Z.requires_grad:
                self.Z.grad += dZ
        return loss.detach()
