"""Matrix-free (streamed) SGPR objective for the diagonal Gibbs kernel: the collapsed bound of DiagonalSparseGP
(reference models/nonstationary_models.py:64-89 driving models/gibbs_kernels.py:187-261)
evaluated WITHOUT materialising the N x M root, so that it runs at the scale of BASELINE config 3 (N = 4M rows,
M = 2048), where the reference's `k_ux1.matmul(inv_root)` would need 68 GB.

With K = Gibbs(X, Z) (unscaled, rows streamed in chunks) and the root rows G = K L^-T (Kzz = L L^T) the bound depends on
the rows only through
    W = G^T G (M x M SYRK),   c = G^T y,   y^T y,
so one pass over the rows accumulates just those (each chunk is whitened by a triangular GEMM first: accumulating the
unwhitened K^T K is cheaper but numerically unusable at N ~ 1e6, see neg_objective_and_grad).  Everything else is M x M
algebra on the blocked Cholesky / GEMM kernels (differentiated by the autograd Functions of `functional.py`).  The
data-side gradient needs a second pass: dObj/dG_chunk = G_chunk (dW + dW^T) + y_chunk dc^T, pushed back through the
whitening (dK = dG L^-1, dL^-1 += dG^T K) and consumed inside the analytic Gibbs backward kernel (never stored as a
gradient matrix); the lengthscale field is interpolated matrix free in both passes.  Rows shard over ranks: W, c, y^T y
and the gradients are sums.  The large products run on the int8 tensor cores (exact Ozaki split, csrc/ozaki.cu) when M
is a multiple of 128, on the FP64 tensor pipe otherwise."""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import functional as F
from . import ops

LOG2PI = math.log(2.0 * math.pi)


def _softplus(x):
    return torch.nn.functional.softplus(x)


def _syrk(Fc, out, use_i8=True):
    """F^T F of a chunk: exact int8 tensor-core path when the width allows (multiple of 128), else FP64 DMMA."""
    if use_i8 and Fc.shape[1] % 128 == 0 and Fc.is_contiguous():
        return ops.syrk_i8(Fc, out=out)
    return ops.wsyrk(Fc, out=out)


def _rowmul(Fc, Csym, T, use_i8=True):
    """T = F_chunk @ Csym (Csym symmetric) on the int8 tensor cores when the width allows (multiple of 64)."""
    if use_i8 and Fc.shape[1] % 64 == 0 and Fc.is_contiguous():
        return ops.rowquad_i8(Fc, Csym, need_q=False, T=T)
    return ops.rowquad(Fc, Csym, need_q=False, T=T)


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
        `all_reduce(t)` sums a tensor over ranks (W, c, y^T y after pass 1; the data-side gradients after pass 2).

        Rows are whitened chunk by chunk before they are accumulated, G = K L^-T (the reference's own root,
        gibbs_kernels.py:225), so that B = I + (s/noise) G^T G has eigenvalues >= 1 and log det Kzz cancels: accumulating
        the unwhitened Gram matrix K^T K instead (12 instead of 22 M^2 flop per row) was measured to move the objective by
        4e-5 at N = 4.2e6 when K^T K changed in its last bits -- `Kzz + (s/noise) K^T K` is too ill-conditioned there."""
        dev = x.device
        n_loc = x.shape[0]
        n = n_total if n_total is not None else n_loc
        M, D = self.Z.shape
        use_i8 = getattr(self, "use_i8", True)
        ell_z, alpha, (_, _, P), lp = self._z_side()
        ell_zd, alphad, Pd = ell_z.detach().contiguous(), alpha.detach().contiguous(), P.detach()

        # ---- pass 1: W = G^T G, c = G^T y, y^T y
        W = torch.zeros(M, M, dtype=torch.float64, device=dev)
        c = torch.zeros(M, dtype=torch.float64, device=dev)
        yy = torch.zeros((), dtype=torch.float64, device=dev)
        rows = min(chunk, n_loc)
        Kbuf = torch.empty(rows, M, dtype=torch.float64, device=dev)
        Gbuf = torch.empty_like(Kbuf)
        Wc = torch.empty_like(W)
        for lo in range(0, n_loc, chunk):
            xc, yc = x[lo:lo + chunk].contiguous(), y[lo:lo + chunk].contiguous()
            _, K = self._chunk_forward(xc, ell_zd, alphad, Kbuf[:xc.shape[0]])
            G = ops.dgemm(K, Pd, transB=True, tri_b=2, C=Gbuf[:xc.shape[0]])
            W += _syrk(G, Wc, use_i8)
            ops.colwsum(G, w=yc, out=c)
            yy += (yc * yc).sum()
        if all_reduce is not None:
            packed = torch.cat([W.reshape(-1), c, yy.reshape(1)])
            all_reduce(packed)
            W, c, yy = packed[:M * M].reshape(M, M), packed[M * M:M * M + M], packed[-1]
        W = W.clone().requires_grad_(True)
        c = c.clone().requires_grad_(True)

        # ---- M x M algebra: B = I + (s/noise) W
        s = _softplus(self.raw_outputscale).reshape(())
        noise = (1e-4 + _softplus(self.raw_noise)).reshape(())
        Bm = torch.eye(M, dtype=torch.float64, device=dev) + (s / noise) * (0.5 * (W + W.T))
        LB, PB = F.psd_safe_chol_inv(Bm)
        w = F.matmul(PB, c)
        quad = yy / noise - (s / (noise * noise)) * (w * w).sum()
        logdet = 2.0 * torch.log(torch.diagonal(LB)).sum() + n * torch.log(noise)
        ll = -0.5 * (quad + logdet + n * LOG2PI)
        trace = -0.5 * (n - torch.diagonal(W).sum()) / noise  # unscaled kernel (gibbs_kernels.py:256-260)
        obj = (ll + trace + lp) / n
        loss = -obj
        dW, dc = torch.autograd.grad(loss, [W, c], retain_graph=True)
        dW2 = (dW + dW.T).contiguous()
        dc = dc.contiguous()

        # ---- pass 2: dG = G (dW + dW^T) + y dc^T;  dK = dG P (to the analytic Gibbs backward);  dP += dG^T K
        d_ell_z = torch.zeros_like(ell_zd)
        dZ = torch.zeros_like(self.Z)
        dalpha = torch.zeros(D, M, 1, dtype=torch.float64, device=dev)
        dP = torch.zeros(M, M, dtype=torch.float64, device=dev)
        Tbuf = torch.empty_like(Kbuf)
        need_dz = self.Z.requires_grad
        for lo in range(0, n_loc, chunk):
            xc, yc = x[lo:lo + chunk].contiguous(), y[lo:lo + chunk].contiguous()
            ell_x, K = self._chunk_forward(xc, ell_zd, alphad, Kbuf[:xc.shape[0]])
            G = ops.dgemm(K, Pd, transB=True, tri_b=2, C=Gbuf[:xc.shape[0]])
            T, _ = _rowmul(G, dW2, Tbuf[:xc.shape[0]], use_i8)
            T.addcmul_(yc.unsqueeze(1), dc.unsqueeze(0))
            ops.dgemm(T, K, transA=True, beta=1.0, C=dP)
            dK = ops.dgemm(T, Pd, tri_b=1, C=Gbuf[:xc.shape[0]])  # G is no longer needed: reuse its storage
            r = ops.gibbs_diag_bwd(xc, ell_x, self.Z.detach(), ell_zd, None, G=dK, need_dx2=need_dz)
            d_ell_z += r["d_ell2"]
            if need_dz:
                dZ += r["d_x2"]
            dlog = (r["d_ell1"] * ell_x).unsqueeze(-1)
            da, dzf = ops.rbf_matvec_bwd(xc, self.Z.detach(), self.prior_lam, self.prior_os, alphad.unsqueeze(-1), dlog,
                                         need_dz=need_dz)
            dalpha += da
            if need_dz:
                dZ += dzf
        if all_reduce is not None:
            parts = [d_ell_z, dZ, dalpha, dP]
            packed = torch.cat([t.reshape(-1) for t in parts])
            all_reduce(packed)
            out, off = [], 0
            for t in parts:
                out.append(packed[off:off + t.numel()].reshape(t.shape))
                off += t.numel()
            d_ell_z, dZ, dalpha, dP = out

        # ---- finish: Z-side graph (Kzz -> P, prior, alpha) + the streamed contributions
        for p_ in self.parameters():
            p_.grad = None
        torch.autograd.backward([loss, alpha, P], [torch.ones_like(loss), dalpha.squeeze(-1), torch.tril(dP)])
        with torch.no_grad():
            self.log_ell_z.grad += d_ell_z * ell_zd
            if need_dz:
                self.Z.grad += dZ
        return loss.detach()


class SGPRSpatioTemporalStream(torch.nn.Module):
    """Matrix-free collapsed bound of SparseSpatioTemporal_Nonstationary (reference models/spatio_temporal_models.py:35-60
    in training mode, scored by ExactMarginalLogLikelihood): Nystrom (outputscale >= 7)(RBF x Periodic) on time (column 0)
    plus scaled Nystrom Gibbs on (lon, lat) (columns 1, 2), both on ONE inducing set Z (M, 3).  The sum of the two
    Nystrom kernels has the rank-2M root [K_t U_t^-1, sqrt(s) K_s U_s^-1], so the rows enter only through
        A = F^T F (2M x 2M),  b = F^T y,  y^T y,        F = [K_t, K_s]  (n x 2M feature rows, streamed in chunks)
    and the N x 2M root (137 GB at N = 4 194 304, M = 2048) is never formed.  As in the reference the temporal kernel's
    inducing points are a frozen alias of Z (:43-44): Z receives gradients from the spatial kernel only.  Temporal
    hyperparameters: lengthscale_rbf, lengthscale_per, period = softplus(raw), outputscale = 7 + softplus(raw)."""

    def __init__(self, Z, log_ell_z, prior_c, prior_os, prior_lam, hyp_t=(1.0, 1.0, 1.0, 7.7), outputscale_s=0.644,
                 noise=0.011, outputscale_t_lower=7.0, learn_inducing_locations=True, include_prior=True):
        super().__init__()
        assert Z.shape[0] % 2 == 0, "M must be even (16-byte aligned column blocks of the n x 2M feature rows)"
        dev = Z.device
        self.os_t_lower = float(outputscale_t_lower)
        self.Z = torch.nn.Parameter(Z.clone(), requires_grad=learn_inducing_locations)
        self.log_ell_z = torch.nn.Parameter(log_ell_z.clone())
        lr, lp, per, os_t = (float(v) for v in hyp_t)
        self.raw_hyp_t = torch.nn.Parameter(torch.tensor(
            [_inv_softplus(lr), _inv_softplus(lp), _inv_softplus(per), _inv_softplus(os_t - self.os_t_lower)],
            dtype=torch.float64, device=dev))
        self.raw_outputscale = torch.nn.Parameter(torch.tensor([_inv_softplus(outputscale_s)], dtype=torch.float64,
                                                               device=dev))
        self.raw_noise = torch.nn.Parameter(torch.tensor([_inv_softplus(noise - 1e-4)], dtype=torch.float64, device=dev))
        self.register_buffer("prior_c", prior_c.clone())
        self.register_buffer("prior_os", prior_os.clone())
        self.register_buffer("prior_lam", prior_lam.clone())
        self.include_prior = include_prior

    def hyp_t(self):
        sp = _softplus(self.raw_hyp_t)
        return torch.cat([sp[:3], self.os_t_lower + sp[3:]])

    @staticmethod
    def _chol_ladder(Kzz, what):
        eye = torch.eye(Kzz.shape[0], dtype=torch.float64, device=Kzz.device)
        for jit in (0.0, 1e-8, 1e-7, 1e-6):  # psd_safe_cholesky ladder, keeping the jittered matrix (reused in Sigma)
            Kj = Kzz if jit == 0.0 else Kzz + jit * eye
            L, P, info = F.chol_inv(Kj)
            if int(info) == 0:
                return Kj, L, P
        raise RuntimeError("%s not positive definite after adding jitter up to 1e-6" % what)

    def _z_side(self):
        M = self.Z.shape[0]
        zs = self.Z[:, 1:3]
        zt = self.Z.detach()[:, 0].contiguous()  # frozen alias (spatio_temporal_models.py:43-44)
        D = 2
        ell_z = torch.exp(self.log_ell_z)
        eye = torch.eye(M, dtype=torch.float64, device=self.Z.device)
        alphas, lp = [], self.Z.new_zeros(())
        z_prior = self.Z[:, 0:2]  # the reference's prior term sees (time, lon): its closure passes the full (M,3) inducing
        # points and the prior's active_dims=(0,1) pick columns 0,1 (spatio_temporal_models.py:52-55, spatio_temporal_exp.py:111)
        for b in range(D):
            Kp = F.rbf_ard(zs, zs, self.prior_lam[b], self.prior_os[b]) + 1e-4 * eye
            Lb, Pb = F.psd_safe_chol_inv(Kp)
            r = self.log_ell_z[b] - self.prior_c[b]
            alphas.append(F.spd_solve(Pb, r))
            if self.include_prior:
                Kq = F.rbf_ard(z_prior, z_prior, self.prior_lam[b], self.prior_os[b]) + 1e-4 * eye
                Lq, Pq = F.psd_safe_chol_inv(Kq)
                lp = lp + (-0.5 * (r * F.spd_solve(Pq, r)).sum() - torch.log(torch.diagonal(Lq)).sum()
                           - 0.5 * M * LOG2PI) / M
        hyp = self.hyp_t()
        spatial = self._chol_ladder(F.gibbs_diag(zs, ell_z, zs, ell_z), "spatial Kzz")
        temporal = self._chol_ladder(ops.rbf_periodic(zt, zt, hyp), "temporal Kzz")
        return ell_z, torch.stack(alphas), hyp, zt, spatial, temporal, lp

    def _chunk_features(self, xc, zs, zt, hypd, ell_z, alpha, Fbuf):
        """Fbuf[:, :M] = K_t (outputscale included), Fbuf[:, M:] = unscaled Gibbs K_s; returns ell(x) of the chunk."""
        M = zs.shape[0]
        xt, xs = xc[:, 0].contiguous(), xc[:, 1:3].contiguous()
        ops.rbfper_fwd(xt, zt, hypd, out=Fbuf[:, :M])
        ell_x = ops.rbf_matvec_fwd(xs, zs, self.prior_lam, self.prior_os, alpha.unsqueeze(-1), bias=self.prior_c,
                                   apply_exp=True).squeeze(-1)
        ops.gibbs_diag_fwd(xs, ell_x, zs, ell_z, out=Fbuf[:, M:])
        return xt, xs, ell_x

    def _whiten(self, Fc, Gc, Pt, Ps):
        """Root rows of the chunk: G = [F_t P_t^T, F_s P_s^T] (the reference's k_ux1.matmul(inv_root),
        gibbs_kernels.py:225), two triangular DMMA GEMMs."""
        M = Pt.shape[0]
        ops.dgemm(Fc[:, :M], Pt, transB=True, tri_b=2, C=Gc[:, :M])
        ops.dgemm(Fc[:, M:], Ps, transB=True, tri_b=2, C=Gc[:, M:])

    def neg_objective_and_grad(self, x, y, chunk: int = 32768, n_total: Optional[int] = None, all_reduce=None):
        """As SGPRGibbsStream.neg_objective_and_grad, for x (n, 3) = (time, lon, lat).

        Numerics: Kzz of the temporal kernel (M inducing times on a smooth 1-D kernel) is nearly singular, and at
        N ~ 4e6 rows the unwhitened Gram matrix F^T F / noise has norm ~1e12, so `Kzz + F^T F / noise` cannot be
        factored in fp64.  Rows are therefore whitened chunk by chunk BEFORE they are accumulated, G = F L^-T, exactly
        as the reference forms its root; then B = I + D G^T G D / noise has eigenvalues >= 1 and the log-determinants of
        Kzz cancel analytically.  Cost per row: 22 M^2 flop for both passes instead of 12 M^2."""
        dev = x.device
        n_loc = x.shape[0]
        n = n_total if n_total is not None else n_loc
        M = self.Z.shape[0]
        M2 = 2 * M
        ell_z, alpha, hyp, zt, (_, _, Ps), (_, _, Pt), lp = self._z_side()
        ell_zd, alphad, hypd = ell_z.detach().contiguous(), alpha.detach().contiguous(), hyp.detach().contiguous()
        Ptd, Psd = Pt.detach(), Ps.detach()
        zs = self.Z.detach()[:, 1:3].contiguous()

        # ---- pass 1: W = G^T G, c = G^T y, y^T y
        W = torch.zeros(M2, M2, dtype=torch.float64, device=dev)
        c = torch.zeros(M2, dtype=torch.float64, device=dev)
        yy = torch.zeros((), dtype=torch.float64, device=dev)
        rows = min(chunk, n_loc)
        Fbuf = torch.empty(rows, M2, dtype=torch.float64, device=dev)
        Gbuf = torch.empty_like(Fbuf)
        Wc = torch.empty_like(W)
        for lo in range(0, n_loc, chunk):
            xc, yc = x[lo:lo + chunk], y[lo:lo + chunk].contiguous()
            Fc, Gc = Fbuf[:xc.shape[0]], Gbuf[:xc.shape[0]]
            self._chunk_features(xc, zs, zt, hypd, ell_zd, alphad, Fc)
            self._whiten(Fc, Gc, Ptd, Psd)
            W += _syrk(Gc, Wc, getattr(self, 'use_i8', True))
            ops.colwsum(Gc, w=yc, out=c)
            yy += (yc * yc).sum()
        if all_reduce is not None:
            packed = torch.cat([W.reshape(-1), c, yy.reshape(1)])
            all_reduce(packed)
            W, c, yy = packed[:M2 * M2].reshape(M2, M2), packed[M2 * M2:M2 * M2 + M2], packed[-1]
        W = W.clone().requires_grad_(True)
        c = c.clone().requires_grad_(True)

        # ---- 2M x 2M algebra: B = I + D W D / noise, D = diag(1, sqrt(s))
        s = _softplus(self.raw_outputscale).reshape(())
        noise = (1e-4 + _softplus(self.raw_noise)).reshape(())
        sv = torch.cat([torch.ones(M, dtype=torch.float64, device=dev), torch.sqrt(s).expand(M)])
        Ws = 0.5 * (W + W.T) * sv[:, None] * sv[None, :]
        cs = c * sv
        Bm = torch.eye(M2, dtype=torch.float64, device=dev) + Ws / noise
        LB, PB = F.psd_safe_chol_inv(Bm)
        w = F.matmul(PB, cs)
        quad = yy / noise - (w * w).sum() / (noise * noise)
        logdet = 2.0 * torch.log(torch.diagonal(LB)).sum() + n * torch.log(noise)
        ll = -0.5 * (quad + logdet + n * LOG2PI)
        dW_ = torch.diagonal(W)
        trace_t = -0.5 * (n * hyp[3] - dW_[:M].sum()) / noise
        trace_s = -0.5 * (n - dW_[M:].sum()) / noise  # unscaled: GibbsSafeScaleKernel wraps the Nystrom kernel
        obj = (ll + trace_t + trace_s + lp) / n
        loss = -obj
        dW, dc = torch.autograd.grad(loss, [W, c], retain_graph=True)
        dW2 = (dW + dW.T).contiguous()
        dc = dc.contiguous()

        # ---- pass 2: dG = G (dW + dW^T) + y dc^T;  dF = dG P (to the analytic kernel backwards);  dP += dG^T F
        d_ell_z = torch.zeros_like(ell_zd)
        dZs = torch.zeros(M, 2, dtype=torch.float64, device=dev)
        dalpha = torch.zeros(2, M, 1, dtype=torch.float64, device=dev)
        dhyp = torch.zeros(4, dtype=torch.float64, device=dev)
        dPt = torch.zeros(M, M, dtype=torch.float64, device=dev)
        dPs = torch.zeros(M, M, dtype=torch.float64, device=dev)
        Tbuf = torch.empty_like(Fbuf)
        need_dz = self.Z.requires_grad
        for lo in range(0, n_loc, chunk):
            xc, yc = x[lo:lo + chunk], y[lo:lo + chunk].contiguous()
            Fc, Gc, Tc = Fbuf[:xc.shape[0]], Gbuf[:xc.shape[0]], Tbuf[:xc.shape[0]]
            xt, xs, ell_x = self._chunk_features(xc, zs, zt, hypd, ell_zd, alphad, Fc)
            self._whiten(Fc, Gc, Ptd, Psd)
            _rowmul(Gc, dW2, Tc, getattr(self, 'use_i8', True))
            Tc.addcmul_(yc.unsqueeze(1), dc.unsqueeze(0))
            ops.dgemm(Tc[:, :M], Fc[:, :M], transA=True, beta=1.0, C=dPt)
            ops.dgemm(Tc[:, M:], Fc[:, M:], transA=True, beta=1.0, C=dPs)
            dFc = Gc  # G is no longer needed: reuse its storage for dF
            ops.dgemm(Tc[:, :M], Ptd, tri_b=1, C=dFc[:, :M])
            ops.dgemm(Tc[:, M:], Psd, tri_b=1, C=dFc[:, M:])
            g4, _ = ops.rbfper_bwd(xt, zt, hypd, dFc[:, :M])
            dhyp += g4
            r = ops.gibbs_diag_bwd(xs, ell_x, zs, ell_zd, None, G=dFc[:, M:], need_dx2=need_dz)
            d_ell_z += r["d_ell2"]
            dlog = (r["d_ell1"] * ell_x).unsqueeze(-1)
            da, dzf = ops.rbf_matvec_bwd(xs, zs, self.prior_lam, self.prior_os, alphad.unsqueeze(-1), dlog, need_dz=need_dz)
            dalpha += da
            if need_dz:
                dZs += r["d_x2"] + dzf
        if all_reduce is not None:
            parts = [d_ell_z, dZs, dalpha, dhyp, dPt, dPs]
            packed = torch.cat([t.reshape(-1) for t in parts])
            all_reduce(packed)
            out, off = [], 0
            for t in parts:
                out.append(packed[off:off + t.numel()].reshape(t.shape))
                off += t.numel()
            d_ell_z, dZs, dalpha, dhyp, dPt, dPs = out

        for p_ in self.parameters():
            p_.grad = None
        torch.autograd.backward([loss, alpha, hyp, Pt, Ps],
                                [torch.ones_like(loss), dalpha.squeeze(-1), dhyp, torch.tril(dPt), torch.tril(dPs)])
        with torch.no_grad():
            self.log_ell_z.grad += d_ell_z * ell_zd
            if need_dz:
                self.Z.grad[:, 1:3] += dZs
        return loss.detach()
