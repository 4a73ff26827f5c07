"""Autograd-enabled building blocks over the npgp C ABI, used by the mirrored reference classes (``models/``).

Everything here runs on the hand-written CUDA kernels: kernel matrices (``gibbs_diag``, ``gibbs_full``, ``rbf_ard``),
matrix-free field interpolation (``rbf_matvec``), FP64 tensor-core products (``matmul``), the blocked Cholesky with
inverse factor (``chol_inv``) and what is built from them (``spd_solve``, ``mvn_log_prob``).  Backward passes are the
analytic kernels of ``ops`` or compositions of the same GEMM / Cholesky kernels -- no library fallback.

Also holds the four helpers of the reference's ``utils/functional.py`` that the hot path uses (``op``, ``dot``,
``mv``, ``t``; utils/functional.py:14-33,60-64)."""
from __future__ import annotations

import math

import torch

from . import ops
from .ops import gibbs_diag, gibbs_full, rbf_matvec, sigma_from_h, sym_pack, sym_unpack  # noqa: F401  (re-exported)

LOG2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------------------------------------------------
# dense products on the FP64 tensor pipe
# ----------------------------------------------------------------------------------------------------------------------
def _aligned(t: torch.Tensor) -> torch.Tensor:
    """2-D fp64 view usable by npgp_dgemm: unit inner stride, even leading dimension, 16-byte aligned."""
    if t.stride(-1) == 1 and t.stride(0) % 2 == 0 and t.data_ptr() % 16 == 0 and t.stride(0) >= t.shape[1]:
        return t
    r, c = t.shape
    buf = torch.empty(r, c + (c & 1), dtype=torch.float64, device=t.device)
    buf[:, :c] = t
    if c & 1:
        buf[:, c] = 0.0
    return buf[:, :c]


def _mm(A, B, transA=False, transB=False, alpha=1.0):
    return ops.dgemm(_aligned(A), _aligned(B), transA, transB, alpha=alpha)


class MatmulFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, A, B):
        ctx.save_for_backward(A, B)
        return _mm(A.detach(), B.detach())

    @staticmethod
    def backward(ctx, G):
        A, B = ctx.saved_tensors
        dA = _mm(G, B, transB=True) if ctx.needs_input_grad[0] else None
        dB = _mm(A, G, transA=True) if ctx.needs_input_grad[1] else None
        return dA, dB


def matmul(A: torch.Tensor, B: torch.Tensor) -> torch.Tensor:
    """A @ B for 2-D fp64 CUDA tensors (vectors are promoted), differentiable."""
    va, vb = A.dim() == 1, B.dim() == 1
    A2 = A.unsqueeze(0) if va else A
    B2 = B.unsqueeze(1) if vb else B
    out = MatmulFn.apply(A2, B2)
    if va:
        out = out.squeeze(0)
    if vb:
        out = out.squeeze(-1)
    return out


# ----------------------------------------------------------------------------------------------------------------------
# Cholesky with inverse factor
# ----------------------------------------------------------------------------------------------------------------------
class CholInvFn(torch.autograd.Function):
    """A (SPD) -> (L, P = L^-1).  Backward: dA = sym(P^T Phi(L^T dL - dP P^T) P)   (Murray 2016 + d(L^-1))."""

    @staticmethod
    def forward(ctx, A):
        Ac = A.detach().clone().contiguous()
        if Ac.shape[0] & 1:  # odd size: embed in an even leading dimension
            buf = torch.zeros(Ac.shape[0], Ac.shape[0] + 1, dtype=torch.float64, device=A.device)
            buf[:, :-1] = Ac
            Ac = buf[:, :-1]
            L, P, info = _potrf_inv_view(Ac)
        else:
            L, P, info = ops.potrf_inv(Ac, overwrite=True)
        ctx.save_for_backward(L, P)
        ctx.info = info
        ctx.mark_non_differentiable(info)
        return L, P, info

    @staticmethod
    def backward(ctx, dL, dP, _):
        L, P = ctx.saved_tensors
        n = L.shape[0]
        X = torch.zeros(n, n, dtype=torch.float64, device=L.device)
        if dL is not None:
            X = X + _mm(L, dL, transA=True)
        if dP is not None:
            X = X - _mm(dP, P, transB=True)
        X = X.contiguous()
        ops.phi_mask_(X, 1.0)
        dA = _mm(_mm(P, X, transA=True), P)
        return 0.5 * (dA + dA.T)


def _potrf_inv_view(Av):
    """potrf_inv on a strided (even-ld) view."""
    from ._lib import check, lib, ptr, stream
    M = Av.shape[0]
    Pbuf = torch.empty(M, Av.stride(0), dtype=torch.float64, device=Av.device)
    P = Pbuf[:, :M]
    nbytes = lib().npgp_potrf_workspace_bytes(M)
    work = torch.empty(nbytes // 8 + 1, dtype=torch.float64, device=Av.device)
    info = torch.zeros((), dtype=torch.int32, device=Av.device)
    check(lib().npgp_potrf_inv_lower(M, ptr(Av), Av.stride(0), ptr(P), P.stride(0), ptr(work), nbytes, ptr(info),
                                     stream()), "npgp_potrf_inv_lower")
    return Av, P, info


def chol_inv(A: torch.Tensor):
    """(L, P = L^-1, info) of a symmetric positive definite matrix; differentiable in L and P."""
    return CholInvFn.apply(A)


def psd_safe_chol_inv(A: torch.Tensor, max_tries: int = 3):
    """gpytorch.utils.cholesky.psd_safe_cholesky semantics (reference models/gibbs_kernels.py:201): plain attempt, then
    jitter 1e-8 * 10^i (fp64) on the diagonal, up to `max_tries` times.  Reads `info` back (one host sync per attempt, as
    the reference's torch.linalg.cholesky_ex does)."""
    L, P, info = chol_inv(A)
    if int(info) == 0:
        return L, P
    eye = torch.eye(A.shape[0], dtype=A.dtype, device=A.device)
    for i in range(max_tries):
        L, P, info = chol_inv(A + (1e-8 * 10 ** i) * eye)
        if int(info) == 0:
            return L, P
    raise RuntimeError("matrix not positive definite after adding jitter up to %.1e" % (1e-8 * 10 ** (max_tries - 1)))


def spd_solve(P: torch.Tensor, rhs: torch.Tensor) -> torch.Tensor:
    """K^-1 rhs given P = chol(K)^-1."""
    return matmul(P.T, matmul(P, rhs))


def mvn_log_prob(y, mean, cov, jitter_ladder: bool = True):
    """log N(y | mean, cov) by Cholesky (the dense path GPyTorch takes for n <= 800)."""
    n = y.shape[0]
    L, P = psd_safe_chol_inv(cov) if jitter_ladder else chol_inv(cov)[:2]
    w = matmul(P, y - mean)
    return -0.5 * (w * w).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * n * LOG2PI


# ----------------------------------------------------------------------------------------------------------------------
# kernel matrices
# ----------------------------------------------------------------------------------------------------------------------
def rbf_ard(x1, x2, lengthscale, outputscale=None):
    """GPyTorch RBFKernel(ard) [* ScaleKernel]: os * exp(-0.5 |(x - x')/l|^2), as a diagonal-Gibbs matrix with constant
    lengthscales (prod sqrt(2 l l/(l^2+l^2)) = 1, exponent sum d^2/(2 l^2)); gradients to x, l, os through the analytic
    Gibbs backward kernel."""
    d = x1.shape[1]
    lam = lengthscale.reshape(-1)
    if lam.numel() == 1:
        lam = lam.expand(d)
    e1 = lam.reshape(d, 1).expand(d, x1.shape[0])
    e2 = lam.reshape(d, 1).expand(d, x2.shape[0])
    scale = None if outputscale is None else outputscale.reshape(1)
    return gibbs_diag(x1, e1, x2, e2, scale)


# ----------------------------------------------------------------------------------------------------------------------
# reference utils/functional.py helpers used on the hot path
# ----------------------------------------------------------------------------------------------------------------------
def dot(v1, v2):
    """Batch dot product (utils/functional.py:14-16)."""
    return (v1 * v2).sum(-1)


def t(x):
    """Matrix transpose (utils/functional.py:19-21)."""
    return torch.transpose(x, -1, -2)


def op(v1, v2=None):
    """Vector outer product (utils/functional.py:60-64) -- broadcast multiply instead of a K=1 batched GEMM."""
    if v2 is None:
        v2 = v1
    return v1.unsqueeze(-1) * v2.unsqueeze(-2)


def mv(matrix, vector, invert=False):
    """Matrix-vector product, or solve when invert=True (utils/functional.py:29-33; the reference uses an LU solve, here
    the matrix is SPD on every hot-path call site and the blocked Cholesky is used)."""
    if not invert:
        return matmul(matrix, vector)
    _, P = psd_safe_chol_inv(matrix)
    return spd_solve(P, vector)
