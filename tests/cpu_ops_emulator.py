"""TEST INFRASTRUCTURE: a CPU stand-in for ``nonstationary_precip_b200.ops`` built from the oracle + torch autograd.

It lets the CPU test tier run the *orchestration* code of the product (svgp.py, models) -- the analytic gradient
chain, the flat-buffer layout, the multi-rank weighting -- without a GPU, by injecting this module as ``ops=``.  It is
never imported by the product; on a GPU box the same code runs on the CUDA kernels and is checked by the -m gpu tests."""
import torch

from oracle import gibbs_oracle as o


def sym_pack(S):
    d = S.shape[-1]
    iu = torch.triu_indices(d, d)
    return S[..., iu[0], iu[1]].contiguous()


def sym_unpack(Sp, d):
    iu = torch.triu_indices(d, d)
    S = Sp.new_zeros(*Sp.shape[:-1], d, d)
    S[..., iu[0], iu[1]] = Sp
    S[..., iu[1], iu[0]] = Sp
    return S


def _d_from_P(P):
    return {1: 1, 3: 2, 6: 3}[P]


def _G(n1, n2, G, rowscale, rowvec, colvec):
    out = torch.zeros(n1, n2, dtype=torch.float64)
    if G is not None:
        out = out + (G if rowscale is None else rowscale[:, None] * G)
    if rowvec is not None:
        out = out + rowvec[:, None] * colvec[None, :]
    return out


def gibbs_diag_fwd(x1, ell1, x2, ell2, scale=None, u=None, out=None):
    K = o.gibbs_diag_K(x1, x2, ell1, ell2)
    if scale is not None:
        K = K * scale
    if out is not None:
        out.copy_(K)
        K = out
    return K if u is None else (K, K @ u)


def gibbs_diag_bwd(x1, ell1, x2, ell2, scale=None, G=None, rowscale=None, rowvec=None, colvec=None, need_dx1=False,
                   need_dx2=False, need_dscale=False):
    x1, ell1, x2, ell2 = [t.detach().clone().requires_grad_(True) for t in (x1, ell1, x2, ell2)]
    s = (scale.detach().clone() if scale is not None else torch.ones(1, dtype=torch.float64)).requires_grad_(True)
    K = s * o.gibbs_diag_K(x1, x2, ell1, ell2)
    Gf = _G(x1.shape[0], x2.shape[0], G, rowscale, rowvec, colvec)
    gx1, ge1, gx2, ge2, gs = torch.autograd.grad((K * Gf).sum(), (x1, ell1, x2, ell2, s))
    return dict(d_ell1=ge1, d_ell2=ge2, d_x1=gx1 if need_dx1 else None, d_x2=gx2 if need_dx2 else None,
                d_scale=gs.sum() if need_dscale else None)


def gibbs_full_fwd(x1, S1p, x2, S2p, jitter=1e-5, scale=None, u=None, out=None):
    d = x1.shape[1]
    K = o.gibbs_full_K(x1, x2, sym_unpack(S1p, d), sym_unpack(S2p, d), jitter)
    if scale is not None:
        K = K * scale
    if out is not None:
        out.copy_(K)
        K = out
    return K if u is None else (K, K @ u)


def gibbs_full_bwd(x1, S1p, x2, S2p, jitter=1e-5, scale=None, G=None, rowscale=None, rowvec=None, colvec=None,
                   need_dx1=False, need_dx2=False, need_dscale=False):
    d = x1.shape[1]
    x1, x2 = [t.detach().clone().requires_grad_(True) for t in (x1, x2)]
    S1 = sym_unpack(S1p.detach(), d).requires_grad_(True)
    S2 = sym_unpack(S2p.detach(), d).requires_grad_(True)
    s = (scale.detach().clone() if scale is not None else torch.ones(1, dtype=torch.float64)).requires_grad_(True)
    K = s * o.gibbs_full_K(x1, x2, S1, S2, jitter)
    Gf = _G(x1.shape[0], x2.shape[0], G, rowscale, rowvec, colvec)
    gx1, gS1, gx2, gS2, gs = torch.autograd.grad((K * Gf).sum(), (x1, S1, x2, S2, s))
    sym = lambda g: sym_pack(0.5 * (g + g.transpose(-1, -2)))
    return dict(d_S1=sym(gS1), d_S2=sym(gS2), d_x1=gx1 if need_dx1 else None, d_x2=gx2 if need_dx2 else None,
                d_scale=gs.sum() if need_dscale else None)


def sigma_from_h_fwd(H, Dm):
    return sym_pack(o.sigma_from_H(H, Dm))


def sigma_from_h_bwd(H, Dm, dS, need_dD=True):
    d = H.shape[1]
    H, Dm = H.detach().clone().requires_grad_(True), Dm.detach().clone().requires_grad_(True)
    S = o.sigma_from_H(H, Dm)
    gH, gD = torch.autograd.grad((S * sym_unpack(dS, d)).sum(), (H, Dm))
    return gH, (gD if need_dD else None)


def _rbf(x, z, lam, os):
    K = o.rbf_ard_K(x, z, lam)  # (nb,n,m)
    if os is not None:
        K = K * os.reshape(-1, 1, 1)
    return K


def rbf_matvec_fwd(x, z, lam, os, V, bias=None, apply_exp=False):
    out = _rbf(x, z, lam, os) @ V
    if bias is not None:
        out = out + bias.reshape(-1, 1, 1)
    return torch.exp(out) if apply_exp else out


def rbf_matvec_bwd(x, z, lam, os, V, dOut, need_dz=True):
    z, V = z.detach().clone().requires_grad_(True), V.detach().clone().requires_grad_(True)
    out = _rbf(x, z, lam, os) @ V
    gz, gV = torch.autograd.grad((out * dOut).sum(), (z, V))
    return gV, (gz if need_dz else None)


def dgemm(A, B, transA=False, transB=False, alpha=1.0, beta=0.0, C=None, tri_a=0, tri_b=0, out_tri=0):
    opA, opB = (A.T if transA else A), (B.T if transB else B)
    for t, op in ((tri_a, opA), (tri_b, opB)):  # the structure hints must be true
        if t == 1:
            assert torch.triu(op, 1).abs().max() == 0
        if t == 2:
            assert torch.tril(op, -1).abs().max() == 0
    R = alpha * (opA @ opB)
    if C is not None:
        R = R + beta * C
        C.copy_(R)
        return C
    return R


def rowquad(K, Cm, need_q=True, T=None):
    Tn = K @ Cm
    if T is not None:
        T.copy_(Tn)
        Tn = T
    return Tn, ((Tn * K).sum(-1) if need_q else None)


def wsyrk(K, w=None, alpha=1.0, out=None, uniform_count=None, uniform_target=0.0):
    R = alpha * (K.T @ (K if w is None else w[:, None] * K))
    if out is not None:
        out.copy_(R)
        return out
    return R


def colwsum(K, w=None, out=None):
    R = K.sum(0) if w is None else K.T @ w
    if out is not None:
        out += R
        return out
    return R


def gemv_n(A, v):
    return A @ v


def potrf_inv(A, overwrite=False):
    L, info = torch.linalg.cholesky_ex(A)
    P = torch.linalg.solve_triangular(L, torch.eye(A.shape[0], dtype=A.dtype), upper=False)
    if overwrite:
        A.copy_(L)
        L = A
    return L, P, info.to(torch.int32)


def gauss_ell(y, mu, q, kdiag, noise, jitter_xx=1e-4, min_var=1e-6, wscale=1.0, want_var=False):
    v = kdiag + jitter_xx + q
    clamped = v < min_var
    v = torch.where(clamped, torch.full_like(v, min_var), v)
    r = y - mu
    e = -0.5 * ((r * r + v) / noise + torch.log(noise) + o.LOG2PI)
    acc = torch.stack([e.sum(), (r * r + v).sum(), (~clamped).double().sum()])
    gmu = wscale * r / noise
    gv = torch.where(clamped, torch.zeros_like(v), -0.5 * wscale / noise * torch.ones_like(v))
    return acc, gmu, gv, (v if want_var else None)


def phi_mask_(X, alpha=1.0):
    R = alpha * (torch.tril(X, -1) + 0.5 * torch.diag(torch.diagonal(X)))
    X.copy_(R)
    return X


def adam_step_(p, g, m, v, step, lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8, gscale=1.0, mask=None):
    gi = gscale * g
    mn = beta1 * m + (1 - beta1) * gi
    vn = beta2 * v + (1 - beta2) * gi * gi
    upd = lr * (mn / (1 - beta1 ** step)) / (torch.sqrt(vn / (1 - beta2 ** step)) + eps)
    if mask is not None:
        keep = mask != 0
        mn, vn, upd = torch.where(keep, mn, m), torch.where(keep, vn, v), torch.where(keep, upd, torch.zeros_like(upd))
    m.copy_(mn)
    v.copy_(vn)
    p -= upd
    return p
