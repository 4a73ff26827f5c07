"""Parity of every CUDA op (called through the C ABI) against the CPU oracle / plain fp64 torch on the same inputs."""
import os

import pytest
import torch

from oracle import gibbs_oracle as o

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)


@pytest.fixture(scope="module")
def ops():
    from nonstationary_precip_b200 import ops as _ops
    return _ops


def rel(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def dev(*ts):
    return [t.cuda() for t in ts]


# ---------------------------------------------------------------------------------------------------------------- diag
@pytest.mark.parametrize("tag", ["d1", "d2", "d3", "d5"])
def test_gibbs_diag_fwd_golden(ops, golden, tag):
    g = golden("gibbs_diag_" + tag)
    x1, x2, e1, e2 = dev(g["x1"], g["x2"], g["ell1"], g["ell2"])
    assert rel(ops.gibbs_diag_fwd(x1, e1, x2, e2), g["K12"]) < 1e-12
    K11 = ops.gibbs_diag_fwd(x1, e1, x1, e1)
    assert rel(K11, g["K11"]) < 1e-12
    assert (torch.diagonal(K11).cpu() - 1).abs().max() < 1e-14


@pytest.mark.parametrize("n1,n2,D", [(1, 1, 2), (33, 257, 3), (300, 1023, 3), (1000, 511, 2), (64, 1024, 1)])
def test_gibbs_diag_fwd_bwd_random(ops, n1, n2, D):
    g = torch.Generator().manual_seed(n1 * 7 + n2)
    x1 = torch.rand(n1, D, generator=g) * 2 - 1
    x2 = torch.rand(n2, D, generator=g) * 2 - 1
    e1 = torch.exp(0.4 * torch.randn(D, n1, generator=g) - 1)
    e2 = torch.exp(0.4 * torch.randn(D, n2, generator=g) - 1)
    s = torch.tensor(0.644)
    u = torch.randn(n2, generator=g)
    G = torch.randn(n1, n2, generator=g)
    cx1, ce1, cx2, ce2, cs = [t.clone().requires_grad_(True) for t in (x1, e1, x2, e2, s)]
    Kc = cs * o.gibbs_diag_K(cx1, cx2, ce1, ce2)
    (Kc * G).sum().backward()
    K, Ku = ops.gibbs_diag_fwd(*dev(x1, e1, x2, e2, s), u=u.cuda())
    assert rel(K, Kc) < 1e-12
    assert rel(Ku, Kc.detach() @ u) < 1e-11
    gx1, ge1, gx2, ge2, gs = [t.clone().requires_grad_(True) for t in dev(x1, e1, x2, e2, s)]
    Kg = ops.gibbs_diag(gx1, ge1, gx2, ge2, gs)
    (Kg * G.cuda()).sum().backward()
    for name, a, b in (("x1", gx1, cx1), ("ell1", ge1, ce1), ("x2", gx2, cx2), ("ell2", ge2, ce2), ("scale", gs, cs)):
        assert rel(a.grad, b.grad) < 1e-10, name


def test_gibbs_diag_bwd_composed_gradient(ops):
    """G_ij = rowscale_i * T_ij + rowvec_i * colvec_j formed inside the kernel."""
    g = torch.Generator().manual_seed(5)
    n1, n2, D = 130, 300, 3
    x1, x2 = torch.rand(n1, D, generator=g), torch.rand(n2, D, generator=g)
    e1, e2 = torch.rand(D, n1, generator=g) + 0.2, torch.rand(D, n2, generator=g) + 0.2
    T, rs, rv, cv = (torch.randn(n1, n2, generator=g), torch.randn(n1, generator=g), torch.randn(n1, generator=g),
                     torch.randn(n2, generator=g))
    G = rs[:, None] * T + rv[:, None] * cv[None, :]
    a = ops.gibbs_diag_bwd(*dev(x1, e1, x2, e2), G=G.cuda(), need_dx2=True, need_dscale=True)
    b = ops.gibbs_diag_bwd(*dev(x1, e1, x2, e2), G=T.cuda(), rowscale=rs.cuda(), rowvec=rv.cuda(), colvec=cv.cuda(),
                           need_dx2=True, need_dscale=True)
    for k in ("d_ell1", "d_ell2", "d_x2", "d_scale"):
        assert rel(b[k], a[k]) < 1e-11, k


# ---------------------------------------------------------------------------------------------------------------- full
def test_gibbs_full_fwd_golden(ops, golden):
    g = golden("gibbs_full_d2_f64")
    S1, S2 = o.sigma_from_H(g["H1"], g["Dm"]), o.sigma_from_H(g["H2"], g["Dm"])
    x1, x2 = dev(g["x1"], g["x2"])
    K = ops.gibbs_full_fwd(x1, ops.sym_pack(S1.cuda()), x2, ops.sym_pack(S2.cuda()))
    assert rel(K, g["K12"]) < 1e-12
    K11 = ops.gibbs_full_fwd(x1, ops.sym_pack(S1.cuda()), x1, ops.sym_pack(S1.cuda()))
    assert rel(K11, g["K11"]) < 1e-12
    Sg = ops.sigma_from_h_fwd(g["H1"].cuda(), g["Dm"].cuda())
    assert rel(Sg, ops.sym_pack(S1)) < 1e-14
    g32 = golden("gibbs_full_d2_f32sigma")  # the reference as shipped (float32 Sigma): secondary 1e-6 check
    assert rel(K, g32["K12"].double()) < 5e-6


def test_gibbs_full_fwd_golden_d3(ops, golden):
    """d = 3 fixture produced by the reference's own forward lines (tests/golden/make_golden.py multivariate_d3_case)."""
    g = golden("gibbs_full_d3_f64")
    x1, x2, Dm = dev(g["x1"], g["x2"], g["Dm"])
    S1, S2 = ops.sigma_from_h_fwd(g["H1"].cuda(), Dm), ops.sigma_from_h_fwd(g["H2"].cuda(), Dm)
    assert rel(ops.gibbs_full_fwd(x1, S1, x2, S2), g["K12"]) < 1e-12
    assert rel(ops.gibbs_full_fwd(x1, S1, x1, S1), g["K11"]) < 1e-12


@pytest.mark.parametrize("n1,n2,d", [(1, 3, 2), (45, 258, 2), (200, 515, 3), (513, 1024, 3)])
def test_gibbs_full_fwd_bwd_random(ops, n1, n2, d):
    g = torch.Generator().manual_seed(n1 + n2)
    x1 = torch.rand(n1, d, generator=g) * 2 - 1
    x2 = torch.rand(n2, d, generator=g) * 2 - 1
    H1, H2 = torch.randn(n1, d, generator=g), torch.randn(n2, d, generator=g)
    Dm = torch.diag(0.3 + torch.rand(d, generator=g))
    s = torch.tensor(0.644)
    G = torch.randn(n1, n2, generator=g)
    u = torch.randn(n2, generator=g)
    cx1, cH1, cx2, cH2, cD, cs = [t.clone().requires_grad_(True) for t in (x1, H1, x2, H2, Dm, s)]
    Kc = cs * o.gibbs_full_K(cx1, cx2, o.sigma_from_H(cH1, cD), o.sigma_from_H(cH2, cD))
    (Kc * G).sum().backward()
    gx1, gH1, gx2, gH2, gD, gs = [t.clone().requires_grad_(True) for t in dev(x1, H1, x2, H2, Dm, s)]
    Kg = ops.gibbs_full(gx1, ops.sigma_from_h(gH1, gD), gx2, ops.sigma_from_h(gH2, gD), gs)
    assert rel(Kg, Kc) < 1e-12
    (Kg * G.cuda()).sum().backward()
    for name, a, b in (("x1", gx1, cx1), ("H1", gH1, cH1), ("x2", gx2, cx2), ("H2", gH2, cH2), ("D", gD, cD),
                       ("scale", gs, cs)):
        assert rel(a.grad, b.grad) < 1e-9, name
    K2, Ku = ops.gibbs_full_fwd(gx1.detach(), ops.sigma_from_h_fwd(gH1.detach(), gD.detach()), gx2.detach(),
                                ops.sigma_from_h_fwd(gH2.detach(), gD.detach()), scale=gs.detach(), u=u.cuda())
    assert rel(Ku, Kc.detach() @ u) < 1e-11


# --------------------------------------------------------------------------------------------------------------- field
@pytest.mark.parametrize("tag", ["d2", "d3"])
def test_field_interp_diag_golden(ops, golden, tag):
    g = golden("lognormal_field_" + tag)
    D = g["c"].shape[0]
    m = g["xg"].shape[0]
    Kgg = o.rbf_ard_K(g["xg"], g["xg"], g["lam"], g["os"]) + 1e-4 * torch.eye(m)
    alpha = torch.linalg.solve(Kgg, (torch.log(g["ell_g"]) - g["c"][:, None]).unsqueeze(-1)).squeeze(-1)
    out = ops.rbf_matvec_fwd(g["x"].cuda(), g["xg"].cuda(), g["lam"].cuda(), g["os"].cuda(), alpha.cuda().unsqueeze(-1),
                             bias=g["c"].cuda(), apply_exp=True)
    assert rel(out.squeeze(-1), g["ell_x"]) < 1e-11


@pytest.mark.parametrize("d,nb,nv,n,m", [(3, 3, 1, 500, 300), (3, 1, 3, 257, 1024), (2, 2, 1, 31, 65), (2, 1, 2, 1, 7)])
def test_rbf_matvec_fwd_bwd(ops, d, nb, nv, n, m):
    g = torch.Generator().manual_seed(n + m)
    x, z = torch.rand(n, d, generator=g) * 2 - 1, torch.rand(m, d, generator=g) * 2 - 1
    lam, os_ = 0.7 + torch.rand(nb, d, generator=g), 0.5 + torch.rand(nb, generator=g)
    V, bias = torch.randn(nb, m, nv, generator=g), torch.randn(nb, generator=g)
    W = torch.randn(nb, n, nv, generator=g)
    cz, cV = z.clone().requires_grad_(True), V.clone().requires_grad_(True)
    Kc = o.rbf_ard_K(x, cz, lam, os_)  # (nb,n,m)
    outc = torch.exp(bias[:, None, None] + Kc @ cV)
    (outc * W).sum().backward()
    gz, gV = z.cuda().requires_grad_(True), V.cuda().requires_grad_(True)
    outg = ops.rbf_matvec(x.cuda(), gz, lam.cuda(), os_.cuda(), gV, bias.cuda(), True)
    assert rel(outg, outc) < 1e-12
    (outg * W.cuda()).sum().backward()
    assert rel(gV.grad, cV.grad) < 1e-11
    assert rel(gz.grad, cz.grad) < 1e-11


# ---------------------------------------------------------------------------------------------------------------- gemm
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1, 2, 2), (257, 130, 77), (1024, 1024, 1024), (300, 64, 1030)])
@pytest.mark.parametrize("tA,tB", [(False, False), (False, True), (True, False), (True, True)])
def test_dgemm_layouts(ops, M, N, K, tA, tB):
    g = torch.Generator().manual_seed(M + N + K)
    Kp = K + (K % 2)
    A = torch.randn((Kp, M + M % 2) if tA else (M, Kp), generator=g).cuda()
    B = torch.randn((N, Kp) if tB else (Kp, N + N % 2), generator=g).cuda()
    Av = A[:K, :M] if tA else A[:, :K]
    Bv = B[:, :K] if tB else B[:K, :N]
    C0 = torch.randn(M, N, generator=g).cuda()
    want = 0.7 * (Av.T if tA else Av) @ (Bv.T if tB else Bv) - 0.3 * C0
    got = ops.dgemm(Av, Bv, tA, tB, alpha=0.7, beta=-0.3, C=C0.clone())
    assert rel(got, want) < 1e-13


def test_dgemm_triangular_skipping(ops):
    g = torch.Generator().manual_seed(0)
    M = 640
    L = torch.tril(torch.randn(M, M, generator=g)).cuda()
    X = torch.randn(M, M, generator=g).cuda()
    assert rel(ops.dgemm(L, X, tri_a=1), L @ X) < 1e-13
    assert rel(ops.dgemm(L, X, transA=True, tri_a=2), L.T @ X) < 1e-13
    assert rel(ops.dgemm(X, L, tri_b=1), X @ L) < 1e-13
    assert rel(ops.dgemm(X, L, transB=True, tri_b=2), X @ L.T) < 1e-13
    S = ops.dgemm(X, X, transB=True, out_tri=1, C=torch.zeros(M, M, device="cuda"), beta=0.0)
    assert rel(torch.tril(S), torch.tril(X @ X.T)) < 1e-13


@pytest.mark.parametrize("n,M", [(2048, 1024), (777, 256), (130, 130)])
def test_rowquad_and_wsyrk(ops, n, M):
    g = torch.Generator().manual_seed(n)
    Mp = M + M % 2
    K = torch.randn(n, Mp, generator=g).cuda()[:, :M]
    Cm = torch.randn(Mp, Mp, generator=g).cuda()[:M, :M]
    Cm = Cm + Cm.T
    w = torch.randn(n, generator=g).cuda()
    T, q = ops.rowquad(K, Cm)
    Tw = K @ Cm
    assert rel(T, Tw) < 1e-13
    assert rel(q, (Tw * K).sum(-1)) < 1e-12
    assert rel(ops.wsyrk(K, w, alpha=-0.5), -0.5 * K.T @ (w[:, None] * K)) < 1e-12
    assert rel(ops.wsyrk(K), K.T @ K) < 1e-12


# ---------------------------------------------------------------------------------------------------------------- chol
@pytest.mark.parametrize("impl", ["flow", "steps"])
@pytest.mark.parametrize("M", [2, 64, 100, 300, 1024, 1536, 2048])
def test_potrf_inv(ops, M, impl):
    g = torch.Generator().manual_seed(M)
    X = torch.rand(M, 3, generator=g) * 2 - 1
    ell = torch.full((3, M), 0.3)
    A = (o.gibbs_diag_K(X, X, ell, ell) + 1e-6 * torch.eye(M)).cuda()
    L, P, info = ops.potrf_inv(A, impl=impl)
    assert int(info) == 0
    Lw = torch.linalg.cholesky(A)
    assert rel(L, Lw) < 1e-9  # forward error ~ cond * eps
    assert rel(L @ L.T, A) < 1e-14
    eye = torch.eye(M, device="cuda")
    assert (P @ L - eye).abs().max() < 1e-8
    assert torch.triu(L, 1).abs().max() == 0 and torch.triu(P, 1).abs().max() == 0


@pytest.mark.parametrize("impl", ["flow", "steps"])
def test_potrf_reports_first_bad_pivot(ops, impl):
    M = 200
    A = torch.eye(M, device="cuda")
    A[130, 130] = -1.0
    _, _, info = ops.potrf_inv(A, impl=impl)
    assert int(info) == 131


@pytest.mark.parametrize("n,M", [(2, 1024), (4, 1024), (3, 300), (4, 64)])
def test_potrf_flow_batch(ops, n, M):
    """All factorisations of a step (Kzz + the field prior's kernel matrices) in ONE interleaved dataflow launch."""
    g = torch.Generator().manual_seed(7 * n + M)
    mats = []
    for k in range(n):
        X = torch.rand(M, 3, generator=g) * 2 - 1
        ell = torch.full((3, M), 0.25 + 0.05 * k)
        mats.append((o.gibbs_diag_K(X, X, ell, ell) + 1e-6 * torch.eye(M)).cuda())
    if n == 4:
        mats[2] = mats[2].clone()
        mats[2][M // 2, M // 2] = -1.0  # one failing matrix must not disturb the others
    eye = torch.eye(M, device="cuda")
    for _ in range(2):
        outs = ops.potrf_inv_batch(mats)
    for k, ((L, P, info), A) in enumerate(zip(outs, mats)):
        if n == 4 and k == 2:
            assert int(info) == M // 2 + 1
            continue
        assert int(info) == 0
        assert rel(L @ L.T, A) < 1e-14 and (P @ L - eye).abs().max() < 1e-8
        L1, P1, _ = ops.potrf_inv(A, impl="flow")
        assert torch.equal(L, L1) and torch.equal(P, P1)  # same tile arithmetic as the single-matrix launch


def test_potrf_flow_two_factorisations_on_two_streams(ops):
    """The SVGP step factors Kzz and the field prior's kernel matrix concurrently: two dataflow kernels sharing the SMs."""
    g = torch.Generator().manual_seed(5)
    mats = []
    for _ in range(2):
        X = torch.rand(1024, 3, generator=g) * 2 - 1
        ell = torch.full((3, 1024), 0.3)
        mats.append((o.gibbs_diag_K(X, X, ell, ell) + 1e-6 * torch.eye(1024)).cuda())
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = []
    for _ in range(3):
        with torch.cuda.stream(s1):
            r1 = ops.potrf_inv(mats[0], impl="flow")
        with torch.cuda.stream(s2):
            r2 = ops.potrf_inv(mats[1], impl="flow")
        outs = [r1, r2]
    torch.cuda.synchronize()
    for (L, P, info), A in zip(outs, mats):
        assert int(info) == 0
        assert rel(L @ L.T, A) < 1e-14
        assert (P @ L - torch.eye(1024, device="cuda")).abs().max() < 1e-8


def test_wsyrk_uniform_weight_hint(ops):
    """Device-side hint: equal weights -> unweighted inner loop with the weight folded into the epilogue; if the hint
    does not hold the weights are applied per fragment."""
    g = torch.Generator().manual_seed(12)
    n, M = 1500, 192
    K = torch.randn(n, M, generator=g).cuda()
    w_const = torch.full((n,), -0.37, device="cuda")
    count = torch.tensor([float(n)], device="cuda")
    want = K.T @ (w_const[:, None] * K)
    assert rel(ops.wsyrk(K, w_const, uniform_count=count, uniform_target=float(n)), want) < 1e-12
    w_var = w_const.clone()
    w_var[::7] = 0.0  # e.g. clamped variances
    count2 = torch.tensor([float((w_var != 0).sum())], device="cuda")
    want2 = K.T @ (w_var[:, None] * K)
    assert rel(ops.wsyrk(K, w_var, uniform_count=count2, uniform_target=float(n)), want2) < 1e-12
