"""Direct checks of the four helpers of the reference's utils/functional.py that the hot path uses (op, dot, mv, t;
reference utils/functional.py:14-33,60-64), through both import paths (package and compat top-level `utils.functional`)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)


def _ref_op(v1, v2=None):  # the reference's definition, restated: batched (…,n,1) @ (…,1,m)
    v2 = v1 if v2 is None else v2
    return v1.unsqueeze(-1) @ v2.unsqueeze(-2)


@pytest.mark.parametrize("via_compat", [False, True])
def test_op_dot_mv_t(via_compat):
    if via_compat:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compat"))
        import utils.functional as fn
    else:
        from nonstationary_precip_b200.utils import functional as fn
    g = torch.Generator().manual_seed(0)
    a = torch.randn(5, 7, generator=g).cuda()
    b = torch.randn(5, 7, generator=g).cuda()
    assert torch.allclose(fn.dot(a, b), (a * b).sum(-1), rtol=0, atol=1e-15)
    assert torch.equal(fn.t(a), a.transpose(-1, -2))
    assert torch.allclose(fn.op(a, b), _ref_op(a, b), rtol=1e-15, atol=0) and fn.op(a).shape == (5, 7, 7)
    # mv: matrix-vector product on the FP64 tensor pipe, and the solve used by LogNormalPriorProcess.conditional_sample
    # (reference gibbs_kernels.py:89-93 calls fn.mv(K + jitter, rhs, invert=True) on an SPD matrix)
    A = torch.randn(300, 300, generator=g).cuda()
    v = torch.randn(300, generator=g).cuda()
    assert ((fn.mv(A, v) - A @ v).abs().max() / (A.abs() @ v.abs()).max()).item() < 1e-14
    S = A @ A.T / 300 + 0.5 * torch.eye(300, device="cuda")
    x = fn.mv(S, v, invert=True)
    want = torch.linalg.solve(S, v)  # the reference's LU solve
    assert ((x - want).abs().max() / want.abs().max()).item() < 1e-11
    # gradients flow through mv (the field interpolation is differentiated w.r.t. the inducing lengthscales)
    vr = v.clone().requires_grad_(True)
    fn.mv(S, vr, invert=True).sum().backward()
    assert ((vr.grad - torch.linalg.solve(S, torch.ones(300, device="cuda"))).abs().max()).item() < 1e-10


@pytest.mark.parametrize("M,k", [(200, 6), (512, 64)])
def test_trsm_through_the_inverse_factor(M, k):
    """npgp_trsm: X = L^-1 B / L^-T B with P = L^-1 from npgp_potrf_inv_* (the reference's triangular_solve(eye, chol) + matmul,
    models/gibbs_kernels.py:205-208,222-225) against torch.linalg.solve_triangular."""
    from nonstationary_precip_b200 import ops
    from nonstationary_precip_b200._lib import check, lib, ptr, stream
    g = torch.Generator().manual_seed(M)
    A = torch.randn(M, M, generator=g, dtype=torch.float64)
    K = (A @ A.T / M + torch.eye(M, dtype=torch.float64)).cuda()
    L, P, info = ops.potrf_inv(K.clone())
    assert int(info) == 0
    B = torch.randn(M, k, generator=g, dtype=torch.float64).cuda()
    for trans in (0, 1):
        X = torch.empty_like(B)
        check(lib().npgp_trsm(trans, M, k, ptr(P), P.stride(0), ptr(B), B.stride(0), ptr(X), X.stride(0), stream()), "npgp_trsm")
        want = torch.linalg.solve_triangular(L.T if trans else L, B, upper=bool(trans))
        assert ((X - want).abs().max() / want.abs().max()).item() < 1e-12
