"""Exact int8 (Ozaki split) row-quadratic GEMM on tcgen05 (csrc/ozaki.cu) against the FP64 DMMA kernel, a torch fp64
product and, for the whole SVGP step, the oracle."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)


def _case(n, M, seed, spread=3.0):
    g = torch.Generator().manual_seed(seed)
    K = (torch.rand(n, M, generator=g) * torch.exp(spread * torch.randn(n, 1, generator=g))).cuda()
    A = torch.randn(M, M, generator=g).cuda()
    C = A @ A.T / M - 0.3 * torch.eye(M, device="cuda")
    return K, 0.5 * (C + C.T)


@pytest.mark.parametrize("n,M", [(1, 64), (127, 64), (128, 128), (300, 192), (1000, 256), (4099, 1024)])
def test_rowquad_i8_matches_fp64(n, M):
    from nonstationary_precip_b200 import ops
    K, C = _case(n, M, seed=n + M)
    T0, q0 = ops.rowquad(K, C)
    T1, q1 = ops.rowquad_i8(K, C)
    scale = K.abs() @ C.abs()  # the quantity the FP64 rounding bound of a dot product refers to
    assert ((T1 - T0).abs() / scale).max().item() < 4e-15
    assert ((T1 - K @ C).abs() / scale).max().item() < 4e-15
    assert ((q1 - q0).abs() / (scale * K.abs()).sum(1)).max().item() < 4e-15
    T2, q2 = ops.rowquad_i8(K, C, need_q=False)
    assert q2 is None and torch.equal(T2, T1)  # deterministic


def test_rowquad_i8_exact_on_integer_data():
    """Operands whose entries are small integers times powers of two are represented exactly by the slices, so the result
    must equal the exact product bit for bit (checked in integer arithmetic)."""
    from nonstationary_precip_b200 import ops
    g = torch.Generator().manual_seed(5)
    n, M = 256, 128
    Ki = torch.randint(-1000, 1000, (n, M), generator=g)
    Ci = torch.randint(-1000, 1000, (M, M), generator=g)
    Ci = Ci + Ci.T
    T, q = ops.rowquad_i8(Ki.double().cuda() * 2.0 ** -7, Ci.double().cuda() * 2.0 ** 5)
    want = (Ki @ Ci).double() * 2.0 ** -2
    assert torch.equal(T.cpu(), want)
    assert torch.equal(q.cpu(), (want * Ki.double() * 2.0 ** -7).sum(1))


def test_rowquad_i8_edge_values():
    from nonstationary_precip_b200 import ops
    K, C = _case(200, 64, seed=9)
    K[3] = 0.0              # an all-zero row
    K[7] *= 1e-200          # tiny and huge rows keep their relative accuracy (per-row exponents)
    K[9] *= 1e150
    C[:, 5] = 0.0
    C[5, :] = 0.0
    T0, _ = ops.rowquad(K, C)
    T1, q1 = ops.rowquad_i8(K, C)
    assert torch.isfinite(T1).all() and (T1[3] == 0).all() and (T1[:, 5] == 0).all()
    scale = (K.abs() @ C.abs()).clamp_min(1e-300)
    assert ((T1 - T0).abs() / scale).max().item() < 4e-15


def test_rowquad_i8_rejects_unsupported_shapes():
    from nonstationary_precip_b200 import ops
    from nonstationary_precip_b200._lib import NpgpError
    K, C = torch.rand(10, 48).cuda(), torch.eye(48).cuda()
    with pytest.raises(NpgpError):
        ops.rowquad_i8(K, C)


@pytest.mark.parametrize("n,M", [(1, 128), (33, 128), (1000, 128), (5000, 256), (20000, 1024)])
def test_syrk_i8_matches_fp64(n, M):
    from nonstationary_precip_b200 import ops
    g = torch.Generator().manual_seed(n + M)
    K = (torch.rand(n, M, generator=g) * torch.exp(2 * torch.randn(1, M, generator=g))).cuda()
    w0 = torch.tensor([-0.37], device="cuda")
    got = ops.syrk_i8(K, w0=w0, alpha=2.0)
    want = -0.74 * (K.T @ K)
    scale = 0.74 * (K.abs().T @ K.abs())
    assert ((got - want).abs() / scale).max().item() < 1e-14
    assert torch.equal(got, got.T)
    ref = ops.wsyrk(K, alpha=-0.74)
    assert ((got - ref).abs() / scale).max().item() < 1e-14


def test_syrk_i8_exact_on_integer_data():
    from nonstationary_precip_b200 import ops
    g = torch.Generator().manual_seed(11)
    Ki = torch.randint(-2000, 2000, (3000, 128), generator=g)
    got = ops.syrk_i8(Ki.double().cuda() * 2.0 ** -9)
    assert torch.equal(got.cpu(), (Ki.T @ Ki).double() * 2.0 ** -18)


def test_wsyrk_i8_device_gate():
    """Equal weights -> integer tensor-core path; one differing weight -> the FP64 weighted kernel; both enqueued."""
    from nonstationary_precip_b200 import ops
    g = torch.Generator().manual_seed(12)
    n, M = 3000, 256
    K = torch.rand(n, M, generator=g).cuda()
    w = torch.full((n,), -0.25, device="cuda")
    cnt = torch.tensor([float(n)], device="cuda")
    got = ops.wsyrk_i8(K, w, cnt, float(n))
    want = -0.25 * (K.T @ K)
    assert ((got - want).abs().max() / want.abs().max()).item() < 1e-13
    w2 = w.clone()
    w2[17] = 0.0
    cnt2 = torch.tensor([float(n - 1)], device="cuda")
    got2 = ops.wsyrk_i8(K, w2, cnt2, float(n))
    want2 = K.T @ (w2[:, None] * K)
    assert ((got2 - want2).abs().max() / want2.abs().max()).item() < 1e-13
    assert torch.equal(got2, got2.T)


@pytest.mark.parametrize("variant", ["full", "diag"])
def test_svgp_step_with_i8_rowquad_matches_dmma_and_oracle(variant):
    """Whole ELBO step (loss + flat gradient) and prediction with rowquad_impl='i8': against the DMMA step on the same
    state, and against the oracle at the north-star tolerances."""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from svgp_cases import make_problem, oracle_loss_and_grads
    from nonstationary_precip_b200.svgp import SVGPGibbs
    x, y, Z, p, N = make_problem(variant, device="cuda", B=700, M=128, d=3, seed=21)
    outs = []
    for impl in ("dmma", "i8"):
        model = SVGPGibbs(variant, Z, N, **p)
        model.rowquad_impl = impl
        loss = model.loss_and_grad(x, y)
        assert int(model.last["info"]) == 0
        xs = torch.rand(500, 3, device="cuda") * 2 - 1
        outs.append((loss.item(), model.grad.clone(), model.predict(xs.clone().fill_(0.25) + 0.5 * x[:500]), model))
    (l0, g0, (m0, v0), _), (l1, g1, (m1, v1), model) = outs
    assert abs(l0 - l1) < 1e-11 * abs(l0)
    assert (g0 - g1).abs().max().item() < 1e-8 * g0.abs().max().item()
    assert (m0 - m1).abs().max().item() < 1e-10 and ((v0 - v1).abs() / v0).max().item() < 1e-9
    want_loss, want = oracle_loss_and_grads(variant, x, y, Z, p, N)
    assert abs(l1 - want_loss.item()) < 1e-9 * abs(want_loss.item())
    for name, gw in want.items():
        g = model.g[name].detach().cpu()
        assert ((g - gw).abs().max() / gw.abs().max().clamp_min(1e-300)).item() < 1e-6, name
