"""Matrix-free (streamed) SGPR objective (nonstationary_precip_b200/sgpr.py) against the oracle's dense collapsed bound."""
import math

import pytest
import torch

from oracle import gibbs_oracle as o

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)


def rel(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


@pytest.mark.parametrize("n,M,D,chunk", [(700, 48, 2, 256), (1000, 130, 3, 333), (1500, 128, 2, 512)])  # M = 128: int8 GEMMs
def test_streamed_sgpr_matches_dense_oracle(n, M, D, chunk):
    from nonstationary_precip_b200.sgpr import SGPRGibbsStream, _inv_softplus
    g = torch.Generator().manual_seed(n)
    x = torch.rand(n, D, generator=g) * 2 - 1
    y = torch.sin(3 * x[:, 0]) + 0.1 * torch.randn(n, generator=g)
    Z = x[torch.randperm(n, generator=g)[:M]].clone()
    le = math.log(0.3) + 0.2 * torch.randn(D, M, generator=g)
    c, os_, lam = torch.full((D,), math.log(0.3)), torch.ones(D), torch.full((D, D), 1.3)
    model = SGPRGibbsStream(Z.cuda(), le.cuda(), c.cuda(), os_.cuda(), lam.cuda(), outputscale=0.8, noise=0.05)
    loss = model.neg_objective_and_grad(x.cuda(), y.cuda(), chunk=chunk)

    Zc, lec = Z.clone().requires_grad_(True), le.clone().requires_grad_(True)
    ro = torch.tensor(_inv_softplus(0.8), requires_grad=True)
    rn = torch.tensor(_inv_softplus(0.05 - 1e-4), requires_grad=True)
    want = -o.sgpr_gibbs_objective(x, y, Zc, lec, o.softplus(ro), 1e-4 + o.softplus(rn), c, os_, lam)
    want.backward()
    assert abs(loss.item() - want.item()) < 1e-8 * abs(want.item())
    assert rel(model.log_ell_z.grad, lec.grad) < 1e-5
    assert rel(model.Z.grad, Zc.grad) < 1e-5
    assert rel(model.raw_outputscale.grad.reshape(()), ro.grad) < 1e-6
    assert rel(model.raw_noise.grad.reshape(()), rn.grad) < 1e-6


def test_streamed_sgpr_rank_sharding_by_emulated_all_reduce():
    """Two 'ranks' each stream half of the rows.  The all-reduce is emulated on one GPU by running each rank's passes and
    exchanging the packed partial sums by hand; the result must equal the single-rank objective and gradients."""
    from nonstationary_precip_b200.sgpr import SGPRGibbsStream
    g = torch.Generator().manual_seed(3)
    n, M, D = 600, 40, 2
    x = (torch.rand(n, D, generator=g) * 2 - 1).cuda()
    y = torch.sin(3 * x[:, 0])
    Z = x[:M].clone()
    le = torch.full((D, M), math.log(0.3)).cuda()
    c, os_, lam = torch.full((D,), math.log(0.3)).cuda(), torch.ones(D).cuda(), torch.full((D, D), 1.3).cuda()
    halves = [(x[:n // 2], y[:n // 2]), (x[n // 2:], y[n // 2:])]
    new = lambda: SGPRGibbsStream(Z, le, c, os_, lam)  # noqa: E731
    full = new()
    l_full = full.neg_objective_and_grad(x, y, chunk=128)

    # round 1: every rank's first-pass sums (A, b, y^T y)
    first = []
    for xs, ys in halves:
        got = []
        new().neg_objective_and_grad(xs, ys, chunk=128, n_total=n, all_reduce=lambda t, got=got: got.append(t.clone()))
        first.append(got[0])
    tot1 = first[0] + first[1]

    # round 2: with the global first-pass sums, every rank's second-pass sums (data-side gradients)
    def reducer(second_of_other=None, capture=None):
        state = {"k": 0}

        def ar(t):
            if state["k"] == 0:
                t.copy_(tot1)
            elif second_of_other is not None:
                t += second_of_other
            elif capture is not None:
                capture.append(t.clone())
            state["k"] += 1
        return ar

    cap = []
    new().neg_objective_and_grad(*halves[1], chunk=128, n_total=n, all_reduce=reducer(capture=cap))
    r0 = new()
    l0 = r0.neg_objective_and_grad(*halves[0], chunk=128, n_total=n, all_reduce=reducer(second_of_other=cap[0]))
    assert abs(l0.item() - l_full.item()) < 1e-10 * abs(l_full.item())
    assert rel(r0.log_ell_z.grad, full.log_ell_z.grad) < 1e-8
    assert rel(r0.Z.grad, full.Z.grad) < 1e-8


@pytest.mark.parametrize("n,M,chunk", [(500, 24, 192), (900, 40, 900), (1200, 64, 512)])  # 2M = 128: int8 GEMMs
def test_streamed_spatio_temporal_sgpr_matches_dense_oracle(n, M, chunk):
    """Config-3 model (Nystrom RBF x Periodic on time + scaled Nystrom Gibbs on lon/lat, one inducing set) evaluated
    matrix free, against the oracle's dense rank-2M bound (spatio_temporal_models.py:35-60)."""
    from nonstationary_precip_b200.sgpr import SGPRSpatioTemporalStream, _inv_softplus
    g = torch.Generator().manual_seed(n + M)
    x = torch.cat([torch.rand(n, 1, generator=g) * 6 - 3, torch.rand(n, 2, generator=g) * 2 - 1], 1)
    y = torch.sin(2 * math.pi * x[:, 0] / 1.7) * torch.exp(-(x[:, 1:] ** 2).sum(-1)) + 0.1 * torch.randn(n, generator=g)
    Z = x[torch.randperm(n, generator=g)[:M]].clone()
    le = math.log(0.4) + 0.2 * torch.randn(2, M, generator=g)
    c, os_, lam = torch.full((2,), math.log(0.4)), torch.ones(2), torch.full((2, 2), 1.3)
    # 64 inducing times under a smooth kernel make the temporal Kzz numerically singular (jitter ladder, objective
    # sensitive at cond * eps); the larger case uses a shorter temporal lengthscale so that the comparison stays sharp
    hyp0 = (0.9, 1.3, 1.7, 7.6) if M < 64 else (0.12, 0.8, 1.7, 7.6)
    model = SGPRSpatioTemporalStream(Z.cuda(), le.cuda(), c.cuda(), os_.cuda(), lam.cuda(), hyp_t=hyp0,
                                     outputscale_s=0.8, noise=0.05)
    loss = model.neg_objective_and_grad(x.cuda(), y.cuda(), chunk=chunk)
    if (2 * M) % 128 == 0:  # this case ran on the int8 tensor-core GEMMs: the FP64 DMMA path must give the same numbers
        ref = SGPRSpatioTemporalStream(Z.cuda(), le.cuda(), c.cuda(), os_.cuda(), lam.cuda(), hyp_t=hyp0,
                                       outputscale_s=0.8, noise=0.05)
        ref.use_i8 = False
        loss_ref = ref.neg_objective_and_grad(x.cuda(), y.cuda(), chunk=chunk)
        assert abs(loss.item() - loss_ref.item()) < 1e-12 * abs(loss_ref.item())
        assert rel(model.raw_hyp_t.grad, ref.raw_hyp_t.grad) < 1e-7 and rel(model.Z.grad, ref.Z.grad) < 1e-7

    Zc, lec = Z.clone().requires_grad_(True), le.clone().requires_grad_(True)
    raw = torch.tensor([_inv_softplus(hyp0[0]), _inv_softplus(hyp0[1]), _inv_softplus(hyp0[2]),
                        _inv_softplus(hyp0[3] - 7.0)], requires_grad=True)
    ro = torch.tensor(_inv_softplus(0.8), requires_grad=True)
    rn = torch.tensor(_inv_softplus(0.05 - 1e-4), requires_grad=True)
    sp = o.softplus(raw)
    hyp = torch.stack([sp[0], sp[1], sp[2], 7.0 + sp[3]])
    # the temporal kernel's inducing points are a frozen alias of Z: no gradient through column 0 of the KERNELS; the prior
    # term, however, is handed the spatial kernel's trainable (M,3) inducing points and its active_dims pick (time, lon)
    # (reference models/spatio_temporal_models.py:52-55), so column 0 does receive the prior's gradient
    Zo = torch.cat([Zc[:, :1].detach(), Zc[:, 1:]], 1)
    want = -o.st_sgpr_objective(x, y, Zo, lec, hyp, o.softplus(ro), 1e-4 + o.softplus(rn), c, os_, lam, Z_prior=Zc)
    want.backward()
    # (with 64 inducing times the Cholesky factors of the two implementations already differ at 1e-8 of the objective)
    assert abs(loss.item() - want.item()) < (1e-8 if M < 64 else 1e-6) * abs(want.item())
    assert rel(model.log_ell_z.grad, lec.grad) < 1e-5
    assert rel(model.Z.grad, Zc.grad) < 1e-5
    assert float(Zc.grad[:, 0].abs().max()) > 0.0  # prior only
    assert rel(model.raw_hyp_t.grad, raw.grad) < 1e-5
    assert rel(model.raw_outputscale.grad.reshape(()), ro.grad) < 1e-6
    assert rel(model.raw_noise.grad.reshape(()), rn.grad) < 1e-6
