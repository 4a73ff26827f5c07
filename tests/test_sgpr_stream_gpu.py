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


@pytest.mark.parametrize("n,M,D,chunk", [(700, 48, 2, 256), (1000, 130, 3, 333)])
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
    """Two 'ranks' each stream half of the rows; with the sums exchanged the objective equals the single-rank one."""
    from nonstationary_precip_b200.sgpr import SGPRGibbsStream
    g = torch.Generator().manual_seed(3)
    n, M, D = 600, 40, 2
    x = (torch.rand(n, D, generator=g) * 2 - 1).cuda()
    y = torch.sin(3 * x[:, 0])
    Z = x[:M].clone()
    le = torch.full((D, M), math.log(0.3)).cuda()
    c, os_, lam = torch.full((D,), math.log(0.3)).cuda(), torch.ones(D).cuda(), torch.full((D, D), 1.3).cuda()
    full = SGPRGibbsStream(Z, le, c, os_, lam)
    l_full = full.neg_objective_and_grad(x, y, chunk=128)
    # emulate the all-reduce: rank 1's partial sums are precomputed and added to rank 0's
    r1 = SGPRGibbsStream(Z, le, c, os_, lam)
    stash = []
    r1.neg_objective_and_grad(x[n // 2:], y[n // 2:], chunk=128, n_total=n, all_reduce=lambda t: stash.append(t.clone()))
    it = iter(stash)
    r0 = SGPRGibbsStream(Z, le, c, os_, lam)

    def fake_all_reduce(t):
        other = next(it)
        t += other

    # rank 1's second-pass sums depend on dA/db from the GLOBAL objective, so recompute them with rank 0's sums first
    stash0 = []
    r0.neg_objective_and_grad(x[:n // 2], y[:n // 2], chunk=128, n_total=n, all_reduce=lambda t: stash0.append(t.clone()))
    tot1 = stash[0] + stash0[0]

    class TwoPass:
        def __init__(self, other_second=None):
            self.k, self.other_second = 0, other_second

        def __call__(self, t):
            if self.k == 0:
                t.copy_(tot1)
            elif self.other_second is not None:
                t += self.other_second
            self.k += 1

    cap = []
    r1b = SGPRGibbsStream(Z, le, c, os_, lam)

    def ar1(t, st=[0]):
        if st[0] == 0:
            t.copy_(tot1)
        else:
            cap.append(t.clone())
        st[0] += 1

    r1b.neg_objective_and_grad(x[n // 2:], y[n // 2:], chunk=128, n_total=n, all_reduce=ar1)
    r0b = SGPRGibbsStream(Z, le, c, os_, lam)
    l0 = r0b.neg_objective_and_grad(x[:n // 2], y[:n // 2], chunk=128, n_total=n, all_reduce=TwoPass(cap[0]))
    assert abs(l0.item() - l_full.item()) < 1e-10 * abs(l_full.item())
    assert rel(r0b.log_ell_z.grad, full.log_ell_z.grad) < 1e-8
    assert rel(r0b.Z.grad, full.Z.grad) < 1e-8
