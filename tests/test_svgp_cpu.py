"""The product's SVGP-Gibbs step (nonstationary_precip_b200/svgp.py: analytic backward, flat buffers, rank weighting)
run on CPU through tests/cpu_ops_emulator.py and checked against autograd of the oracle's ELBO."""
import pytest
import torch

import cpu_ops_emulator as emu
from nonstationary_precip_b200.svgp import SVGPGibbs
from svgp_cases import make_problem, oracle_loss_and_grads

torch.set_default_dtype(torch.float64)


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


@pytest.mark.parametrize("variant,d", [("diag", 3), ("diag", 2), ("full", 3), ("full", 2)])
def test_loss_and_analytic_gradients_match_oracle_autograd(variant, d):
    x, y, Z, kw, N = make_problem(variant, B=80, M=20, d=d)
    model = SVGPGibbs(variant, Z, N, ops=emu, **kw)
    loss = model.loss_and_grad(x, y)
    want_loss, want = oracle_loss_and_grads(variant, x, y, Z, kw, N)
    assert abs(loss.item() - want_loss.item()) < 1e-9 * abs(want_loss.item())
    for name, gw in want.items():
        assert rel(model.g[name], gw) < 1e-7, name


def test_rank_sharded_gradients_sum_to_single_rank():
    """Rows split over 2 'ranks': the sum of the flat gradient buffers equals the 1-rank buffer."""
    for variant in ("diag", "full"):
        x, y, Z, kw, N = make_problem(variant, B=64, M=16, d=3, seed=3)
        full = SVGPGibbs(variant, Z, N, ops=emu, **kw)
        full.loss_and_grad(x, y)
        tot = torch.zeros_like(full.grad)
        for r in range(2):
            part = SVGPGibbs(variant, Z, N, ops=emu, **kw)
            part.loss_and_grad(x[r::2].contiguous(), y[r::2].contiguous(), world_size=2)
            tot += part.grad
        assert rel(tot, full.grad) < 1e-11


def test_adam_step_matches_torch_optim():
    x, y, Z, kw, N = make_problem("diag", B=48, M=12, d=3, seed=5)
    model = SVGPGibbs("diag", Z, N, ops=emu, **kw)
    ref = torch.nn.Parameter(model.theta.clone())
    opt = torch.optim.Adam([ref], lr=0.01)
    for _ in range(3):
        model.loss_and_grad(x, y)
        ref.grad = (model.grad[:ref.numel()] * model.mask).clone()
        model.adam_step(0.01)
        opt.step()
    # frozen (masked) entries: torch moves nothing for zero grads either
    assert rel(model.theta, ref.detach()) < 1e-12


def test_loss_decreases_under_training():
    x, y, Z, kw, N = make_problem("full", B=128, M=16, d=3, seed=7)
    model = SVGPGibbs("full", Z, N, ops=emu, **kw)
    losses = [model.train_step(x, y, lr=0.02).item() for _ in range(25)]
    assert losses[-1] < losses[0]
