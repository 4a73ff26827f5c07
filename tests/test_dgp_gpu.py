"""Deep GP with DSVI (mirror of models/dgps.py on the CUDA kernels) against the oracle's dgp_elbo with the same
parameters and the same N(0,1) draws; plus statistical checks of the Philox sampling kernel."""
import pytest
import torch

from oracle import gibbs_oracle as o

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)


def rel(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def layer_dict(layer, requires_grad=True):
    vs, vd = layer.variational_strategy, layer.variational_strategy._variational_distribution
    c = lambda t: t.detach().cpu().clone().requires_grad_(requires_grad)
    d = dict(Z=c(vs.inducing_points), m=c(vd.variational_mean), Ls=c(vd.chol_variational_covar),
             raw_os=c(layer.covar_module.raw_outputscale),
             raw_ls=c(layer.covar_module.base_kernel.raw_lengthscale.squeeze(-2)))
    if hasattr(layer.mean_module, "weights"):
        d["W"], d["b"] = c(layer.mean_module.weights), c(layer.mean_module.bias)
    else:
        d["c"] = c(layer.mean_module.constant.reshape(()))
    return d


@pytest.mark.parametrize("c_layer", [True, False])
@pytest.mark.parametrize("num_layers,B,M,S", [(1, 300, 24, 4), (2, 300, 24, 4), (1, 2048, 128, 4)])
def test_dgp_elbo_and_gradients_match_oracle(num_layers, B, M, S, c_layer, monkeypatch):
    """c_layer: every layer output through npgp_dsvi_layer_fwd / _bwd (one C call each) or through the kernel-by-kernel
    composition.  (1, 2048, 128, 4): the last layer sees S * B = 8192 rows of width 128 -- T = K C and the equal-weights
    K^T diag(dvar) K of its backward run on the int8 tensor cores."""
    from nonstationary_precip_b200.models import dgps
    monkeypatch.setattr(dgps, "USE_C_LAYER", c_layer)
    torch.manual_seed(5)
    d, N = 2, 4 * B
    g = torch.Generator().manual_seed(6)
    x = torch.rand(B, d, generator=g) * 2 - 1
    y = torch.sin(3 * x[:, 0]) + 0.1 * torch.randn(B, generator=g)
    model = dgps.DeepGP(num_layers, x.shape, num_inducing=M).cuda().double()
    with torch.no_grad():  # non-trivial variational parameters
        for layer in model.variational_layers():
            vd = layer.variational_strategy._variational_distribution
            vd.variational_mean.add_(0.3 * torch.randn_like(vd.variational_mean))
            vd.chol_variational_covar.mul_(0.7).add_(0.05 * torch.tril(torch.randn_like(vd.chol_variational_covar)))
    eps = [torch.randn(S, B, 2, generator=g) for _ in range(num_layers)]
    mll = dgps.DeepApproximateMLL(dgps.VariationalELBO(model.likelihood, model, N))
    with dgps.num_likelihood_samples(S):
        out = model(x.cuda(), eps=[e.cuda() for e in eps])
        elbo = mll(out, y.cuda())
    elbo.backward()

    hid = layer_dict(model.layers[0])
    last = layer_dict(model.last_layer)
    raw_noise = model.likelihood.noise_covar.raw_noise.detach().cpu().clone().requires_grad_(True)
    want = o.dgp_elbo(x, y, N, [hid] * num_layers, last, raw_noise[0], eps)
    want.backward()
    assert abs(elbo.item() - want.item()) < 1e-8 * abs(want.item())
    hl, ll = model.layers[0], model.last_layer
    pairs = [
        (hl.variational_strategy.inducing_points.grad, hid["Z"].grad, "Z_hidden"),
        (hl.variational_strategy._variational_distribution.variational_mean.grad, hid["m"].grad, "m_hidden"),
        (torch.tril(hl.variational_strategy._variational_distribution.chol_variational_covar.grad),
         torch.tril(hid["Ls"].grad), "Ls_hidden"),
        (hl.covar_module.raw_outputscale.grad, hid["raw_os"].grad, "os_hidden"),
        (hl.covar_module.base_kernel.raw_lengthscale.grad.squeeze(-2), hid["raw_ls"].grad, "ls_hidden"),
        (hl.mean_module.weights.grad, hid["W"].grad, "W"),
        (ll.variational_strategy.inducing_points.grad, last["Z"].grad, "Z_last"),
        (ll.variational_strategy._variational_distribution.variational_mean.grad, last["m"].grad, "m_last"),
        (ll.covar_module.base_kernel.raw_lengthscale.grad.reshape(-1), last["raw_ls"].grad.reshape(-1), "ls_last"),
        (model.likelihood.noise_covar.raw_noise.grad, raw_noise.grad, "noise"),
    ]
    for a, b, name in pairs:
        assert rel(a, b) < 1e-6, name


def test_dsvi_sampling_kernel_statistics_and_determinism():
    from nonstationary_precip_b200 import ops
    n = 1 << 20
    mu = torch.full((n,), 2.0, device="cuda")
    var = torch.full((n,), 0.25, device="cuda")
    h1 = ops.dsvi_sample(mu, var, None, seed=11, offset=0)
    h2 = ops.dsvi_sample(mu, var, None, seed=11, offset=0)
    h3 = ops.dsvi_sample(mu, var, None, seed=12, offset=0)
    assert torch.equal(h1, h2) and not torch.equal(h1, h3)
    # sharding invariance: the second half generated with an index offset equals the second half of the full draw
    hb = ops.dsvi_sample(mu[n // 2:], var[n // 2:], None, seed=11, offset=n // 2)
    assert torch.equal(hb, h1[n // 2:])
    z = (h1 - 2.0) / 0.5
    assert abs(z.mean().item()) < 5e-3 and abs(z.var().item() - 1.0) < 5e-3
    assert abs((z ** 3).mean().item()) < 2e-2 and abs((z ** 4).mean().item() - 3.0) < 5e-2


def test_dgp_sample_sharding_is_world_size_invariant():
    """SURVEY 8(e), DSVI row: S likelihood samples split over 2 and 4 emulated ranks; the summed loss and gradients must
    equal the single-rank step (same Philox draws by global element index)."""
    from nonstationary_precip_b200.models import dgps
    torch.manual_seed(11)
    g = torch.Generator().manual_seed(11)
    B, S = 96, 8
    x = (torch.rand(B, 2, generator=g, dtype=torch.float64) * 2 - 1).cuda()  # tied hidden layers need d_in = 2 (= width)
    y = torch.sin(3 * x[:, 0]).contiguous()
    model = dgps.DeepGP(2, x.shape, num_inducing=24).cuda().double()
    mll = dgps.DeepApproximateMLL(dgps.VariationalELBO(model.likelihood, model, 1000))

    def step(rank, world):
        for p in model.parameters():
            p.grad = None
        with dgps.num_likelihood_samples(S), dgps.sample_shard(rank, world):
            out = model(x, seed=77)
            assert out.mean.shape[0] == S // world
            loss = -mll(out, y)
        loss.backward()
        return loss.detach(), torch.cat([p.grad.reshape(-1) for p in model.parameters()])

    l1, g1 = step(0, 1)
    for world in (2, 4):
        parts = [step(r, world) for r in range(world)]
        lw, gw = sum(p[0] for p in parts), sum(p[1] for p in parts)
        assert abs(lw.item() - l1.item()) < 1e-12 * abs(l1.item())
        assert (gw - g1).abs().max().item() < 1e-11 * g1.abs().max().item()
    with pytest.raises(ValueError):
        step(0, 3)
    # the flat-buffer gradient all-reduce helper (identity "collective": x2)
    step(0, 1)
    n = dgps.allreduce_gradients(model, lambda t: t.mul_(2.0))
    assert n == g1.numel()
    assert torch.allclose(torch.cat([p.grad.reshape(-1) for p in model.parameters()]), 2.0 * g1)
