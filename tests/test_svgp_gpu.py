"""SVGP-Gibbs ELBO step on the CUDA kernels (through the C ABI) against the CPU oracle: loss, every gradient,
prediction, a short training trace, and full-size (B=65536, M=1024) size-independent properties."""
import pytest
import torch

from svgp_cases import make_problem, oracle_loss_and_grads

pytestmark = pytest.mark.gpu
torch.set_default_dtype(torch.float64)


def rel(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def build(variant, **kw):
    from nonstationary_precip_b200.svgp import SVGPGibbs
    x, y, Z, p, N = make_problem(variant, device="cuda", **kw)
    return SVGPGibbs(variant, Z, N, **p), x, y, Z, p, N


@pytest.mark.parametrize("variant,d,B,M", [("diag", 3, 700, 160), ("diag", 2, 300, 64), ("full", 3, 700, 160),
                                            ("full", 2, 513, 130)])
def test_elbo_and_gradients_match_oracle(variant, d, B, M):
    model, x, y, Z, p, N = build(variant, B=B, M=M, d=d, seed=1)
    loss = model.loss_and_grad(x, y)
    assert int(model.last["info"]) == 0
    want_loss, want = oracle_loss_and_grads(variant, x, y, Z, p, N)
    # north_star tolerance: ELBO within 1e-6 relative
    assert abs(loss.item() - want_loss.item()) < 1e-9 * abs(want_loss.item())
    for name, gw in want.items():
        assert rel(model.g[name], gw) < 1e-6, name


@pytest.mark.parametrize("variant", ["diag", "full"])
def test_predict_matches_oracle(variant):
    from oracle import gibbs_oracle as o
    from nonstationary_precip_b200.svgp import _inv_softplus
    model, x, y, Z, p, N = build(variant, B=400, M=96, d=3, seed=2)
    xs = torch.rand(1000, 3, device="cuda") * 2 - 1
    mu, var = model.predict(xs, chunk=384)
    c = lambda t: t.detach().cpu()
    if variant == "diag":
        extra = dict(log_ell_z=c(p["log_ell_z"]), prior_c=c(p["prior_c"]), prior_os=c(p["prior_os"]),
                     prior_lam=c(p["prior_lam"]))
    else:
        extra = dict(H=c(p["H"]), Dm=c(p["Dm"]), row_os=torch.tensor(float(p["row_os"])), row_lam=c(p["row_lam"]))
    mu_w, var_w = o.svgp_gibbs_predict(c(xs), c(Z), c(p["m"]), c(p["Ls"]), torch.tensor(_inv_softplus(p["outputscale"])),
                                       variant, **extra)
    # north_star tolerance: predictive mean / variance within 1e-6 relative
    assert rel(mu, mu_w) < 1e-8
    assert rel(var, var_w) < 1e-8


def test_training_trace_matches_oracle_adam():
    """First Adam steps: CUDA analytic-gradient path vs torch.optim.Adam on the oracle's autograd (data
    fixture row: 'loss trace of the first k Adam steps')."""
    from oracle import gibbs_oracle as o
    from nonstationary_precip_b200.svgp import _inv_softplus
    variant = "diag"
    model, x, y, Z, p, N = build(variant, B=256, M=48, d=3, seed=4)
    c = lambda t: t.detach().cpu().clone()
    P = dict(Z=c(Z), log_ell_z=c(p["log_ell_z"]), m=c(p["m"]), Ls=c(p["Ls"]),
             ro=torch.tensor(_inv_softplus(p["outputscale"])), rn=torch.tensor(_inv_softplus(p["noise"] - 1e-4)))
    for v in P.values():
        v.requires_grad_(True)
    opt = torch.optim.Adam(list(P.values()), lr=0.01)
    xc, yc = c(x), c(y)
    got, want = [], []
    for _ in range(5):
        got.append(model.train_step(x, y, lr=0.01).item())
        opt.zero_grad()
        loss = -o.svgp_gibbs_elbo(xc, yc, N, P["Z"], P["m"], P["Ls"], P["ro"], P["rn"], variant,
                                  log_ell_z=P["log_ell_z"], prior_c=c(p["prior_c"]), prior_os=c(p["prior_os"]),
                                  prior_lam=c(p["prior_lam"]))
        loss.backward()
        P["Ls"].grad = torch.tril(P["Ls"].grad)
        opt.step()
        want.append(loss.item())
    for a, b in zip(got, want):
        assert abs(a - b) < 1e-6 * abs(b)


@pytest.mark.parametrize("variant", ["full", "diag"])
def test_full_size_properties(variant):
    """BASELINE config 2 shapes (B=65536, M=1024, d=3): properties that do not need the (slow) oracle at this size."""
    model, x, y, Z, p, N = build(variant, B=65536, M=1024, d=3, seed=9, N_total=1 << 20)
    loss = model.loss_and_grad(x, y)
    assert int(model.last["info"]) == 0 and torch.isfinite(loss)
    assert torch.isfinite(model.grad).all()
    g_full = model.grad.clone()
    # (1) linearity over rows: gradients of two half batches (world_size=2 weighting) sum to the full-batch gradient
    tot = torch.zeros_like(g_full)
    for r in range(2):
        model.loss_and_grad(x[r::2].contiguous(), y[r::2].contiguous(), world_size=2)
        tot += model.grad
    assert rel(tot, g_full) < 1e-7  # summation order x conditioning of the C = P^T (S-I) P form
    # (2) K(Z,Z) has a unit diagonal (outputscale 1) and is symmetric
    fz, _, _ = model._field_forward(x[:64].contiguous())
    Kzz = model._kernel_fwd(model.p["Z"], fz, model.p["Z"], fz, None)
    assert (torch.diagonal(Kzz) - 1).abs().max() < 1e-12
    assert (Kzz - Kzz.T).abs().max() < 1e-12
    # (3) with S = I and m = 0 the whitened posterior equals the prior: mean 0, variance s + jitter
    model.p["m"].zero_()
    model.p["Ls"].copy_(torch.eye(1024))
    mu, var = model.predict(x[:4096].contiguous())
    s = torch.nn.functional.softplus(model.p["raw_outputscale"])
    assert mu.abs().max() < 1e-12
    assert (var - (s + 1e-4)).abs().max() < 1e-9


def test_failed_cholesky_is_flagged_and_leaves_parameters_untouched():
    """Reference behaviour: psd_safe_cholesky (models/gibbs_kernels.py:201) retries with jitter 1e-8, 1e-7, 1e-6 and then
    raises.  A replayed graph cannot raise: the step sets a sticky device flag, the guarded Adam skips the update, the host
    polls the flag and climbs the same ladder."""
    from nonstationary_precip_b200.svgp import SVGPGibbs
    x, y, Z, p, N = make_problem("diag", device="cuda", B=512, M=128, d=3, seed=3)
    # (a) an indefinite Kzz that no jitter of the ladder repairs
    bad = SVGPGibbs("diag", Z, N, jitter_zz=-1.0, **p)
    before = bad.theta.clone()
    for _ in range(2):
        bad.train_step(x, y, lr=0.01)
    assert bad.check_status() & 1
    assert torch.equal(bad.theta, before) and float(bad.step_dev) == 0.0  # nothing was updated, not even to NaN
    for _ in range(3):
        bad.recover()
        bad.train_step(x, y, lr=0.01)
        assert bad.check_status() & 1
    with pytest.raises(RuntimeError):
        bad.recover()
    # (b) a marginally indefinite Kzz (two coincident inducing points, pivot -1e-9): the first rung of the ladder repairs it
    Z2 = Z.clone()
    Z2[1] = Z2[0]
    p2 = dict(p)
    p2["log_ell_z"] = p["log_ell_z"].clone()
    p2["log_ell_z"][:, 1] = p2["log_ell_z"][:, 0]
    m = SVGPGibbs("diag", Z2, N, jitter_zz=-1e-9, learn_inducing_locations=False, **p2)
    m.capture(512, 1, 512, lr=0.01)
    before = m.theta.clone()
    m.train_step_graph(x, y)
    assert m.check_status() & 1 and torch.equal(m.theta, before)
    assert m.recover() is True  # the graph has to be re-captured: the jitter is baked into it
    m.capture(512, 1, 512, lr=0.01)
    loss = m.train_step_graph(x, y)
    assert m.check_status() == 0 and torch.isfinite(loss) and not torch.equal(m.theta, before)
