"""The oracle (oracle/gibbs_oracle.py) pinned against fixtures produced by executing the reference's own source
(tests/golden/make_golden.py), the known-answer vectors of SURVEY.md Appendix C, and 50-digit mpmath closed forms."""
import mpmath as mp
import pytest
import torch

from oracle import gibbs_oracle as o

torch.set_default_dtype(torch.float64)


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("tag", ["d1", "d2", "d3", "d5"])
def test_gibbs_diag_vs_reference_lines(golden, tag):
    g = golden("gibbs_diag_" + tag)
    assert rel(o.gibbs_diag_K(g["x1"], g["x2"], g["ell1"], g["ell2"]), g["K12"]) < 1e-14
    K11 = o.gibbs_diag_K(g["x1"], g["x1"], g["ell1"], g["ell1"])
    assert rel(K11, g["K11"]) < 1e-14
    assert torch.allclose(torch.diagonal(K11), torch.ones(K11.shape[0]), atol=1e-15)


def test_gibbs_diag_kv1():
    """SURVEY.md Appendix C, KV1."""
    x1 = torch.tensor([[0, 0], [1, 0.5], [-0.25, 0.75]])
    x2 = torch.tensor([[0.2, -0.1], [0.9, 0.4]])
    e1 = torch.tensor([[0.3, 0.5, 0.8], [0.4, 0.6, 0.25]])
    e2 = torch.tensor([[0.7, 0.35], [0.2, 0.45]])
    want = torch.tensor([[6.75755476837800906e-01, 1.40867383537459524e-02],
                         [1.28984556050633470e-01, 9.08294735675667986e-01],
                         [7.13894898363300274e-04, 8.77873422265434550e-02]])
    assert rel(o.gibbs_diag_K(x1, x2, e1, e2), want) < 1e-15


@pytest.mark.parametrize("tag,tol", [("f64", 1e-13), ("f32sigma", 5e-6)])
def test_gibbs_full_vs_reference_lines(golden, tag, tol):
    g = golden("gibbs_full_d2_" + tag)
    S1, S2 = o.sigma_from_H(g["H1"], g["Dm"]), o.sigma_from_H(g["H2"], g["Dm"])
    assert rel(o.gibbs_full_K(g["x1"], g["x2"], S1, S2), g["K12"].double()) < tol
    assert rel(o.gibbs_full_K(g["x1"], g["x1"], S1, S1), g["K11"].double()) < tol


def test_gibbs_full_d3_vs_reference_lines(golden):
    """d = 3 (the dimension of BASELINE config 2): the reference's forward lines run with d = 3 (make_golden.py
    multivariate_d3_case; only its constructor hard-codes 2)."""
    g = golden("gibbs_full_d3_f64")
    S1, S2 = o.sigma_from_H(g["H1"], g["Dm"]), o.sigma_from_H(g["H2"], g["Dm"])
    assert rel(o.gibbs_full_K(g["x1"], g["x2"], S1, S2), g["K12"]) < 1e-13
    assert rel(o.gibbs_full_K(g["x1"], g["x1"], S1, S1), g["K11"]) < 1e-13


def test_lognormal_prior_active_dims_vs_reference_lines(golden):
    """The spatio-temporal model's prior term: log_prob on the full (M,3) inducing points sees columns (0,1) = (time, lon);
    the field interpolation sees the (lon, lat) slice."""
    g = golden("lognormal_prior_active_dims")
    lp_tl = o.lognormal_prior_log_prob(g["Z3"][:, 0:2], g["log_ell"], g["c"], g["os"], g["lam"])
    assert rel(lp_tl, g["log_prob_full_Z"]) < 1e-11
    lp_ll = o.lognormal_prior_log_prob(g["Z3"][:, 1:3], g["log_ell"], g["c"], g["os"], g["lam"])
    assert rel(lp_ll, g["log_prob_lonlat"]) < 1e-11
    assert rel(lp_ll, g["log_prob_full_Z"]) > 1e-2  # the two differ: the quirk is observable
    ell_x = o.field_interp_diag(g["x2"], g["Z3"][:, 1:3], torch.exp(g["log_ell"]), g["c"], g["os"], g["lam"])
    assert rel(ell_x, g["ell_x"]) < 1e-12


def test_gibbs_full_mpmath_kv2():
    """KV2 inputs of SURVEY.md Appendix C evaluated with 50 digits.  (The 'all-fp64' numbers printed in the survey are
    themselves off by ~2e-9; the mpmath values below are the closed form of multivariate_gibbs_kernel.py:98-150.)"""
    mp.mp.dps = 50
    x1 = [[0, 0], [1, 0.5], [-0.25, 0.75]]
    x2 = [[0.2, -0.1], [0.9, 0.4]]
    H1 = [[0.5, -1.0], [1.2, 0.3], [-0.7, 0.9]]
    H2 = [[0.1, 0.8], [-1.1, 0.6]]
    D = [[0.6, 0.0], [0.0, -0.9]]

    def sig(h):
        S = mp.matrix(2, 2)
        for a in range(2):
            for b in range(2):
                u = mp.mpf(h[a]) * mp.mpf(h[b])
                S[a, b] = mp.log(1 + mp.exp(u * u)) + mp.mpf(D[a][b]) ** 2
        return S

    want = torch.zeros(3, 2)
    for i in range(3):
        for j in range(2):
            Si, Sj = sig(H1[i]), sig(H2[j])
            A = (Si + Sj) / 2
            pref = mp.det(Si) ** 0.25 * mp.det(Sj) ** 0.25 * mp.det(A) ** -0.5
            dl = mp.matrix([mp.mpf(x1[i][k]) - mp.mpf(x2[j][k]) for k in range(2)])
            Q = (dl.T * mp.inverse(A + mp.mpf("1e-5") * mp.eye(2)) * dl)[0]
            want[i, j] = float(pref * mp.exp(-Q))
    Dm = torch.tensor(D)
    got = o.gibbs_full_K(torch.tensor(x1), torch.tensor(x2), o.sigma_from_H(torch.tensor(H1), Dm),
                         o.sigma_from_H(torch.tensor(H2), Dm))
    assert rel(got, want) < 1e-14
    assert abs(got[0, 0].item() - 0.920746395513261659) < 1e-15


@pytest.mark.parametrize("tag", ["d2", "d3"])
def test_lognormal_field_vs_reference_lines(golden, tag):
    g = golden("lognormal_field_" + tag)
    ell_x = o.field_interp_diag(g["x"], g["xg"], g["ell_g"], g["c"], g["os"], g["lam"])
    assert rel(ell_x, g["ell_x"]) < 1e-12
    lp = o.lognormal_prior_log_prob(g["xg"], torch.log(g["ell_g"]), g["c"], g["os"], g["lam"])
    assert rel(lp, g["log_prob"]) < 1e-11
    K = o.gibbs_diag_K(g["xg"], g["x"], g["ell_g"], ell_x)
    assert rel(K, g["K_cond"]) < 1e-12


def test_sparse_multivariate_vs_reference_lines(golden):
    g = golden("sparse_multivariate_d2")
    Hx = o.field_interp_H(g["x"], g["Z"], g["H"], g["row_os"], g["row_lam"])
    assert rel(Hx, g["Hx"]) < 1e-9  # the reference forms explicit inverses of K_row + 1e-5 I (cond ~1e5)
    col = torch.eye(2)
    assert rel(o.field_interp_H_kron(g["x"], g["Z"], g["H"], g["row_os"], g["row_lam"], col), g["Hx"]) < 1e-10
    Sz = o.sigma_from_H(g["H"], g["Dm"])
    Sx = o.sigma_from_H(g["Hx"], g["Dm"])
    assert rel(o.gibbs_full_K(g["x"], g["Z"], Sx, Sz), g["Kxz"]) < 1e-13
    assert rel(o.gibbs_full_K(g["Z"], g["Z"], Sz, Sz), g["Kzz"]) < 1e-13
    assert rel(o.gibbs_full_K(g["x"], g["x"], Sx, Sx), g["Kxx"]) < 1e-13
    row = o.rbf_ard_K(g["Z"], g["Z"], g["row_lam"], g["row_os"])
    lp = o.matrix_normal_log_prob(g["H"], row, col)
    assert abs(lp.item() - g["prior_H_log_prob"].item()) < 1e-8 * abs(lp.item())


def test_diag_is_special_case_relationship():
    """diagonal kernel == full kernel with Sigma=diag(l^2) but exponent -Q/2 instead of -Q."""
    g = torch.Generator().manual_seed(1)
    x1, x2 = torch.rand(7, 3, generator=g), torch.rand(5, 3, generator=g)
    e1, e2 = torch.rand(3, 7, generator=g) + 0.2, torch.rand(3, 5, generator=g) + 0.2
    S1, S2 = torch.diag_embed(e1.T ** 2), torch.diag_embed(e2.T ** 2)
    Kd = o.gibbs_diag_K(x1, x2, e1, e2)
    Kf = o.gibbs_full_K(x1 / 2 ** 0.5, x2 / 2 ** 0.5, S1, S2, jitter=0.0)
    assert rel(Kf, Kd) < 1e-13


def test_constant_lengthscale_is_rbf():
    g = torch.Generator().manual_seed(2)
    x1, x2 = torch.rand(9, 2, generator=g), torch.rand(4, 2, generator=g)
    ell = 0.37
    K = o.gibbs_diag_K(x1, x2, torch.full((2, 9), ell), torch.full((2, 4), ell))
    want = o.rbf_ard_K(x1, x2, torch.tensor([ell, ell]))
    assert rel(K, want) < 1e-14
