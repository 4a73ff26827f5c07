"""Host-side helpers (utils.dataprep / utils.metrics / utils.metrics2) against outputs of the reference's own functions
on its data fixture (tests/golden/uib_spatial_dataprep.npz, written by tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from nonstationary_precip_b200.utils import dataprep, metrics, metrics2


@pytest.fixture(scope="module")
def gold(golden):
    return golden("uib_spatial_dataprep")


def _t(a, dtype=None):
    t = torch.from_numpy(np.asarray(a))
    return t.to(dtype) if dtype is not None else t


def test_dataprep_matches_reference_outputs(gold, tmp_path):
    csv = tmp_path / "uib_spatial.csv"
    np.savetxt(csv, gold["raw"], delimiter=",", header="lon,lat,tp", comments="", fmt="%.17g")
    data = dataprep.download_data(str(csv))
    assert data.dtype == torch.float32 and torch.equal(data, _t(gold["data_f32"]))
    assert torch.allclose(dataprep.prep_inputs(data), _t(gold["prep_inputs"]), rtol=0, atol=1e-6)
    y_bc, lam = dataprep.prep_outputs(data)
    assert abs(lam - float(gold["boxcox_lambda"])) < 1e-9
    np.testing.assert_allclose(y_bc, gold["boxcox_y"], rtol=1e-6)
    (x_in, (y2, lam2)) = dataprep.box_cox_transform(data)
    assert lam2 == lam and x_in.shape == (394, 2)
    xw, yw, meanx, stdx, meany, stdy = dataprep.whitening_transform(data)
    for got, key in ((xw, "xw"), (yw, "yw"), (meanx, "meanx"), (stdx, "stdx"), (meany, "meany"), (stdy, "stdy")):
        assert torch.allclose(got, _t(gold[key]), rtol=1e-6, atol=1e-6), key
    trx, try_, tex, tey = dataprep.train_test_split(xw, yw, 0.8)
    assert trx.shape == (315, 2) and try_.shape == (315,) and tex.shape == (79, 2)
    assert torch.allclose(trx, _t(gold["split_train_x"]), atol=1e-6) and torch.allclose(tey, _t(gold["split_test_y"]), atol=1e-6)
    assert trx.is_contiguous() and tex.is_contiguous()


def test_metrics_match_reference_outputs(gold):
    y, mu, var, ystd = (_t(gold[k]) for k in ("m_y", "m_mu", "m_var", "m_ystd"))
    pred = torch.distributions.MultivariateNormal(mu, torch.diag(var))
    assert abs(float(metrics.rmse(mu, y, ystd)) - float(gold["rmse1"])) < 1e-12
    assert abs(float(metrics2.rmse(mu, y, ystd)) - float(gold["rmse2"])) < 1e-12
    assert abs(float(metrics.nlpd(pred, y, ystd)) - float(gold["nlpd1"])) < 1e-12
    assert abs(float(metrics2.nlpd(pred, y, ystd)) - float(gold["nlpd1"])) < 1e-12
    assert abs(float(metrics.negative_log_predictive_density(y, mu, var)) - float(gold["nlpd_marg"])) < 1e-12


def test_trainable_param_names(capsys):
    m = torch.nn.Linear(3, 2)
    m.bias.requires_grad = False
    assert metrics.get_trainable_param_names(m) == ["weight"]
    metrics.print_trainable_param_names(m)
    out = capsys.readouterr().out
    assert "weight" in out and "bias" not in out and "Total Trainable Params: 6" in out


def test_kmeans_inducing_points(gold):
    X = _t(gold["x_norm64"])
    Z = dataprep.kmeans_inducing_points(50, X, seed=3)
    assert Z.shape == (50, 2) and Z.dtype == X.dtype and torch.isfinite(Z).all()
    assert torch.equal(Z, dataprep.kmeans_inducing_points(50, X, seed=3))  # deterministic
    assert (Z.min(0).values >= X.min(0).values - 1e-12).all() and (Z.max(0).values <= X.max(0).values + 1e-12).all()
    # fixed point of Lloyd's iteration: every centroid is the mean of the rows assigned to it (scaled space)
    s = X.std(0)
    lab = torch.cdist(X / s, Z / s).argmin(1)
    assert len(torch.unique(lab)) == 50
    for k in range(50):
        assert torch.allclose(X[lab == k].mean(0), Z[k], atol=1e-9)
    # quantisation error decreases against a random subset of the rows
    rand = X[torch.randperm(394, generator=torch.Generator().manual_seed(0))[:50]]
    err = lambda C: torch.cdist(X, C).min(1).values.pow(2).mean()  # noqa: E731
    assert err(Z) < err(rand)
    assert dataprep.kmeans_inducing_points(394, X).shape == (394, 2)
    with pytest.raises(ValueError):
        dataprep.kmeans_inducing_points(500, X)


def test_experiment_drivers_defaults_and_data_without_gpu():
    """The experiment drivers keep the reference scripts' option defaults (experiments/spatial_exp.py:54-81,
    spatio_temporal_exp.py:143-145, deepgp_spatial_bench.py:34-37,66) and load the committed data tables; running them
    needs a GPU and must say so."""
    from experiments import deepgp_spatial_bench as dg, spatial_exp as se, spatio_temporal_exp as ste
    a = se.parse_args([])
    assert (a.n_iter, a.splits, a.lr, a.prior_scale, a.prior_ell, a.prior_mean, a.noise, a.scale, a.train_percent) == (
        5000, 10, 1e-2, 1.0, 1.3, 0.3, 0.011, 0.644, 80.0)
    x, y = se.load_khyber_data()
    assert x.shape == (394, 2) and y.shape == (394,) and x.dtype == torch.float64
    b = ste.parse_args([])
    assert (b.n_iter, b.lr) == (500, 0.015)
    xtr, ytr, xte, yte, meany, stdy = ste.load_train_test()
    assert xtr.shape == (172, 3) and xte.shape == (43, 3) and ytr.shape == (172,)
    c = dg.parse_args([])
    assert (c.num_epochs, c.num_samples, c.num_layers, c.batch_size, c.states) == (400, 3, 4, 315, 10)
    assert dg.load_table().shape == (394, 3)
    if not torch.cuda.is_available():
        for mod in (se, ste, dg):
            with pytest.raises(RuntimeError, match="CUDA"):
                mod.main([])
