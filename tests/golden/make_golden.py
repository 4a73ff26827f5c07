#!/usr/bin/env python3
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE (read-only at /root/reference) in this container.

GPyTorch is not installable here, so `gpytorch` is replaced by tests/golden/_gpytorch_stub.py (a ~100-line restatement of
the few upstream classes the reference's kernel modules touch).  The arithmetic of the hot path -- the lines cited in
each case below -- is the reference's, run verbatim through its classes' `forward` / `conditional_sample` / `log_prob`.

Run:  python tests/golden/make_golden.py      (only in the build container; the GPU box has no /root/reference)
The resulting fixtures are committed; tests never need /root/reference.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import _gpytorch_stub  # noqa: E402

_gpytorch_stub.install()

import importlib  # noqa: E402

gk = importlib.import_module("models.gibbs_kernels")
lp = importlib.import_module("models.latent_priors")
sys.modules["kernels"] = sys.modules["models"]  # sparse_multivariate_gibbs_kernel.py:11 imports a non-existent package
sys.modules["kernels.latent_priors"] = lp
mgk = importlib.import_module("models.multivariate_gibbs_kernel")
smgk = importlib.import_module("models.sparse_multivariate_gibbs_kernel")


def npz(name, **arrs):
    out = {}
    for k, v in arrs.items():
        out[k] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    np.savez(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: v.shape for k, v in out.items()})


def gibbs_diag_cases():
    """GibbsKernel.forward, models/gibbs_kernels.py:135-162 (explicit ell1, ell2; x1 != x2 and x1 == x2 branches)."""
    torch.set_default_dtype(torch.float64)
    g = torch.Generator().manual_seed(173)
    for tag, (n1, n2, D) in {"d1": (17, 9, 1), "d2": (37, 23, 2), "d3": (64, 50, 3), "d5": (12, 20, 5)}.items():
        x1 = torch.rand(n1, D, generator=g) * 2 - 1
        x2 = torch.rand(n2, D, generator=g) * 2 - 1
        ell1 = torch.exp(0.5 * torch.randn(D, n1, generator=g) - 1.0)
        ell2 = torch.exp(0.5 * torch.randn(D, n2, generator=g) - 1.0)
        kern = gk.GibbsKernel(lengthscale_prior=None)
        K12 = kern.forward(x1, x2, ell1=ell1, ell2=ell2)
        K11 = kern.forward(x1, x1, ell1=ell1)  # torch.equal branch: ell2 = ell1
        npz("gibbs_diag_" + tag, x1=x1, x2=x2, ell1=ell1, ell2=ell2, K12=K12, K11=K11)


def lognormal_field_cases():
    """LogNormalPriorProcess.conditional_sample / log_prob, models/gibbs_kernels.py:80-109, and GibbsKernel.forward with
    ell2=None (conditional branch :151-153)."""
    torch.set_default_dtype(torch.float64)
    g = torch.Generator().manual_seed(174)
    for tag, (n, m, D) in {"d2": (40, 15, 2), "d3": (33, 20, 3)}.items():
        prior = gk.LogNormalPriorProcess(input_dim=D)
        os_ = 0.5 + torch.rand(D, generator=g)
        lam = 0.8 + torch.rand(D, 1, D, generator=g)
        c = torch.log(torch.tensor(0.3)) + 0.1 * torch.randn(D, 1, generator=g)
        prior.covar_module.outputscale = os_
        prior.covar_module.base_kernel.lengthscale = lam
        prior.mean_module.constant.data = c
        xg = torch.rand(m, D, generator=g) * 2 - 1
        x = torch.rand(n, D, generator=g) * 2 - 1
        ell_g = torch.exp(torch.log(torch.tensor(0.3)) + 0.3 * torch.randn(D, m, generator=g))
        with torch.no_grad():
            ell_x = prior.conditional_sample(x, given=(xg, ell_g))
            logp = prior.log_prob((xg, torch.log(ell_g)))
            kern = gk.GibbsKernel(lengthscale_prior=prior)
            K = kern.forward(xg, x, ell1=ell_g)  # ell2 sampled conditionally at x given (xg, ell_g)
        npz("lognormal_field_" + tag, x=x, xg=xg, ell_g=ell_g, c=c.squeeze(-1), os=os_, lam=lam.squeeze(1), ell_x=ell_x,
            log_prob=logp, K_cond=K)


def _bare(cls, H, D, d):
    k = cls.__new__(cls)
    torch.nn.Module.__init__(k)
    k.d = d
    k.H = torch.nn.Parameter(H.clone())
    k.D = torch.nn.Parameter(D.clone())
    return k


def multivariate_cases():
    """MultivariateGibbsKernel.forward, models/multivariate_gibbs_kernel.py:77-150 (same-input branch :79-106 and cross
    branch :108-143), once with default dtype float64 (all-fp64 through the reference's own lines) and once with the
    default float32 as shipped (Sigma silently float32)."""
    for dtname, dt in (("f64", torch.float64), ("f32sigma", torch.float32)):
        torch.set_default_dtype(dt)
        g = torch.Generator().manual_seed(175)
        n1, n2, d = 30, 19, 2
        x1 = (torch.rand(n1, d, generator=g, dtype=torch.float64) * 2 - 1)
        x2 = (torch.rand(n2, d, generator=g, dtype=torch.float64) * 2 - 1)
        H1 = torch.randn(n1, d, generator=g, dtype=torch.float64)
        H2 = torch.randn(n2, d, generator=g, dtype=torch.float64)
        Dm = torch.diag(torch.tensor([0.6, -0.9], dtype=torch.float64))
        k = _bare(mgk.MultivariateGibbsKernel, H1.to(dt), Dm, d)
        k.expectation_conditional_matrix_variate_dist = lambda xs: H2.to(dt)  # conditional mean supplied directly
        import contextlib
        import io
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            K11 = k.forward(x1, x1)
            K12 = k.forward(x1, x2)
        npz("gibbs_full_d2_" + dtname, x1=x1, x2=x2, H1=H1, H2=H2, Dm=Dm, K11=K11, K12=K12)
    torch.set_default_dtype(torch.float64)


def multivariate_d3_case():
    """The same forward lines (models/multivariate_gibbs_kernel.py:77-150) with d = 3, the dimension of BASELINE config 2.
    Only the CONSTRUCTOR of the reference hard-codes d = 2 (:46,54,62); `forward` is dimension-generic, so the kernel
    object is built without the constructor (`_bare`) and every arithmetic line runs verbatim.  |D_kk| >= 0.7 keeps the
    3-D Sigma(h) positive definite."""
    torch.set_default_dtype(torch.float64)
    g = torch.Generator().manual_seed(177)
    n1, n2, d = 28, 17, 3
    x1 = torch.rand(n1, d, generator=g) * 2 - 1
    x2 = torch.rand(n2, d, generator=g) * 2 - 1
    H1 = torch.randn(n1, d, generator=g)
    H2 = torch.randn(n2, d, generator=g)
    Dm = torch.diag(torch.tensor([0.9, -1.1, 0.75]))
    k = _bare(mgk.MultivariateGibbsKernel, H1, Dm, d)
    k.expectation_conditional_matrix_variate_dist = lambda xs: H2
    import contextlib
    import io
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        K11 = k.forward(x1, x1)
        K12 = k.forward(x1, x2)
    npz("gibbs_full_d3_f64", x1=x1, x2=x2, H1=H1, H2=H2, Dm=Dm, K11=K11, K12=K12)


def lognormal_active_dims_case():
    """LogNormalPriorProcess(input_dim=2, active_dims=(0,1)) as experiments/spatio_temporal_exp.py:111 builds it, called
    the two ways SparseSpatioTemporal_Nonstationary calls it: `log_prob` on the FULL (M,3) inducing points
    (models/spatio_temporal_models.py:52-55: the prior's active_dims then select columns 0,1 = time, lon) and
    `conditional_sample` on already-sliced (lon, lat) inputs (models/gibbs_kernels.py:310-316)."""
    torch.set_default_dtype(torch.float64)
    g = torch.Generator().manual_seed(178)
    M, n, D = 14, 21, 2
    prior = gk.LogNormalPriorProcess(input_dim=D, active_dims=(0, 1))
    os_ = 0.5 + torch.rand(D, generator=g)
    lam = 0.8 + torch.rand(D, 1, D, generator=g)
    c = torch.log(torch.tensor(0.3)) + 0.1 * torch.randn(D, 1, generator=g)
    prior.covar_module.outputscale = os_
    prior.covar_module.base_kernel.lengthscale = lam
    prior.mean_module.constant.data = c
    Z3 = torch.randn(M, 3, generator=g)
    x2 = torch.randn(n, 2, generator=g)
    log_ell = torch.log(torch.tensor(0.3)) + 0.3 * torch.randn(D, M, generator=g)
    with torch.no_grad():
        logp_full = prior.log_prob((Z3, log_ell))
        logp_lonlat = prior.log_prob((Z3[:, 1:3], log_ell))
        ell_x = prior.conditional_sample(x2, given=(Z3[:, 1:3], torch.exp(log_ell)))
    npz("lognormal_prior_active_dims", Z3=Z3, x2=x2, log_ell=log_ell, c=c.squeeze(-1), os=os_, lam=lam.squeeze(1),
        log_prob_full_Z=logp_full, log_prob_lonlat=logp_lonlat, ell_x=ell_x)


def sparse_multivariate_cases():
    """SparseMultivariateGibbsKernel: full constructor (:29-65), conditional mean of H (:67-80) and forward on inputs
    whose row count differs from M (:92-100, :113-123); default dtype float64."""
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(176)
    g = torch.Generator().manual_seed(176)
    M, n, d = 12, 25, 2
    Z = torch.rand(M, d, generator=g) * 2 - 1
    x = torch.rand(n, d, generator=g) * 2 - 1
    import contextlib
    import io
    k = smgk.SparseMultivariateGibbsKernel(Z, d, Z.clone())
    k.H.data = k.H.data.double()
    k.D.data = torch.diag(torch.tensor([0.7, 0.5]))
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        Hx = k.expectation_conditional_matrix_variate_dist(x)
        Kxz = k.forward(x, Z)
        Kzz = k.forward(Z, Z)
        Kxx = k.forward(x, x)
        logp = k.H_matrix_prior.log_prob(k.H.data)
    npz("sparse_multivariate_d2", Z=Z, x=x, H=k.H.data, Dm=k.D.data, row_os=k.row_covar_kernel.outputscale,
        row_lam=k.row_covar_kernel.base_kernel.lengthscale.reshape(-1), Hx=Hx, Kxz=Kxz, Kzz=Kzz, Kxx=Kxx,
        prior_H_log_prob=logp)


def dataprep_and_metrics_cases():
    """The reference's utils/dataprep.py:9-52 and utils/metrics.py:37-54 / metrics2.py:36-47 run on its own data fixture
    data/uib_spatial.csv (394 rows: lon, lat, tp), plus the split of experiments/spatial_exp.py:126-141 for i = 0.
    The raw table is stored too: it is the input of BASELINE config 1 and /root/reference does not exist on the GPU box."""
    torch.set_default_dtype(torch.float32)
    dp = importlib.import_module("utils.dataprep")
    m1 = importlib.import_module("utils.metrics")
    m2 = importlib.import_module("utils.metrics2")
    path = os.path.join(REF, "data", "uib_spatial.csv")
    import pandas as pd
    raw = np.asarray(pd.read_csv(path, dtype=np.float64))
    data = dp.download_data(path)
    x_in = dp.prep_inputs(data)
    y_bc, bc_lam = dp.prep_outputs(data)
    xw, yw, meanx, stdx, meany, stdy = dp.whitening_transform(data)
    trx, try_, tex, tey = dp.train_test_split(xw, yw, 0.8)
    # spatial_exp.py:126-141 in float64 (load_khyber_data, :36-40), split i = 0
    x64, y64 = torch.tensor(raw[:, 0:2]), torch.tensor(raw[:, -1])
    sx, mx = torch.std_mean(x64, dim=-2)
    sy, my = torch.std_mean(y64)
    rng = np.random.default_rng(173 + 0)
    idx = np.arange(0, y64.shape[0], 1)
    rng.shuffle(idx)
    # metrics on a fixed synthetic prediction
    g = torch.Generator().manual_seed(7)
    yt = torch.randn(40, generator=g, dtype=torch.float64)
    mu = yt + 0.3 * torch.randn(40, generator=g, dtype=torch.float64)
    var = 0.2 + torch.rand(40, generator=g, dtype=torch.float64)
    ystd = torch.tensor(2.5, dtype=torch.float64)
    pred = torch.distributions.MultivariateNormal(mu, torch.diag(var))
    npz("uib_spatial_dataprep", raw=raw, data_f32=data, prep_inputs=x_in, boxcox_y=y_bc, boxcox_lambda=np.float64(bc_lam),
        xw=xw, yw=yw, meanx=meanx, stdx=stdx, meany=meany, stdy=stdy, split_train_x=trx, split_test_y=tey,
        x_norm64=(x64 - mx) / sx, y_norm64=(y64 - my) / sy, stdy64=sy, shuffle_idx_seed173=idx,
        m_y=yt, m_mu=mu, m_var=var, m_ystd=ystd, rmse1=m1.rmse(mu, yt, ystd), rmse2=m2.rmse(mu, yt, ystd),
        nlpd1=m1.nlpd(pred, yt, ystd), nlpd_marg=m1.negative_log_predictive_density(yt, mu, var))
    torch.set_default_dtype(torch.float64)


def spatio_temporal_data_case():
    """Train/test tensors of experiments/spatio_temporal_exp.py:29-56 (load_uib_data + load_train_test: year 2000, months
    1-4 train = 172 rows, month 5 test = 43 rows; float32 z-scored (time, lon, lat) and tp), obtained by running those
    lines on data/uib_spatio_temporal.csv.  (The script itself cannot be imported: cartopy / pymc3 at module top.)"""
    import pandas as pd
    data = pd.read_csv(os.path.join(REF, "data", "uib_spatio_temporal.csv"))
    data = data[data['time'] < 2001].copy()
    data['month'] = data['time'].rank(method='dense').astype('int')
    train_test_data = data[data['month'] < 6]
    x, y = torch.Tensor(np.array(train_test_data))[:, 1:4], torch.Tensor(np.array(train_test_data)[:, -2])
    with torch.no_grad():
        stdx, meanx = torch.std_mean(x, dim=-2)
        x_norm = (x - meanx) / stdx
        stdy, meany = torch.std_mean(y)
        y_norm = (y - meany) / stdy
    split_idx = len(np.where(train_test_data['month'] < 5)[0])
    npz("uib_spatio_temporal_2000", table=np.array(train_test_data, dtype=np.float64), x_train=x_norm[0:split_idx],
        y_train=y_norm[0:split_idx], x_test=x_norm[split_idx:], y_test=y_norm[split_idx:], meany=meany, stdy=stdy)


if __name__ == "__main__":
    cases = [gibbs_diag_cases, lognormal_field_cases, multivariate_cases, multivariate_d3_case, lognormal_active_dims_case,
             sparse_multivariate_cases, dataprep_and_metrics_cases, spatio_temporal_data_case]
    want = sys.argv[1:]  # optional: names of the cases to (re)generate
    for fn in cases:
        if not want or fn.__name__ in want:
            fn()
