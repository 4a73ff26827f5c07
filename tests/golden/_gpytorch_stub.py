"""Minimal stand-in for the parts of GPyTorch (1.5-1.8 API) that the reference's kernel modules touch at import and
call time.  Used ONLY by tests/golden/make_golden.py, in this container, to execute the reference's own source lines
(GPyTorch itself is not installable here: no network).  Semantics restated from upstream:
RBF = exp(-0.5 ||(x-x')/l||^2), Positive constraint = softplus, ConstantMean(batch_shape=(D,)) broadcasts to (D,n)."""
import sys
import types

import torch


class _Lazy:
    def __init__(self, t):
        self._t = t

    def evaluate(self):
        return self._t

    def __add__(self, other):
        return self._t + (other._t if isinstance(other, _Lazy) else other)

    __radd__ = __add__


def _inv_softplus(v):
    return v + torch.log(-torch.expm1(-v))


class Kernel(torch.nn.Module):
    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None, **kwargs):
        super().__init__()
        self.ard_num_dims = ard_num_dims
        self._batch_shape = batch_shape
        self.active_dims = active_dims
        self._priors = {}

    def register_parameter(self, name, parameter=None, **kw):  # GPyTorch names the 2nd argument `parameter`
        super().register_parameter(name, parameter)

    def register_prior(self, name, prior, closure, setting_closure=None):
        self._priors[name] = (prior, closure)
        if isinstance(prior, torch.nn.Module):
            self.add_module(name, prior)
        else:
            object.__setattr__(self, name, prior)

    def __call__(self, x1, x2=None, diag=False, **params):
        # gpytorch.kernels.Kernel.__call__: active_dims are applied HERE (index_select on the last dimension), once, by
        # the outermost kernel that is called; ScaleKernel.forward then calls base_kernel.forward directly.
        if x2 is None:
            x2 = x1
        if self.active_dims is not None:
            idx = torch.as_tensor(self.active_dims, dtype=torch.long).reshape(-1)
            x1, x2 = x1.index_select(-1, idx), x2.index_select(-1, idx)
        return _Lazy(self.forward(x1, x2, diag=diag, **params))


class RBFKernel(Kernel):
    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), **kwargs):
        super().__init__(ard_num_dims=ard_num_dims, batch_shape=batch_shape, **kwargs)  # `lengthscale=` kwarg is swallowed
        nd = 1 if ard_num_dims is None else ard_num_dims
        self.raw_lengthscale = torch.nn.Parameter(torch.zeros(tuple(batch_shape) + (1, nd)))

    @property
    def lengthscale(self):
        return torch.nn.functional.softplus(self.raw_lengthscale)

    @lengthscale.setter
    def lengthscale(self, v):
        v = torch.as_tensor(v, dtype=self.raw_lengthscale.dtype).expand_as(self.raw_lengthscale)
        self.raw_lengthscale.data = _inv_softplus(v.clone())

    def forward(self, x1, x2, diag=False, **params):
        a, b = x1 / self.lengthscale, x2 / self.lengthscale
        d2 = ((a.unsqueeze(-2) - b.unsqueeze(-3)) ** 2).sum(-1)
        return torch.exp(-0.5 * d2)


class ScaleKernel(Kernel):
    def __init__(self, base_kernel, batch_shape=torch.Size([]), **kwargs):
        super().__init__(batch_shape=batch_shape, **kwargs)
        self.base_kernel = base_kernel
        self.raw_outputscale = torch.nn.Parameter(torch.zeros(tuple(batch_shape)))

    @property
    def outputscale(self):
        return torch.nn.functional.softplus(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, v):
        v = torch.as_tensor(v, dtype=self.raw_outputscale.dtype).expand_as(self.raw_outputscale)
        self.raw_outputscale.data = _inv_softplus(v.clone())

    def forward(self, x1, x2, diag=False, **params):
        k = self.base_kernel.forward(x1, x2, diag=diag, **params)
        os = self.outputscale
        return k * (os.view(*os.shape, 1, 1) if os.dim() else os)


class InducingPointKernel(Kernel):
    def __init__(self, base_kernel, inducing_points, likelihood, active_dims=None):
        super().__init__(active_dims=active_dims)
        self.base_kernel, self.likelihood = base_kernel, likelihood
        self.register_parameter("inducing_points", torch.nn.Parameter(inducing_points))


class ConstantMean(torch.nn.Module):
    def __init__(self, batch_shape=torch.Size([])):
        super().__init__()
        self.constant = torch.nn.Parameter(torch.zeros(tuple(batch_shape) + (1,)))

    def forward(self, x):
        return self.constant.expand(*self.constant.shape[:-1], x.shape[-2])


class MultivariateNormal(torch.distributions.MultivariateNormal):
    def __init__(self, mean, covar, **kw):
        covar = covar.evaluate() if isinstance(covar, _Lazy) else covar
        super().__init__(mean, covariance_matrix=covar, validate_args=False)


class MultivariateNormalPrior(torch.distributions.MultivariateNormal):
    def __init__(self, loc, covariance_matrix=None, **kw):
        super().__init__(loc, covariance_matrix=covariance_matrix, validate_args=False)

    def sample_n(self, n):
        return self.sample(torch.Size([n]))


def install():
    g = types.ModuleType("gpytorch")
    for sub in ("kernels", "means", "distributions", "priors", "likelihoods", "lazy", "mlls", "utils", "settings",
                "constraints", "models", "variational"):
        m = types.ModuleType("gpytorch." + sub)
        setattr(g, sub, m)
        sys.modules["gpytorch." + sub] = m
    g.kernels.Kernel, g.kernels.RBFKernel, g.kernels.ScaleKernel = Kernel, RBFKernel, ScaleKernel
    g.kernels.InducingPointKernel = InducingPointKernel
    g.means.ConstantMean = ConstantMean
    g.distributions.MultivariateNormal = MultivariateNormal
    g.priors.MultivariateNormalPrior = MultivariateNormalPrior
    g.likelihoods.Likelihood = torch.nn.Module
    g.__version__ = "stub-1.8"
    sys.modules["gpytorch"] = g
    # modules the reference imports at top level but never uses on this path
    sys.modules["pymc3"] = types.ModuleType("pymc3")
    pt = types.ModuleType("prettytable")
    pt.PrettyTable = object
    sys.modules["prettytable"] = pt
    return g
