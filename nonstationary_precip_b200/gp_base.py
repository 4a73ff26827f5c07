"""Small GP scaffolding used by the mirrored reference classes: the subset of GPyTorch's object model that the
reference's hot path touches, re-implemented on top of the npgp kernels.

It is NOT a GPyTorch re-implementation: only what `models/*.py` of the reference uses is here -- `Module` with
`register_prior` / added loss terms, softplus-constrained parameters, `ConstantMean` / `ZeroMean` / `LinearMean`,
`RBFKernel` / `ScaleKernel` (with batch shapes, as the lengthscale prior needs), `MultivariateNormal` with a dense or a
low-rank-root (+diagonal) covariance, `GaussianLikelihood`, `ExactGP`, `ExactMarginalLogLikelihood`.
Parameter names follow GPyTorch (`raw_outputscale`, `raw_lengthscale`, `raw_noise`, `constant`) so that code such as
experiments/spatial_exp.py:159-186 keeps working."""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import functional as F

LOG2PI = math.log(2.0 * math.pi)


def softplus(x):
    return torch.nn.functional.softplus(x)


def inv_softplus(v: torch.Tensor) -> torch.Tensor:
    return v + torch.log(-torch.expm1(-v))


# ----------------------------------------------------------------------------------------------------------------------
class Module(torch.nn.Module):
    """torch Module + priors and added loss terms (gpytorch.Module semantics used by nonstationary_models.py:35-38 and
    gibbs_kernels.py:256-261)."""

    def __init__(self):
        super().__init__()
        self._priors = {}
        self._added_loss_terms = {}

    def register_parameter(self, name, parameter=None, **kw):  # GPyTorch names the second argument `parameter`
        super().register_parameter(name, parameter)

    def register_prior(self, name, prior, closure, setting_closure=None):
        if isinstance(closure, str):
            attr = closure
            closure = lambda module, _a=attr: getattr(module, _a)  # noqa: E731
        self._priors[name] = (prior, closure)
        if isinstance(prior, torch.nn.Module):
            self.add_module(name, prior)
        else:
            object.__setattr__(self, name, prior)

    def named_priors(self):
        for mod_name, mod in self.named_modules():
            if isinstance(mod, Module):
                for name, (prior, closure) in mod._priors.items():
                    yield (mod_name + "." if mod_name else "") + name, mod, prior, closure

    def update_added_loss_term(self, name, term):
        self._added_loss_terms[name] = term

    def added_loss_terms(self):
        for mod in self.modules():
            if isinstance(mod, Module):
                yield from mod._added_loss_terms.values()


# ----------------------------------------------------------------------------------------------------------------------
# means
# ----------------------------------------------------------------------------------------------------------------------
class ZeroMean(Module):
    def forward(self, x):
        return torch.zeros(x.shape[:-1], dtype=x.dtype, device=x.device)


class ConstantMean(Module):
    """constant of shape (*batch, 1); on (n,d) input returns (*batch, n) -- the origin of the reference's (D,n) layout."""

    def __init__(self, batch_shape=torch.Size([])):
        super().__init__()
        self.constant = torch.nn.Parameter(torch.zeros(tuple(batch_shape) + (1,)))

    def forward(self, x):
        return self.constant.expand(*self.constant.shape[:-1], x.shape[-2])


class LinearMean(Module):
    """x @ W + b with W (d,1), b (1,), both ~ randn (gpytorch.means.LinearMean; models/dgps.py:43)."""

    def __init__(self, input_size):
        super().__init__()
        self.weights = torch.nn.Parameter(torch.randn(input_size, 1))
        self.bias = torch.nn.Parameter(torch.randn(1))

    def forward(self, x):
        return (x * self.weights.reshape(-1)).sum(-1) + self.bias


# ----------------------------------------------------------------------------------------------------------------------
# kernels
# ----------------------------------------------------------------------------------------------------------------------
class Kernel(Module):
    is_stationary = True

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None, **kwargs):
        super().__init__()
        self.ard_num_dims = ard_num_dims
        self._batch_shape = torch.Size(batch_shape)
        if active_dims is not None and not torch.is_tensor(active_dims):
            active_dims = torch.tensor([active_dims] if isinstance(active_dims, int) else list(active_dims),
                                       dtype=torch.long)
        self.register_buffer("active_dims", active_dims)

    @property
    def batch_shape(self):
        return self._batch_shape

    def __call__(self, x1, x2=None, diag=False, **params):
        """Applies active_dims, x2 defaults to x1, and -- like gpytorch.Kernel.__call__ -- takes .diag() of a square
        result when the kernel ignored diag=True."""
        if x1.dim() == 1:
            x1 = x1.unsqueeze(1)
        if x2 is not None and x2.dim() == 1:
            x2 = x2.unsqueeze(1)
        if self.active_dims is not None:
            ad = self.active_dims.to(x1.device)
            x1 = x1.index_select(-1, ad)
            if x2 is not None:
                x2 = x2.index_select(-1, ad)
        if x2 is None:
            x2 = x1
        res = self.forward(x1, x2, diag=diag, **params)
        if diag and torch.is_tensor(res) and res.dim() >= 2 and res.shape[-1] == res.shape[-2] == x1.shape[-2]:
            res = torch.diagonal(res, dim1=-1, dim2=-2)
        return res


class RBFKernel(Kernel):
    has_lengthscale = True

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None, **kwargs):
        super().__init__(ard_num_dims, batch_shape, active_dims)
        nd = 1 if ard_num_dims is None else ard_num_dims
        self.raw_lengthscale = torch.nn.Parameter(torch.zeros(tuple(batch_shape) + (1, nd)))

    @property
    def lengthscale(self):
        return softplus(self.raw_lengthscale)

    @lengthscale.setter
    def lengthscale(self, v):
        v = torch.as_tensor(v, dtype=self.raw_lengthscale.dtype, device=self.raw_lengthscale.device)
        self.raw_lengthscale.data = inv_softplus(v.expand_as(self.raw_lengthscale).clone())

    def forward(self, x1, x2, diag=False, **params):
        ls = self.lengthscale
        if len(self._batch_shape) == 0:
            return F.rbf_ard(x1, x2, ls.reshape(-1))
        return torch.stack([F.rbf_ard(x1, x2, ls[b].reshape(-1)) for b in range(self._batch_shape[0])])


class ScaleKernel(Kernel):
    def __init__(self, base_kernel, batch_shape=torch.Size([]), active_dims=None, outputscale_constraint=None, **kwargs):
        super().__init__(None, batch_shape, active_dims)
        self.base_kernel = base_kernel
        self.raw_outputscale = torch.nn.Parameter(torch.zeros(tuple(batch_shape)))
        self._lower = 0.0 if outputscale_constraint is None else float(outputscale_constraint)

    @property
    def outputscale(self):
        return self._lower + softplus(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, v):
        v = torch.as_tensor(v, dtype=self.raw_outputscale.dtype, device=self.raw_outputscale.device)
        self.raw_outputscale.data = inv_softplus(v.expand_as(self.raw_outputscale).clone() - self._lower)

    def forward(self, x1, x2, diag=False, **params):
        # like gpytorch.ScaleKernel.forward: calls base_kernel.forward directly (skips the base's active_dims)
        k = self.base_kernel.forward(x1, x2, diag=diag, **params)
        os = self.outputscale
        if isinstance(k, LowRankRootCovar):
            return k.scaled(os)
        return k * (os.reshape(*os.shape, 1, 1) if os.dim() else os)


def GreaterThan(lower):  # gpytorch.constraints.GreaterThan(7) -> lower + softplus(raw)
    return float(lower)


# ----------------------------------------------------------------------------------------------------------------------
# covariances and distributions
# ----------------------------------------------------------------------------------------------------------------------
class LowRankRootCovar:
    """R R^T (+ diag): the LowRankRootLazyTensor / LowRankRootAddedDiagLazyTensor of gibbs_kernels.py:225-232."""

    def __init__(self, root, diag: Optional[torch.Tensor] = None):
        self.root, self.added_diag = root, diag

    def scaled(self, s):
        return LowRankRootCovar(self.root * torch.sqrt(s), None if self.added_diag is None else self.added_diag * s)

    def evaluate(self):
        K = F.matmul(self.root, self.root.T)
        return K if self.added_diag is None else K + torch.diag(self.added_diag)

    def diag(self):
        d = (self.root * self.root).sum(-1)
        return d if self.added_diag is None else d + self.added_diag

    @property
    def shape(self):
        n = self.root.shape[0]
        return torch.Size([n, n])


class MultivariateNormal:
    """mean (.., n) and covariance (dense tensor (.., n, n) or LowRankRootCovar), optional iid noise variance."""

    def __init__(self, mean, covariance_matrix, noise: Optional[torch.Tensor] = None):
        self.loc = mean
        self._covar = covariance_matrix
        self._noise = noise

    @property
    def mean(self):
        return self.loc

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def covariance_matrix(self):
        K = self._covar.evaluate() if isinstance(self._covar, LowRankRootCovar) else self._covar
        if self._noise is not None:
            K = K + self._noise * torch.eye(K.shape[-1], dtype=K.dtype, device=K.device)
        return K

    @property
    def variance(self):
        if isinstance(self._covar, LowRankRootCovar):
            v = self._covar.diag()
        else:
            v = torch.diagonal(self._covar, dim1=-1, dim2=-2)
        return v if self._noise is None else v + self._noise

    @property
    def stddev(self):
        return self.variance.sqrt()

    @property
    def event_shape(self):
        return self.loc.shape[-1:]

    @property
    def batch_shape(self):
        return self.loc.shape[:-1]

    def add_noise(self, noise):
        return MultivariateNormal(self.loc, self._covar, noise if self._noise is None else self._noise + noise)

    def log_prob(self, y):
        if self.loc.dim() > 1:  # independent batch of distributions
            return torch.stack([MultivariateNormal(self.loc[b], self._covar[b], self._noise).log_prob(y[b])
                                for b in range(self.loc.shape[0])])
        if isinstance(self._covar, LowRankRootCovar) and self._noise is not None and self._covar.added_diag is None:
            return self._woodbury_log_prob(y)
        return F.mvn_log_prob(y, self.loc, self.covariance_matrix)

    def _woodbury_log_prob(self, y):
        """log N(y | mean, R R^T + s2 I) through the M x M system B = I + R^T R / s2."""
        R, s2 = self._covar.root, self._noise
        n, M = R.shape
        r = y - self.loc
        B = torch.eye(M, dtype=R.dtype, device=R.device) + F.matmul(R.T, R) / s2
        LB, PB = F.psd_safe_chol_inv(B)
        w = F.matmul(PB, F.matmul(R.T, r))
        quad = (r * r).sum() / s2 - (w * w).sum() / (s2 * s2)
        logdet = 2.0 * torch.log(torch.diagonal(LB)).sum() + n * torch.log(s2)
        return -0.5 * (quad + logdet + n * LOG2PI)

    def rsample(self, sample_shape=torch.Size()):
        K = self.covariance_matrix
        if K.dim() == 2:
            L, _ = F.psd_safe_chol_inv(K)
            eps = torch.randn(*sample_shape, K.shape[-1], dtype=K.dtype, device=K.device)
            return self.loc + F.matmul(eps.reshape(-1, K.shape[-1]), L.T).reshape(*sample_shape, K.shape[-1])
        return torch.stack([MultivariateNormal(self.loc[b], K[b]).rsample(sample_shape) for b in range(K.shape[0])],
                           dim=len(sample_shape))


class MultivariateNormalPrior(MultivariateNormal):
    def __init__(self, loc, covariance_matrix):
        super().__init__(loc, covariance_matrix)

    def sample_n(self, n):
        return self.rsample(torch.Size([n]))


# ----------------------------------------------------------------------------------------------------------------------
# likelihood, model base, marginal log likelihood
# ----------------------------------------------------------------------------------------------------------------------
class _NoiseCovar(Module):
    def __init__(self):
        super().__init__()
        self.raw_noise = torch.nn.Parameter(torch.zeros(1))


class GaussianLikelihood(Module):
    """noise = 1e-4 + softplus(raw_noise)  (GreaterThan(1e-4))."""

    def __init__(self):
        super().__init__()
        self.noise_covar = _NoiseCovar()

    @property
    def noise(self):
        return 1e-4 + softplus(self.noise_covar.raw_noise)

    @noise.setter
    def noise(self, v):
        raw = self.noise_covar.raw_noise
        v = torch.as_tensor(v, dtype=raw.dtype, device=raw.device)
        raw.data = inv_softplus(v.expand_as(raw).clone() - 1e-4)

    def forward(self, dist):
        return dist.add_noise(self.noise.reshape(()))

    def expected_log_prob(self, y, mean, variance):
        s2 = self.noise.reshape(())
        return -0.5 * (((y - mean) ** 2 + variance) / s2 + torch.log(s2) + LOG2PI)

    def log_marginal(self, y, mean, variance):
        v = variance + self.noise.reshape(())
        return -0.5 * ((y - mean) ** 2 / v + torch.log(v) + LOG2PI)


class ExactGP(Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        if torch.is_tensor(train_inputs):
            train_inputs = (train_inputs,)
        self.train_inputs = tuple(t.unsqueeze(-1) if t.dim() == 1 else t for t in train_inputs)
        self.train_targets = train_targets
        self.likelihood = likelihood

    def _apply(self, fn, *a, **k):
        self.train_inputs = tuple(fn(t) for t in self.train_inputs)
        self.train_targets = fn(self.train_targets)
        return super()._apply(fn, *a, **k)

    def __call__(self, *inputs, **kwargs):
        return self.forward(*inputs, **kwargs)


class InducingPointKernelAddedLossTerm:
    """-0.5 * sum(prior_diag - q_diag) / noise   (gpytorch.mlls.InducingPointKernelAddedLossTerm, Appendix B.3)."""

    def __init__(self, prior_diag, q_diag, likelihood):
        self.prior_diag, self.q_diag, self.likelihood = prior_diag, q_diag, likelihood

    def loss(self):
        return -0.5 * ((self.prior_diag - self.q_diag) / self.likelihood.noise.reshape(())).sum()


class ExactMarginalLogLikelihood(Module):
    """[log N(y | mu, K + s2 I) + added loss terms + sum of log-priors] / n   (SURVEY Appendix B.3)."""

    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood, self.model = likelihood, model

    def forward(self, output, target):
        res = self.likelihood(output).log_prob(target)
        for term in self.model.added_loss_terms():
            res = res + term.loss()
        for _, module, prior, closure in self.model.named_priors():
            res = res + prior.log_prob(closure(module)).sum()
        return res / target.shape[-1]


# ----------------------------------------------------------------------------------------------------------------------
# stationary kernels of the spatio-temporal model (models/spatio_temporal_models.py:41-44)
# ----------------------------------------------------------------------------------------------------------------------
class PeriodicKernel(Kernel):
    """exp(-2 sin^2(pi |x - x'| / p) / l)  (GPyTorch <= 1.8 convention); 1-D inputs."""

    def __init__(self, active_dims=None, **kwargs):
        super().__init__(None, torch.Size([]), active_dims)
        self.raw_lengthscale = torch.nn.Parameter(torch.zeros(1, 1))
        self.raw_period_length = torch.nn.Parameter(torch.zeros(1, 1))

    @property
    def lengthscale(self):
        return softplus(self.raw_lengthscale)

    @property
    def period_length(self):
        return softplus(self.raw_period_length)

    def forward(self, x1, x2, diag=False, **params):
        one = torch.ones(1, dtype=x1.dtype, device=x1.device)
        hyp = torch.cat([1e150 * one, self.lengthscale.reshape(1), self.period_length.reshape(1), one])
        from . import ops
        return ops.rbf_periodic(x1[:, 0], x2[:, 0], hyp)


class ProductKernel(Kernel):
    """RBFKernel * PeriodicKernel on the same single dimension, evaluated in one fused kernel."""

    def __init__(self, k1, k2):
        super().__init__()
        self.kernels = torch.nn.ModuleList([k1, k2])
        if not (isinstance(k1, RBFKernel) and isinstance(k2, PeriodicKernel)):
            raise NotImplementedError("only RBFKernel * PeriodicKernel is on the hot path")

    def hyper(self, outputscale):
        rbf, per = self.kernels
        return torch.cat([rbf.lengthscale.reshape(-1)[:1], per.lengthscale.reshape(1), per.period_length.reshape(1),
                          outputscale.reshape(1)])

    def forward(self, x1, x2, diag=False, outputscale=None, **params):
        from . import ops
        os = outputscale if outputscale is not None else torch.ones(1, dtype=x1.dtype, device=x1.device)
        return ops.rbf_periodic(x1[:, 0], x2[:, 0], self.hyper(os))


def _kernel_mul(self, other):
    return ProductKernel(self, other)


def _kernel_add(self, other):
    return AdditiveKernel(self, other)


Kernel.__mul__ = _kernel_mul
Kernel.__add__ = _kernel_add


class AdditiveKernel(Kernel):
    """Sum of kernels; low-rank-root summands are merged into one root [R1 R2] (rank M1 + M2)."""

    def __init__(self, *kernels):
        super().__init__()
        self.kernels = torch.nn.ModuleList(kernels)

    def forward(self, x1, x2, diag=False, **params):
        parts = [k(x1, x2, diag=diag, **params) for k in self.kernels]
        return sum_covariances(parts)


def sum_covariances(parts):
    if all(isinstance(p, LowRankRootCovar) for p in parts):
        root = torch.cat([p.root for p in parts], dim=-1)
        diags = [p.added_diag for p in parts if p.added_diag is not None]
        return LowRankRootCovar(root, sum(diags) if diags else None)
    dense = [p.evaluate() if isinstance(p, LowRankRootCovar) else p for p in parts]
    return sum(dense)


class InducingPointKernel(Kernel):
    """gpytorch.kernels.InducingPointKernel restated: root K_xz Kzz^{-1/2}, eval-time diagonal
    correction, training-time added loss term; used for the temporal part (spatio_temporal_models.py:42-44)."""

    def __init__(self, base_kernel, inducing_points, likelihood, active_dims=None):
        super().__init__(active_dims=active_dims)
        self.base_kernel, self.likelihood = base_kernel, likelihood
        if inducing_points.dim() == 1:
            inducing_points = inducing_points.unsqueeze(-1)
        # like upstream: re-wrapped in a new Parameter (aliases the storage of the tensor it was given)
        self.register_parameter("inducing_points", torch.nn.Parameter(inducing_points.data
                                                                     if isinstance(inducing_points, torch.nn.Parameter)
                                                                     else inducing_points))

    def _base(self, a, b):
        return self.base_kernel(a, b)

    def forward(self, x1, x2, diag=False, **params):
        z = self.inducing_points
        Kzz = self._base(z, z)
        _, P = F.psd_safe_chol_inv(Kzz)
        inv_root = P.T
        root1 = F.matmul(self._base(x1, z), inv_root)
        if torch.equal(x1, x2):
            covar = LowRankRootCovar(root1)
            prior_diag = self.base_kernel(x1, x1, diag=True)
            if self.training:
                self.update_added_loss_term("inducing_point_loss_term",
                                            InducingPointKernelAddedLossTerm(prior_diag, covar.diag(), self.likelihood))
            else:
                covar = LowRankRootCovar(root1, (prior_diag - covar.diag()).clamp(0, math.inf))
        else:
            if self.training:
                raise RuntimeError("x1 should equal x2 in training mode")
            covar = F.matmul(root1, F.matmul(self._base(x2, z), inv_root).T)
        if diag:
            return covar.diag() if isinstance(covar, LowRankRootCovar) else torch.diagonal(covar)
        return covar
