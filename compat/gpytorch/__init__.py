"""`gpytorch` facade for the reference's scripts on a box without GPyTorch: the subset of the GPyTorch API the non-stationary
GP hot path touches, backed by nonstationary_precip_b200.gp_base and models.dgps (hand-written CUDA behind the C ABI).
Not a general GPyTorch replacement -- anything else raises AttributeError at the point of use.  See compat/README.md."""
import contextlib
import sys
import types

from nonstationary_precip_b200 import gp_base as _b
from nonstationary_precip_b200.models import dgps as _d

__version__ = "npgp-compat (API subset of gpytorch 1.5-1.8)"
Module = _b.Module


def _sub(name, **attrs):
    m = types.ModuleType("gpytorch." + name)
    m.__dict__.update(attrs)
    sys.modules["gpytorch." + name] = m
    return m


kernels = _sub("kernels", Kernel=_b.Kernel, RBFKernel=_b.RBFKernel, ScaleKernel=_b.ScaleKernel, PeriodicKernel=_b.PeriodicKernel,
               ProductKernel=_b.ProductKernel, AdditiveKernel=_b.AdditiveKernel, InducingPointKernel=_b.InducingPointKernel)
means = _sub("means", ZeroMean=_b.ZeroMean, ConstantMean=_b.ConstantMean, LinearMean=_b.LinearMean)
distributions = _sub("distributions", MultivariateNormal=_b.MultivariateNormal)
priors = _sub("priors", MultivariateNormalPrior=_b.MultivariateNormalPrior)
likelihoods = _sub("likelihoods", GaussianLikelihood=_b.GaussianLikelihood, Likelihood=_b.Module)
constraints = _sub("constraints", GreaterThan=_b.GreaterThan)
mlls = _sub("mlls", ExactMarginalLogLikelihood=_b.ExactMarginalLogLikelihood, VariationalELBO=_d.VariationalELBO,
            DeepApproximateMLL=_d.DeepApproximateMLL, AddedLossTerm=_b.InducingPointKernelAddedLossTerm,
            InducingPointKernelAddedLossTerm=_b.InducingPointKernelAddedLossTerm)
variational = _sub("variational", VariationalStrategy=_d.VariationalStrategy,
                   CholeskyVariationalDistribution=_d.CholeskyVariationalDistribution)
_deep = _sub("models.deep_gps", DeepGPLayer=_d.DeepGPLayer, DeepGP=_d.DeepGP)
models = _sub("models", ExactGP=_b.ExactGP, deep_gps=_deep)


def _delazify(x):
    return x.evaluate() if hasattr(x, "evaluate") else x


lazy = _sub("lazy", delazify=_delazify)


@contextlib.contextmanager
def _noop(*args, **kwargs):
    yield


settings = _sub("settings", num_likelihood_samples=_d.num_likelihood_samples, fast_pred_var=_noop, cholesky_jitter=_noop,
                max_cg_iterations=_noop, cg_tolerance=_noop, debug=_noop)
utils = _sub("utils")
