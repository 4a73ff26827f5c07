"""Mirror of the reference's models/dgps.py:13-111 (deep GP with doubly-stochastic variational inference) on the npgp
kernels, together with the slice of GPyTorch it drives (restated from SURVEY.md Appendix B.2/B.4/B.5):

  CholeskyVariationalDistribution, whitened VariationalStrategy, DeepGPLayer.__call__ (marginal sampling between
  layers, expansion of the deterministic first layer to S = num_likelihood_samples), VariationalELBO and
  DeepApproximateMLL (mean over the S samples of  sum_i E log p(y_i|f_i)/B - KL/N).

Per layer and output dimension: Kzz = s RBF(Z,Z) + jitter I -> blocked Cholesky + inverse; K(h, Z) for all S*B input rows
in one fused tile kernel (RBF-ARD is the constant-lengthscale case of the diagonal Gibbs kernel, so inputs, lengthscales
and outputscale all get analytic gradients); mean = K u, variance = s + 1e-4 + rowdot(K C, K) on the FP64 tensor pipe;
h' = mean + sqrt(var) eps in the DSVI sampling kernel (Philox, sharding-invariant); the expected log-likelihood is reduced
per sample with warp shuffles."""
from __future__ import annotations

import math

import torch

from .. import functional as F
from .. import ops
from ..gp_base import ConstantMean, ExactGP, GaussianLikelihood, LinearMean, Module, MultivariateNormal, RBFKernel, ScaleKernel

num_output_dims = 2
USE_C_LAYER = True  # route every layer output through npgp_dsvi_layer_fwd / _bwd (False: the kernel-by-kernel composition)


class num_likelihood_samples:
    """gpytorch.settings.num_likelihood_samples (default 10)."""
    _value = 10

    def __init__(self, value):
        self.value_, self.prev = value, None

    @classmethod
    def value(cls):
        return cls._value

    def __enter__(self):
        self.prev = num_likelihood_samples._value
        num_likelihood_samples._value = self.value_

    def __exit__(self, *exc):
        num_likelihood_samples._value = self.prev
        return False


class sample_shard:
    """Multi-GPU DSVI: the S likelihood samples are split over `world` ranks, rank r propagating samples
    [r S/world, (r+1) S/world).  The Philox draws are keyed by the element index in the GLOBAL (S, B, width) tensor, so
    the union of the ranks' draws is exactly the single-rank draw, and `DeepApproximateMLL` weights the local mean by
    1/world: summing loss and gradients over ranks (one all-reduce, `allreduce_gradients`) reproduces the single-rank
    step for every world size that divides S."""
    _rank, _world = 0, 1

    def __init__(self, rank, world):
        self.rank_, self.world_, self.prev = int(rank), int(world), None

    @classmethod
    def local_samples(cls):
        S = num_likelihood_samples.value()
        if S % cls._world:
            raise ValueError("num_likelihood_samples (%d) must be divisible by the number of ranks (%d)" % (S, cls._world))
        return S // cls._world

    def __enter__(self):
        self.prev = (sample_shard._rank, sample_shard._world)
        sample_shard._rank, sample_shard._world = self.rank_, self.world_

    def __exit__(self, *exc):
        sample_shard._rank, sample_shard._world = self.prev
        return False


def _dsvi_draw(dist, eps, seed):
    """Sample the previous layer's marginals; offset = this rank's position in the global sample tensor."""
    return ops.dsvi_sample(dist.mean, dist.variance, eps, seed, sample_shard._rank * dist.mean.numel())


def allreduce_gradients(model, all_reduce):
    """Sum the gradients of all parameters over ranks with ONE collective on a flat buffer (`all_reduce(t)` sums the
    tensor in place, e.g. torch.distributed.all_reduce)."""
    ps = [p for p in model.parameters() if p.grad is not None]
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    all_reduce(flat)
    off = 0
    for p in ps:
        p.grad.copy_(flat[off:off + p.numel()].reshape(p.shape))
        off += p.numel()
    return flat.numel()


class MarginalNormal:
    """Independent normal marginals (mean, variance of shape (..., B[, O])) -- all that DSVI propagates."""

    def __init__(self, mean, variance):
        self.mean, self.variance = mean, variance

    @property
    def loc(self):
        return self.mean


class CholeskyVariationalDistribution(Module):
    """q(u) = N(m, L L^T): m zeros + 1e-3 N(0,1) (GPyTorch's first-call initialisation, Appendix B.2), L = I."""

    def __init__(self, num_inducing_points, batch_shape=torch.Size([])):
        super().__init__()
        b = tuple(batch_shape)
        self.variational_mean = torch.nn.Parameter(1e-3 * torch.randn(*b, num_inducing_points))
        self.chol_variational_covar = torch.nn.Parameter(torch.eye(num_inducing_points).expand(*b, -1, -1).clone())


class VariationalStrategy(Module):
    """Whitened variational strategy (Appendix B.5); `jitter_val` is GPyTorch's variational_cholesky_jitter (1e-6 fp64)."""

    def __init__(self, model, inducing_points, variational_distribution, learn_inducing_locations=True, jitter_val=1e-6):
        super().__init__()
        object.__setattr__(self, "model", model)
        self.inducing_points = torch.nn.Parameter(inducing_points.clone(), requires_grad=learn_inducing_locations)
        self._variational_distribution = variational_distribution
        self.jitter_val = jitter_val

    def kl_divergence(self):
        m = self._variational_distribution.variational_mean
        L = torch.tril(self._variational_distribution.chol_variational_covar)
        M = m.shape[-1]
        d = torch.diagonal(L, dim1=-1, dim2=-2)
        return 0.5 * ((L * L).sum() + (m * m).sum() - m.numel() - torch.log(d * d).sum()) + 0.0 * M


class DeepGPLayer(Module):
    """gpytorch.models.deep_gps.DeepGPLayer restated for marginals."""

    def __init__(self, variational_strategy, input_dims, output_dims):
        super().__init__()
        self.variational_strategy = variational_strategy
        self.input_dims, self.output_dims = input_dims, output_dims

    def _marginals_one_output(self, X, o):
        """Whitened SVGP marginals of output dim `o` at the rows X (n, d_in)."""
        vs = self.variational_strategy
        vd = vs._variational_distribution
        batched = self.output_dims is not None
        Z = vs.inducing_points[o] if batched else vs.inducing_points
        m = vd.variational_mean[o] if batched else vd.variational_mean
        Ls = torch.tril(vd.chol_variational_covar[o] if batched else vd.chol_variational_covar)
        ls = self.covar_module.base_kernel.lengthscale
        os = self.covar_module.outputscale
        ls = (ls[o] if batched else ls).reshape(-1)
        os = (os[o] if batched else os).reshape(())
        M = Z.shape[0]
        mean = var = None
        if (USE_C_LAYER and X.is_cuda and X.dtype == torch.float64 and M % 2 == 0 and ls.numel() == X.shape[1]
                and X.shape[1] <= 6):
            # the whole layer output, forward and analytic backward, as one C call each (csrc/dsvi_layer.cu)
            raw_Ls = vd.chol_variational_covar[o] if batched else vd.chol_variational_covar
            mean, var, info = ops.dsvi_layer(X, Z, ls, os, m, raw_Ls, vs.jitter_val)
            if int(info) != 0:  # psd_safe_cholesky's ladder lives in the composed path below
                mean = var = None
        if mean is not None:
            if isinstance(self.mean_module, LinearMean):
                mean = mean + self.mean_module(X)
            else:
                c = self.mean_module.constant
                mean = mean + (c[o] if batched else c).reshape(())
            return mean, var
        eye = torch.eye(M, dtype=X.dtype, device=X.device)
        Kzz = F.rbf_ard(Z, Z, ls, os) + vs.jitter_val * eye
        _, P = F.psd_safe_chol_inv(Kzz)
        u = F.matmul(P.T, m)
        C = F.matmul(P.T, F.matmul(F.matmul(Ls, Ls.T) - eye, P))
        K = F.rbf_ard(X, Z, ls, os)
        mean = F.matmul(K, u)
        var = (os + 1e-4 + ops.rowquad_sym(K, 0.5 * (C + C.T))).clamp_min(1e-6)
        if isinstance(self.mean_module, LinearMean):
            mean = mean + self.mean_module(X)
        else:
            c = self.mean_module.constant
            mean = mean + (c[o] if batched else c).reshape(())
        return mean, var

    def __call__(self, inputs, are_samples=False, eps=None, seed=0, **kwargs):
        deterministic_inputs = not are_samples
        if isinstance(inputs, MarginalNormal):  # DSVI: sample the previous layer's marginals
            inputs = _dsvi_draw(inputs, eps, seed)
            deterministic_inputs = False
        lead = inputs.shape[:-1]
        X = inputs.reshape(-1, inputs.shape[-1]).contiguous()
        if self.output_dims is None:
            mean, var = self._marginals_one_output(X, None)
            out = MarginalNormal(mean.reshape(lead), var.reshape(lead))
        else:
            ms, vs = zip(*[self._marginals_one_output(X, o) for o in range(self.output_dims)])
            out = MarginalNormal(torch.stack(ms, -1).reshape(*lead, self.output_dims),
                                 torch.stack(vs, -1).reshape(*lead, self.output_dims))
        if deterministic_inputs:  # expand to S likelihood samples (same marginals for every sample)
            S = sample_shard.local_samples()
            out = MarginalNormal(out.mean.unsqueeze(0).expand(S, *out.mean.shape),
                                 out.variance.unsqueeze(0).expand(S, *out.variance.shape))
        return out


class DeepGPHiddenLayer(DeepGPLayer):
    """reference models/dgps.py:15-70."""

    def __init__(self, input_dims, output_dims, num_inducing=250, mean_type="constant"):
        if output_dims is None:
            inducing_points = torch.randn(num_inducing, input_dims)
            batch_shape = torch.Size([])
        else:
            inducing_points = torch.randn(output_dims, num_inducing, input_dims)
            batch_shape = torch.Size([output_dims])
        variational_distribution = CholeskyVariationalDistribution(num_inducing_points=num_inducing,
                                                                   batch_shape=batch_shape)
        variational_strategy = VariationalStrategy(self, inducing_points, variational_distribution,
                                                   learn_inducing_locations=True)
        super().__init__(variational_strategy, input_dims, output_dims)
        if mean_type == "constant":
            self.mean_module = ConstantMean(batch_shape=batch_shape)
        else:
            self.mean_module = LinearMean(input_dims)
        self.covar_module = ScaleKernel(RBFKernel(batch_shape=batch_shape, ard_num_dims=input_dims),
                                        batch_shape=batch_shape, ard_num_dims=None)

    def __call__(self, x, *other_inputs, **kwargs):
        """Concatenation-based skip connections, as in the reference (:53-70)."""
        if len(other_inputs):
            if isinstance(x, MarginalNormal):
                x = _dsvi_draw(x, None, kwargs.get("seed", 0))
            processed = [inp.unsqueeze(0).expand(sample_shard.local_samples(), *inp.shape) for inp in other_inputs]
            x = torch.cat([x] + processed, dim=-1)
        return super().__call__(x, are_samples=bool(len(other_inputs)), **kwargs)


class DeepGP(Module):
    """reference models/dgps.py:72-111: `num_layers` x THE SAME hidden layer object (weights tied, dgps.py:88) followed
    by the last layer; GaussianLikelihood."""

    def __init__(self, num_layers, train_x_shape, num_inducing=250):
        hidden_layer = DeepGPHiddenLayer(input_dims=train_x_shape[-1], output_dims=num_output_dims, mean_type="linear",
                                         num_inducing=num_inducing)
        last_layer = DeepGPHiddenLayer(input_dims=hidden_layer.output_dims, output_dims=None, mean_type="constant",
                                       num_inducing=num_inducing)
        super().__init__()
        self.layers = torch.nn.ModuleList([hidden_layer for _ in range(num_layers)])
        self.last_layer = last_layer
        self.likelihood = GaussianLikelihood()

    def forward(self, inputs, eps=None, seed=0):
        """eps: optional list of N(0,1) draws, one per sampling point (for parity tests); else Philox with `seed`."""
        hidden_rep = inputs
        k = 0
        for layer in self.layers:
            e = None
            if isinstance(hidden_rep, MarginalNormal):
                e, k = (eps[k] if eps is not None else None), k + 1
            hidden_rep = layer(hidden_rep, eps=e, seed=seed + 7919 * k)
        e = eps[k] if (eps is not None and isinstance(hidden_rep, MarginalNormal)) else None
        return self.last_layer(hidden_rep, eps=e, seed=seed + 7919 * (k + 1))

    def __call__(self, inputs, **kw):
        return self.forward(inputs, **kw)

    def variational_layers(self):
        seen = []
        for layer in list(self.layers) + [self.last_layer]:
            if not any(layer is s for s in seen):
                seen.append(layer)
        return seen

    def predict(self, test_loader):
        with torch.no_grad():
            mus, variances, lls = [], [], []
            for x_batch, y_batch in test_loader:
                out = self(x_batch)
                noise = self.likelihood.noise.reshape(())
                preds = MarginalNormal(out.mean, out.variance + noise)
                mus.append(preds.mean)
                variances.append(preds.variance)
                lls.append(self.likelihood.log_marginal(y_batch, out.mean, out.variance))
        return preds, torch.cat(mus, dim=-1), torch.cat(variances, dim=-1), torch.cat(lls, dim=-1)


class VariationalELBO(Module):
    """sum_i E_q log p(y_i | f_i) / B - KL / num_data (Appendix B.4), per likelihood sample."""

    def __init__(self, likelihood, model, num_data):
        super().__init__()
        self.likelihood, self.model, self.num_data = likelihood, model, num_data

    def forward(self, output, target):
        mean, var = output.mean, output.variance
        if mean.dim() == 1:
            mean, var = mean.unsqueeze(0), var.unsqueeze(0)
        ell = ops.gauss_ell_batched(target, mean.contiguous(), var.contiguous(), self.likelihood.noise.reshape(()))
        kl = sum(layer.variational_strategy.kl_divergence() for layer in self.model.variational_layers())
        return ell / target.shape[-1] - kl / self.num_data


class DeepApproximateMLL(Module):
    """Mean over the likelihood samples (leading dimension) of the base objective."""

    def __init__(self, base_mll):
        super().__init__()
        self.base_mll = base_mll

    def forward(self, output, target):
        return self.base_mll(output, target).mean(0) / sample_shard._world


class ExactGPModel(ExactGP):
    """The reference's plain exact-GP baseline (models/dgps.py:113-122): wiring only (out of the accelerated scope)."""

    def __init__(self, train_x, train_y, likelihood, kernel):
        super().__init__(train_x, train_y, likelihood)
        self.mean_module = ConstantMean()
        self.covar_module = kernel

    def forward(self, x):
        return MultivariateNormal(self.mean_module(x), self.covar_module(x))
