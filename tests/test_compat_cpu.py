"""compat/: the top-level `models`, `utils`, `gpytorch`, `pymc3` names the reference's experiment drivers import resolve to
the B200-native mirrors.  Import-only: no GPU is touched."""
import ast
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

# what /root/reference/experiments/{spatial_exp,spatio_temporal_exp,deepgp_spatial_bench}.py import from these packages
# (spatial_exp.py:21-27, spatio_temporal_exp.py:18-24, deepgp_spatial_bench.py:10-20)
WANTED = {
    "gpytorch.kernels": ["ScaleKernel", "RBFKernel", "PeriodicKernel", "InducingPointKernel"],
    "gpytorch.constraints": ["GreaterThan"],
    "gpytorch.mlls": ["VariationalELBO", "AddedLossTerm", "DeepApproximateMLL", "ExactMarginalLogLikelihood"],
    "gpytorch.models": ["ExactGP"],
    "gpytorch.likelihoods": ["GaussianLikelihood"],
    "gpytorch.distributions": ["MultivariateNormal"],
    "models.gibbs_kernels": ["LogNormalPriorProcess", "PositivePriorProcess", "GibbsKernel", "GibbsSafeScaleKernel",
                             "InducingGibbsKernel", "InducingGibbsKernelST"],
    "models.nonstationary_models": ["DiagonalSparseGP", "DiagonalExactGP"],
    "models.spatio_temporal_models": ["SparseSpatioTemporal_Nonstationary"],
    "models.dgps": ["DeepGPHiddenLayer", "DeepGP"],
    "models.multivariate_gibbs_kernel": ["MultivariateGibbsKernel"],
    "models.sparse_multivariate_gibbs_kernel": ["SparseMultivariateGibbsKernel"],
    "models.latent_priors": ["MatrixVariateNormalPrior"],
    "utils.config": ["BASE_SEED", "EPSILON", "DATASET_DIR", "RESULTS_DIR"],
    "utils.metrics": ["rmse", "nlpd", "negative_log_predictive_density", "get_trainable_param_names"],
    "utils.metrics2": ["nlpd", "rmse"],
    "utils.dataprep": ["download_data", "prep_inputs", "prep_outputs", "whitening_transform", "train_test_split"],
    "utils.functional": ["op", "dot", "mv", "t"],
    "pymc3.gp.util": ["kmeans_inducing_points"],
}


def _run(code):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "compat"),
                                                       os.path.join(ROOT, "compat", "optional_stubs")]))
    return subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)


def test_top_level_names_resolve_to_the_mirrors():
    code = "import importlib, json, sys\nwanted = %r\nmissing = []\n" % WANTED + (
        "for mod, names in wanted.items():\n"
        "    m = importlib.import_module(mod)\n"
        "    missing += [mod + '.' + n for n in names if not hasattr(m, n)]\n"
        "import models.gibbs_kernels as g, nonstationary_precip_b200.models.gibbs_kernels as h\n"
        "assert g.GibbsKernel is h.GibbsKernel\n"
        "import gpytorch\nassert issubclass(g.GibbsKernel, gpytorch.kernels.Kernel)\n"
        "print(json.dumps(missing))\n")
    r = _run(code)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().splitlines()[-1] == "[]", r.stdout


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree exists only in the build container")
@pytest.mark.parametrize("script", ["spatial_exp.py", "spatio_temporal_exp.py", "deepgp_spatial_bench.py"])
def test_reference_driver_import_block_executes_unchanged(script):
    """The import statements of the reference's own driver, taken verbatim from its source, execute against compat/."""
    src = open(os.path.join(REF, "experiments", script)).read()
    tree = ast.parse(src)
    imports = [ast.get_source_segment(src, n) for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert len(imports) >= 10
    skip = ("tqdm",)  # progress bar: present here, irrelevant
    code = "\n".join(i for i in imports if not any(s in i for s in skip)) + "\nprint('ok')\n"
    r = _run(code)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-3000:]
