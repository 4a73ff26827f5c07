"""`pymc3` stand-in: the reference's experiment drivers use exactly one function of it,
pm.gp.util.kmeans_inducing_points (experiments/spatial_exp.py:153), for inducing-point placement."""
from . import gp  # noqa: F401
