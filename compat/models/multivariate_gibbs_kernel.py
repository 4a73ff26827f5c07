"""Top-level alias of nonstationary_precip_b200.models.multivariate_gibbs_kernel (the reference imports `models.multivariate_gibbs_kernel`)."""
from nonstationary_precip_b200.models.multivariate_gibbs_kernel import *  # noqa: F401,F403
from nonstationary_precip_b200.models import multivariate_gibbs_kernel as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
