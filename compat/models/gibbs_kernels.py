"""Top-level alias of nonstationary_precip_b200.models.gibbs_kernels (the reference imports `models.gibbs_kernels`)."""
from nonstationary_precip_b200.models.gibbs_kernels import *  # noqa: F401,F403
from nonstationary_precip_b200.models import gibbs_kernels as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
