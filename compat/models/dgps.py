"""Top-level alias of nonstationary_precip_b200.models.dgps (the reference imports `models.dgps`)."""
from nonstationary_precip_b200.models.dgps import *  # noqa: F401,F403
from nonstationary_precip_b200.models import dgps as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
