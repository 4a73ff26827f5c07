"""Top-level alias of nonstationary_precip_b200.models.nonstationary_models (the reference imports `models.nonstationary_models`)."""
from nonstationary_precip_b200.models.nonstationary_models import *  # noqa: F401,F403
from nonstationary_precip_b200.models import nonstationary_models as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
