"""Top-level alias of nonstationary_precip_b200.models.spatio_temporal_models (the reference imports `models.spatio_temporal_models`)."""
from nonstationary_precip_b200.models.spatio_temporal_models import *  # noqa: F401,F403
from nonstationary_precip_b200.models import spatio_temporal_models as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
