"""Top-level alias of nonstationary_precip_b200.utils.metrics (the reference imports `utils.metrics`)."""
from nonstationary_precip_b200.utils.metrics import *  # noqa: F401,F403
from nonstationary_precip_b200.utils import metrics as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
