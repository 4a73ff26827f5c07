"""Top-level alias of nonstationary_precip_b200.utils.metrics2 (the reference imports `utils.metrics2`)."""
from nonstationary_precip_b200.utils.metrics2 import *  # noqa: F401,F403
from nonstationary_precip_b200.utils import metrics2 as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
