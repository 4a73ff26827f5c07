"""Top-level alias of nonstationary_precip_b200.utils.config (the reference imports `utils.config`)."""
from nonstationary_precip_b200.utils.config import *  # noqa: F401,F403
from nonstationary_precip_b200.utils import config as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
