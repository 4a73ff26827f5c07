"""Top-level alias of nonstationary_precip_b200.utils.functional (the reference imports `utils.functional`)."""
from nonstationary_precip_b200.utils.functional import *  # noqa: F401,F403
from nonstationary_precip_b200.utils import functional as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
