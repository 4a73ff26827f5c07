"""Top-level alias of nonstationary_precip_b200.utils.dataprep (the reference imports `utils.dataprep`)."""
from nonstationary_precip_b200.utils.dataprep import *  # noqa: F401,F403
from nonstationary_precip_b200.utils import dataprep as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
