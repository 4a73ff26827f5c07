"""Top-level alias of nonstationary_precip_b200.models.latent_priors (the reference imports `models.latent_priors`)."""
from nonstationary_precip_b200.models.latent_priors import *  # noqa: F401,F403
from nonstationary_precip_b200.models import latent_priors as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
