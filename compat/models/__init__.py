"""Top-level `models` package of the reference, backed by nonstationary_precip_b200.models (see compat/README.md)."""
