"""Top-level `utils` package of the reference, backed by nonstationary_precip_b200.utils (see compat/README.md)."""
