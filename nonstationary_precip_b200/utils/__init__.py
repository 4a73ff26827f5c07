"""Mirror of the slice of the reference's `utils` package that the hot path uses."""
