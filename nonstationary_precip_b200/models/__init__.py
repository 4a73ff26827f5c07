"""Mirror of the reference's `models` package (same module and class names) on the npgp CUDA kernels."""
