"""B200-native non-stationary (Gibbs) GP hot path: hand-written sm_100a CUDA kernels behind a C ABI (include/npgp.h),
with a Python host side that mirrors the reference's GPyTorch-style interface.  No CPU fallback."""
__version__ = "0.1.0"
