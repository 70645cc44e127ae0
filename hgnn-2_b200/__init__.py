"""hgnn-2_b200: B200-native (sm_100a) implementation of the HGNN-2 aggregation hot path.
See DESIGN.md.  Host bookkeeping lives in Python; every arithmetic op on features runs in the
hand-written CUDA kernels of ``csrc/`` behind the C ABI declared in ``include/hgnn_b200.h``."""
__version__ = "0.1.0"
