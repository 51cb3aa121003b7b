"""veomni_b200 — B200-native (sm_100a) kernels for VeOmni's data-parallel hot path.

The CUDA library is loaded lazily by :mod:`veomni_b200._lib`; nothing here falls back to PyTorch
eager code when it is missing.
"""

__version__ = "0.1.0"
