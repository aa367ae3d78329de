"""B200-native HybridFusion hot path (sm_100a CUDA kernels behind a C ABI).

Host side of the drop-in for ``src/fusion.py`` / ``src/attention.py`` /
``src/encoders.py`` / ``src/uncertainty.py`` of the reference; the arithmetic
lives in ``libmsf_b200.so`` (``csrc/``, declared in ``include/msf_b200.h``).
There is no CPU or PyTorch-eager fallback: without the library or without a
CUDA device every compute entry point raises.
"""
from . import _native as native  # noqa: F401
from ._native import MsfError, lib, library_path  # noqa: F401

__all__ = ["native", "lib", "library_path", "MsfError"]
