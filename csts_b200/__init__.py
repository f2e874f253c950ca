"""csts_b200 — B200-native (sm_100a) implementation of the CSTS forward/backward hot path.

Layout:
  csrc/        hand-written CUDA kernels + the C-ABI (libcsts_b200.so, declared in include/csts_b200.h)
  _lib.py      ctypes binding of that C-ABI (raises if the library is missing — no fallback)
  kernels.py   tensor-level wrappers (allocation + strides only)
  host/        host-side mirror of the reference's slowfast.models / utils interface for this path
"""
__version__ = "0.1.0"
