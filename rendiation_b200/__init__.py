"""rendiation_b200 — B200-native (sm_100a) BVH closest-hit ray traversal behind rendiation's
ray-tracing / space-query API surface.

The product is the C-ABI library ``librdn_rt.so`` (``include/rdn_rt.h``) built from ``csrc/``;
this package is the thin ctypes host mirror used by tests and ``bench.py``.  There is NO CPU
fallback: if the CUDA library is missing, importing :mod:`rendiation_b200.api` raises.
"""
from __future__ import annotations

__all__ = ["api", "scenes", "build"]
