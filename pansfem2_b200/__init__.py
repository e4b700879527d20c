"""pansfem2_b200: B200-native SIMP topology-optimisation hot path behind PANSFEM2's API.

The product is the CUDA shared library ``libpansfem2_b200.so`` (C ABI in include/pansfem2_b200.h) plus the
C++ header mirror under ``pansfem2_b200/src``.  This Python package is plumbing for tests and bench.py:
ctypes bindings (``capi``), structured meshers (``mesher``) and the synthetic problems (``problems``).
"""
