"""fle_b200: B200-native FLE decode + encoded-domain predicate evaluation.

Importable as ``fle_b200`` through the loader at the repo root (the directory name
``impala-avx2-parquet-scanner_b200`` is not a Python identifier)."""
from . import capi  # noqa: F401
from .capi import (ALL, AND, EQ, GE, GT, IN, LE, LT, NONE, OR, FleContext, FleError,  # noqa: F401
                   load_library)
