"""mytrim_b200 — B200-native cascade transport for MyTRIM's TrimBase::trim() hot path.

The product is ``libmytrim_b200.so`` (hand-written sm_100a kernels behind the C ABI of
``include/mytrim_b200.h``) plus the C++ plugin façade in ``include/mytrim``.  This package is the
thin ctypes harness the tests and ``bench.py`` use.
"""
from . import capi  # noqa: F401
from .capi import Engine, MytrimError, default_config, make_ions  # noqa: F401
