"""cpmd_b200 — B200-native (sm_100a) drop-in for CPMD's Gamma-point ``vpsi`` + ``rhoofr`` path.

Hand-written CUDA kernels behind a C ABI (``include/cpb200.h``, ``cpmd_b200/csrc``); this package
is the host-side mirror of the reference's interface for that path.  No CPU fallback.
"""
from .api import CpbError, CpmdContext, Plan, StopGM, leadim  # noqa: F401
from . import lib  # noqa: F401

__all__ = ["Plan", "CpmdContext", "CpbError", "StopGM", "leadim", "lib"]
