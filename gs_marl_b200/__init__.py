"""gs_marl_b200 — B200-native batched implementation of GS-MARL's env hot path.

Only what the path needs lives here: `csrc/` (sm_100a kernels + the C ABI of
include/gsmarl_b200.h) and the host-side mirror of the reference's env interface
(`environment.py`, `env_wrappers.py`, `scenarios/`).  See DESIGN.md and SPEC.md; the
reference's env sources are withheld, so the model is a *declared* one (parity unpinned).
"""
from .config import WorldConfig  # noqa: F401
from . import scenarios  # noqa: F401

__all__ = ["WorldConfig", "scenarios"]
