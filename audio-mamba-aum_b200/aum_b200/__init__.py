"""aum_b200 — B200 (sm_100a) engine for Audio-Mamba's bidirectional selective-scan hot path.

Host side of the C-ABI library ``lib/libaum_b200.so`` (declared in ``include/aum_b200.h``).
The drop-in surface for the reference lives in the sibling ``mamba_ssm`` / ``causal_conv1d`` packages of this
directory, which mirror the reference's import paths.
"""
from . import _lib, ops, mixer  # noqa: F401
from ._lib import AumError, lib  # noqa: F401

__all__ = ["ops", "mixer", "AumError", "lib"]
