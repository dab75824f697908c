"""Drop-in ``causal_conv1d`` namespace (reference import: selective_scan_interface.py:9, mamba_simple.py:14)."""
from aum_b200.functional import causal_conv1d_fn  # noqa: F401

causal_conv1d_update = None  # single-token decode is outside the AuM hot path

__all__ = ["causal_conv1d_fn", "causal_conv1d_update"]
