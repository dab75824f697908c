"""``from mamba_ssm.modules.mamba_simple import Mamba`` (reference src/models/mamba_models.py:18)."""
from aum_b200.modules import Mamba  # noqa: F401

__all__ = ["Mamba"]
