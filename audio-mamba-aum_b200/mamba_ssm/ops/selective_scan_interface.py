"""Functional operator API of the hot path, same names / argument order / layouts as the reference's
vim-mamba_ssm/mamba_ssm/ops/selective_scan_interface.py (:77, :606, :616, :627)."""
from aum_b200.functional import (  # noqa: F401
    selective_scan_fn, mamba_inner_fn, bimamba_inner_fn, mamba_inner_fn_no_out_proj,
)

__all__ = ["selective_scan_fn", "mamba_inner_fn", "bimamba_inner_fn", "mamba_inner_fn_no_out_proj"]
