"""``from mamba_ssm.ops.triton.layernorm import RMSNorm, layer_norm_fn, rms_norm_fn``
(reference src/models/mamba_models.py:26).  No Triton here: these are the sm_100a CUDA kernel."""
from aum_b200.modules import RMSNorm, layer_norm_fn, rms_norm_fn  # noqa: F401

__all__ = ["RMSNorm", "layer_norm_fn", "rms_norm_fn"]
