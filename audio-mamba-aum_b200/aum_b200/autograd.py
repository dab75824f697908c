"""Training (backward) entry points.  The backward kernels of the hot path (reference:
BiMambaInnerFn.backward, selective_scan_interface.py:519-603) are not built yet: fail loudly instead of
silently falling back to anything else."""


def mamba_mixer_autograd(module, hidden_states):
    raise NotImplementedError(
        "aum_b200: backward of the Mamba mixer is not implemented yet (forward/inference only). "
        "Wrap the call in torch.no_grad() / torch.inference_mode().")


def rms_norm_autograd(*args, **kwargs):
    raise NotImplementedError(
        "aum_b200: backward of add+RMSNorm is not implemented yet (forward/inference only).")
