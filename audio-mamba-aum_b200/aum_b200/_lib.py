"""ctypes binding of libaum_b200.so (C ABI declared in include/aum_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, this raises.
PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# AUM_B200_LIB: an experiment build of the same library (csrc/Makefile VARIANT=...); never a different implementation
LIB_PATH = os.environ.get("AUM_B200_LIB") or os.path.join(_HERE, "lib", "libaum_b200.so")

F32, F16, BF16 = 0, 1, 2
ACT_NONE, ACT_SOFTPLUS, ACT_SILU = 0, 1, 2
SCAN_Z_PREGATED = 1


def act_from(kind: int, col0: int) -> int:
    """AUM_ACT_FROM(kind, col0): activation applied to output columns >= col0 only."""
    return kind | (col0 << 8)
GEMM_AUTO, GEMM_TCGEN05, GEMM_SIMT = 0, 1, 2

_DT = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}


class AumError(RuntimeError):
    pass


class ScanDir(C.Structure):
    """struct aum_scan_dir (include/aum_b200.h)."""
    _fields_ = [
        ("u", C.c_void_p), ("ld_u", C.c_int64),
        ("delta", C.c_void_p), ("ld_delta", C.c_int64), ("delta_dtype", C.c_int),
        ("A", C.c_void_p),
        ("Bm", C.c_void_p), ("ld_B", C.c_int64),
        ("Cm", C.c_void_p), ("ld_C", C.c_int64),
        ("bc_dtype", C.c_int),
        ("D", C.c_void_p),
        ("delta_bias", C.c_void_p),
        ("delta_softplus", C.c_int),
        ("last_state", C.c_void_p),
        ("ckpt", C.c_void_p),
    ]


class ScanBwdDir(C.Structure):
    """struct aum_scan_bwd_dir (include/aum_b200.h)."""
    _fields_ = [
        ("u", C.c_void_p), ("ld_u", C.c_int64),
        ("delta", C.c_void_p), ("ld_delta", C.c_int64),
        ("A", C.c_void_p),
        ("BC", C.c_void_p), ("ld_bc", C.c_int64),
        ("D", C.c_void_p),
        ("du", C.c_void_p), ("ld_du", C.c_int64),
        ("ddelta", C.c_void_p), ("ld_dd", C.c_int64),
        ("dA", C.c_void_p),
        ("dD", C.c_void_p),
        ("dBC", C.c_void_p), ("ld_dbc", C.c_int64),
        ("dbc_ws", C.c_void_p),
        ("ckpt", C.c_void_p),
        ("ckpt_valid", C.c_int),
        ("dgrad_dtype", C.c_int),
        ("delta_dtype", C.c_int),
        ("dA_is_dAlog", C.c_int),
    ]


# symbol -> (restype, argtypes); also the list the CPU test checks the .so against the header with
SIGNATURES = {
    "aum_version": (C.c_int, []),
    "aum_last_error": (C.c_char_p, []),
    "aum_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "aum_gemm_tn": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int,
                              C.c_void_p, C.c_int64, C.c_int,
                              C.c_void_p, C.c_int64, C.c_int, C.c_int,
                              C.c_int, C.c_int, C.c_int,
                              C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "aum_gemm_wgrad": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int,
                                 C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "aum_causal_conv1d_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "aum_conv_xproj_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "aum_selective_scan_fwd": (C.c_int, [C.POINTER(ScanDir), C.POINTER(ScanDir), C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_float, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "aum_selective_scan_bwd_workspace_floats": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "aum_selective_scan_bwd_dbc_ws_floats": (C.c_int64, [C.c_int, C.c_int, C.c_int]),
    "aum_selective_scan_bwd": (C.c_int, [C.POINTER(ScanBwdDir), C.POINTER(ScanBwdDir), C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_float, C.c_int, C.c_void_p]),
    "aum_sum_cast_colsum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p,
                                      C.c_int, C.c_int, C.c_void_p]),
    "aum_causal_conv1d_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                        C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "aum_add_rmsnorm_fwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                      C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                      C.c_float, C.c_void_p]),
    "aum_add_rmsnorm_bwd": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "aum_patchify": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "aum_assemble_tokens": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "aum_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_void_p]),
    "aum_adam_step_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_float,
                                    C.c_void_p, C.c_int, C.c_void_p]),
    "aum_transpose": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64,
                                C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None

# bench.py hook: when set to a list, every kernel-launching C-ABI call appends (symbol, start_event, end_event)
# (CUDA events on the launching stream), so that a kernel's share of a step can be computed from like-for-like times.
PROFILE_ALL = None
_NO_LAUNCH = {"aum_version", "aum_last_error", "aum_device_info", "aum_selective_scan_bwd_workspace_floats",
              "aum_selective_scan_bwd_dbc_ws_floats"}


class _Lib:
    pass


def _timed(name, fn):
    def call(*args):
        if PROFILE_ALL is None:
            return fn(*args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        PROFILE_ALL.append((name, e0, e1))
        return rc
    return call


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise AumError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C audio-mamba-aum_b200/csrc`. There is no CPU / PyTorch fallback.")
        L = C.CDLL(LIB_PATH)
        ns = _Lib()
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)     # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
            setattr(ns, name, fn if name in _NO_LAUNCH else _timed(name, fn))
        ns._cdll = L
        _lib = ns
    return _lib


_launches = 0


def check(rc: int, what: str):
    """Every wrapper calls this once per C-ABI call; each successful call enqueued exactly one kernel."""
    global _launches
    if rc != 0:
        msg = lib().aum_last_error()
        raise AumError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")
    _launches += 1


def launch_count() -> int:
    """Number of kernels this process has launched through the C ABI (bench.py's gpu_launches)."""
    return _launches


def dt(t: torch.dtype) -> int:
    try:
        return _DT[t]
    except KeyError:
        raise AumError(f"unsupported dtype {t}") from None


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


import threading

_tls = threading.local()


def stream():
    """The CURRENT stream of the device the op's tensors live on (require_cuda noted it), not of whatever device is
    current: the library switches to the tensors' device for the call (csrc/common.cuh DeviceGuard), as the
    reference's kernels do with a CUDAGuard."""
    return C.c_void_p(torch.cuda.current_stream(getattr(_tls, "device", None)).cuda_stream)


def require_cuda(*tensors):
    """All tensors of one op must be CUDA tensors on ONE device; remembers that device for stream()."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise AumError("aum_b200 ops need CUDA tensors (there is no CPU fallback)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise AumError(f"aum_b200 op got tensors on different devices ({dev} and {t.device})")
    _tls.device = dev
