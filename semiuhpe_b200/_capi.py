"""ctypes binding of include/semiuhpe_b200.h.

There is NO CPU fallback: if the shared library is missing or a call is made
without CUDA tensors this module raises.  Loading the library itself needs no
GPU (the CPU test-suite checks that every declared symbol is exported).
"""
import ctypes
import os

import torch

from . import _build

_LIB = None

c_f32p = ctypes.c_void_p
c_vp = ctypes.c_void_p
i64, i32, u64, u32, f32 = ctypes.c_int64, ctypes.c_int32, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_float

# name -> (restype, argtypes); mirrors include/semiuhpe_b200.h line by line
SIGNATURES = {
    "suhpe_abi_version": (ctypes.c_int, []),
    "suhpe_error_string": (ctypes.c_char_p, [ctypes.c_int]),
    "suhpe_set_quadrature_cut_bits": (ctypes.c_int, [ctypes.c_int]),
    "suhpe_proper_svd_f32": (ctypes.c_int, [c_vp, i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "suhpe_fisher_fused_f32": (ctypes.c_int, [c_vp, c_vp, i64, f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "suhpe_fisher_from_s_f32": (ctypes.c_int, [c_vp, i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "suhpe_fisher_ce_f32": (ctypes.c_int, [c_vp, c_vp, i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "suhpe_fisher_ce_with_g1_f32": (ctypes.c_int, [c_vp, c_vp, c_vp, i64, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "suhpe_rotate_adjust_f32": (ctypes.c_int, [c_vp, c_vp, i64, i32, c_vp, c_vp]),
    "suhpe_ema_update_f32": (ctypes.c_int, [c_vp, c_vp, c_vp, i32, f32, f32, i32, c_vp]),
    "suhpe_laplace_nll_f32": (ctypes.c_int, [c_vp, c_vp, i64, c_vp, i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "suhpe_select_init": (ctypes.c_int, [c_vp, u64, c_vp]),
    "suhpe_select_hist_f32": (ctypes.c_int, [c_vp, i64, i32, c_vp, c_vp, c_vp]),
    "suhpe_select_scan": (ctypes.c_int, [c_vp, i32, i32, c_vp, c_vp]),
    "suhpe_entropy_threshold_f32": (ctypes.c_int, [c_vp, i64, u64, c_vp, c_vp, c_vp, c_vp]),
    "suhpe_select_threshold_ptr": (c_vp, [c_vp]),
    "suhpe_select_read": (ctypes.c_int, [c_vp, ctypes.POINTER(f32), ctypes.POINTER(u32), ctypes.POINTER(u64), c_vp]),
    "suhpe_entropy_mask_f32": (ctypes.c_int, [c_vp, i64, c_vp, f32, c_vp, c_vp, c_vp]),
    "suhpe_so3_metrics_f32": (ctypes.c_int, [c_vp, c_vp, c_vp, i64, i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "suhpe_fp32_probe": (ctypes.c_int, [c_vp, i32, i32, i32, c_vp]),
    "suhpe_pipeline_create": (ctypes.c_int, [ctypes.POINTER(c_vp), i64, i64]),
    "suhpe_pipeline_destroy": (ctypes.c_int, [c_vp]),
    "suhpe_fisher_filter_host": (ctypes.c_int, [c_vp, c_vp, c_vp, i64, f32, u64, c_vp, c_vp, c_vp, c_vp,
                                                ctypes.POINTER(f32), ctypes.POINTER(u64)]),
    "suhpe_fisher_pool_host": (ctypes.c_int, [c_vp, c_vp, c_vp, i64, f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "suhpe_pipeline_sync": (ctypes.c_int, [c_vp]),
}

EINVAL = -100000
STATUS_NONFINITE = 1
STATUS_TRACE_RANGE = 2
STATUS_NONFINITE_CE = 4
FISHER_CE_WORKSPACE_FLOATS = 10
HIST_BINS = 2048
SELECT_STATE_BYTES = 32


def lib():
    """The loaded C-ABI library; raises if it was never built (no fallback)."""
    global _LIB
    if _LIB is None:
        path = _build.LIB_PATH
        if not os.path.exists(path):
            raise RuntimeError(
                f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  semiuhpe_b200 has no CPU or eager fallback.")
        handle = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the ABI drifted
            fn.restype, fn.argtypes = res, args
        if handle.suhpe_abi_version() != 1:
            raise RuntimeError("libsemiuhpe_b200.so ABI version mismatch; rebuild")
        _LIB = handle
    return _LIB


class CudaKernelError(RuntimeError):
    pass


def check(code, what):
    if code < 0:
        msg = lib().suhpe_error_string(code).decode()
        if code == EINVAL:
            raise ValueError(f"{what}: {msg}")
        raise CudaKernelError(f"{what}: CUDA error {-code}: {msg}")
    return code


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def as_records(t, name, width=9):
    """Contiguous fp32 CUDA (n,width) view of a (n,width) / (n,3,3) tensor.

    The reference reshapes with ``view``/``reshape`` (fisher_utils.py:15,41,75) and
    works in fp32; non-CUDA input is an error here (no CPU path)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: semiuhpe_b200 has no CPU path "
                           "(the oracle under oracle/ is test infrastructure only)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype}); the reference path is fp32")
    t = t.detach().reshape(-1, width)
    return t if t.is_contiguous() else t.contiguous()
