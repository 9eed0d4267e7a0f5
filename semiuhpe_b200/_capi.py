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
    "suhpe_proper_svd_f32": (ctypes.c_int, [c_vp, i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "suhpe_proper_svd_backward_f32": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, i64, c_vp, c_vp]),
    "suhpe_fisher_fused_f32": (ctypes.c_int, [c_vp, c_vp, i64, f32, i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "suhpe_fisher_from_s_f32": (ctypes.c_int, [c_vp, i64, i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "suhpe_fisher_ce_f32": (ctypes.c_int, [c_vp, c_vp, i64, i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "suhpe_fisher_ce_with_g1_f32": (ctypes.c_int, [c_vp, c_vp, c_vp, i64, i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "suhpe_scale_rows_f32": (ctypes.c_int, [c_vp, i64, i32, c_vp, c_vp, c_vp, c_vp, c_vp]),
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
    "suhpe_pipeline_create": (ctypes.c_int, [ctypes.POINTER(c_vp), i64, i64, i32]),
    "suhpe_pipeline_destroy": (ctypes.c_int, [c_vp]),
    "suhpe_fisher_filter_host": (ctypes.c_int, [c_vp, c_vp, c_vp, i64, f32, u64, c_vp, c_vp, c_vp, c_vp,
                                                ctypes.POINTER(f32), ctypes.POINTER(u64)]),
    "suhpe_fisher_pool_host": (ctypes.c_int, [c_vp, c_vp, c_vp, i64, f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "suhpe_pipeline_sync": (ctypes.c_int, [c_vp]),
    "suhpe_ssl_step_create": (ctypes.c_int, [ctypes.POINTER(c_vp)]),
    "suhpe_ssl_step_destroy": (ctypes.c_int, [c_vp]),
    "suhpe_ssl_step_workspace_floats": (i64, [i64, i64]),
    "suhpe_ssl_step_f32": (ctypes.c_int, [c_vp, c_vp, c_vp, i64, c_vp, c_vp, i64, c_vp, i32, c_vp, f32, f32, f32, i32, i32,
                                          c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
}
# include/semiuhpe_b200_probe.h (measurement entry points, not part of the drop-in boundary)
PROBE_SIGNATURES = {
    "suhpe_fp32_probe": (ctypes.c_int, [c_vp, i32, i32, i32, c_vp]),
}

ABI_VERSION = 2
CUT_BITS_DEFAULT = 26
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
        for name, (res, args) in list(SIGNATURES.items()) + list(PROBE_SIGNATURES.items()):
            fn = getattr(handle, name)          # AttributeError if the ABI drifted
            fn.restype, fn.argtypes = res, args
        if handle.suhpe_abi_version() != ABI_VERSION:
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
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream(device_index=None):
    """The current CUDA stream of ``device_index`` (default: the current device) as an integer handle."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device() if device_index is None else device_index)
    return torch.cuda.current_stream(device_index).cuda_stream


class on_device:
    """``with on_device(dev):`` -- make ``dev`` current for the launches inside; free when it already is
    (``torch.cuda.device`` costs several microseconds even then, which a 20 us kernel notices)."""
    __slots__ = ("idx", "prev")

    def __init__(self, device):
        self.idx = device.index if device.index is not None else torch.cuda.current_device()
        self.prev = -1

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.idx:
            self.prev = cur
            torch.cuda.set_device(self.idx)
        return self.idx

    def __exit__(self, *exc):
        if self.prev >= 0:
            torch.cuda.set_device(self.prev)
        return False


def as_records(t, name, width=9):
    """Contiguous fp32 CUDA tensor whose memory is n records of ``width`` floats and whose ``shape[0]`` is n
    (the input itself when it already is (n,width) / (n,3,3) and contiguous, else a detached reshaped copy).

    The reference reshapes with ``view``/``reshape`` (fisher_utils.py:15,41,75) and
    works in fp32; non-CUDA input is an error here (no CPU path)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: semiuhpe_b200 has no CPU path "
                           "(the oracle under oracle/ is test infrastructure only)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype}); the reference path is fp32")
    shape = t.shape
    ok = (len(shape) >= 1 and shape[-1] == width) or (width == 9 and len(shape) >= 2 and shape[-1] == 3 and shape[-2] == 3)
    if not ok:
        # the reference's own view(-1, 3, 3) / reshape(-1, 9) accepts any element count divisible by 9; a (n,3)
        # tensor silently reinterpreted as n/3 matrices is far more likely a bug than an intent
        raise RuntimeError(f"{name}: expected trailing dimensions ({width},)" + (" or (3, 3)" if width == 9 else "")
                           + f", got shape {tuple(shape)}")
    if t.is_contiguous() and 2 <= len(shape) <= (3 if width == 9 else 2):
        return t                              # (n,width) / (n,3,3) as it is: callers only take its pointer and shape[0]
    t = t.detach().reshape(-1, width)
    return t if t.is_contiguous() else t.contiguous()
