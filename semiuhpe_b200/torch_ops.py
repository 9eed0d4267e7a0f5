"""``torch.library`` registration of the hot-path ops (namespace ``semiuhpe_b200``).

The mirrors call the C ABI directly through ``ctypes`` in eager mode -- the shortest host path, which is what a
32-sample training step needs.  That path is opaque to ``torch.compile`` / FX tracing.  The ops below are the same
launches registered as dispatcher ops -- CUDA implementation, fake-tensor (shape) implementation and autograd
formula each -- so that a traced or compiled training step sees them as single nodes instead of graph breaks:

    torch.ops.semiuhpe_b200.fisher_nll(A, R, overreg, want_rot, want_grad, keep) -> (nll, rot, grad)
    torch.ops.semiuhpe_b200.fisher_entropy(A) -> entropy
    torch.ops.semiuhpe_b200.proper_rotation(A) -> R          (.._full(A) -> (R, U, V, S), differentiable; polar_backward)
    torch.ops.semiuhpe_b200.fisher_ce(A1, A2, want_grad, target_G, keep) -> (ce, grad)
    torch.ops.semiuhpe_b200.laplace_nll(pred, gt, grids, want_grad, keep) -> (nll, mode, grad)
    torch.ops.semiuhpe_b200.geodesic_deg(pred, gt) -> degrees
    torch.ops.semiuhpe_b200.scale_rows(rows, weight, keep) -> rows * weight[:, None]   (zero rows where keep is false)

The mirrors (``fisher_utils.vmf_loss`` ...) route through these automatically while ``torch.compiler.is_compiling()``.
``want_grad`` is decided by the caller (``A.requires_grad and torch.is_grad_enabled()``): the per-sample gradient is
an output of the forward launch, saved for the backward, which is one ``scale_rows`` launch.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _ops

_NS = "semiuhpe_b200"


def _rows(t):
    return t.numel() // 9


def _empty_rows(like, n, *tail):
    return like.new_empty((n,) + tail, dtype=torch.float32)


# --------------------------------------------------------------------------- scale_rows
@torch.library.custom_op(f"{_NS}::scale_rows", mutates_args=(), device_types="cuda")
def scale_rows(rows: Tensor, weight: Tensor, keep: Optional[Tensor] = None) -> Tensor:
    return _ops.scale_rows(rows, weight, keep=keep)


@scale_rows.register_fake
def _(rows, weight, keep=None):
    return torch.empty_like(rows, memory_format=torch.contiguous_format)


# --------------------------------------------------------------------------- matrix Fisher
@torch.library.custom_op(f"{_NS}::fisher_nll", mutates_args=(), device_types="cuda")
def fisher_nll(A: Tensor, R: Tensor, overreg: float, want_rot: bool, want_grad: bool,
               keep: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    out = _ops.fisher_fused(A, R, overreg, nll=True, grad=want_grad, rot=want_rot, what="KL_Fisher")
    n = out["nll"].shape[0]
    rot = out["rot"] if want_rot else _empty_rows(out["nll"], 0, 3, 3)
    grad = out["grad"] if want_grad else _empty_rows(out["nll"], 0, 9)
    return out["nll"], rot, grad


@fisher_nll.register_fake
def _(A, R, overreg, want_rot, want_grad, keep=None):
    n = _rows(A)
    return _empty_rows(A, n), _empty_rows(A, n if want_rot else 0, 3, 3), _empty_rows(A, n if want_grad else 0, 9)


def _fisher_nll_setup(ctx, inputs, output):
    A, _R, _overreg, _want_rot, want_grad, keep = inputs
    ctx.a_shape = A.shape
    ctx.have_grad = want_grad
    ctx.save_for_backward(output[2], keep)


def _fisher_nll_backward(ctx, g_nll, _g_rot, _g_grad):
    grad, keep = ctx.saved_tensors
    if not ctx.have_grad:
        raise RuntimeError("semiuhpe_b200::fisher_nll was called with want_grad=False but its input requires grad")
    return torch.ops.semiuhpe_b200.scale_rows(grad, g_nll, keep).view(ctx.a_shape), None, None, None, None, None


fisher_nll.register_autograd(_fisher_nll_backward, setup_context=_fisher_nll_setup)


@torch.library.custom_op(f"{_NS}::fisher_entropy", mutates_args=(), device_types="cuda")
def fisher_entropy(A: Tensor) -> Tensor:
    return _ops.fisher_fused(A, None, 1.0, entropy=True, what="fisher_entropy")["entropy"]


@fisher_entropy.register_fake
def _(A):
    return _empty_rows(A, _rows(A))


@torch.library.custom_op(f"{_NS}::proper_rotation", mutates_args=(), device_types="cuda")
def proper_rotation(A: Tensor) -> Tensor:
    return _ops.proper_svd(A, rot=True, what="batch_torch_A_to_R")["rot"]


@proper_rotation.register_fake
def _(A):
    return _empty_rows(A, _rows(A), 3, 3)


@torch.library.custom_op(f"{_NS}::polar_backward", mutates_args=(), device_types="cuda")
def polar_backward(U: Tensor, V: Tensor, S: Tensor, grad_R: Tensor) -> Tensor:
    return _ops.proper_svd_backward(U, V, S, grad_R)


@polar_backward.register_fake
def _(U, V, S, grad_R):
    return _empty_rows(U, _rows(U), 9)


@torch.library.custom_op(f"{_NS}::proper_rotation_full", mutates_args=(), device_types="cuda")
def proper_rotation_full(A: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    out = _ops.proper_svd(A, rot=True, S=True, U=True, V=True, what="batch_torch_A_to_R")
    return out["rot"], out["U"], out["V"], out["S"]


@proper_rotation_full.register_fake
def _(A):
    n = _rows(A)
    return _empty_rows(A, n, 3, 3), _empty_rows(A, n, 3, 3), _empty_rows(A, n, 3, 3), _empty_rows(A, n, 3)


def _rotation_setup(ctx, inputs, output):
    ctx.a_shape = inputs[0].shape
    ctx.save_for_backward(output[1], output[2], output[3])


def _rotation_backward(ctx, g_rot, _gU, _gV, _gS):
    U, V, S = ctx.saved_tensors
    return torch.ops.semiuhpe_b200.polar_backward(U, V, S, g_rot.contiguous()).view(ctx.a_shape)


proper_rotation_full.register_autograd(_rotation_backward, setup_context=_rotation_setup)


@torch.library.custom_op(f"{_NS}::fisher_ce", mutates_args=(), device_types="cuda")
def fisher_ce(A1: Tensor, A2: Tensor, want_grad: bool, target_G: Optional[Tensor] = None,
              keep: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    out = _ops.fisher_ce(A1, A2, grad=want_grad, target_G=target_G, keep=keep)
    return out["ce"], (out["grad"] if want_grad else _empty_rows(out["ce"], 0, 9))


@fisher_ce.register_fake
def _(A1, A2, want_grad, target_G=None, keep=None):
    n = _rows(A2)
    return _empty_rows(A2, n), _empty_rows(A2, n if want_grad else 0, 9)


def _fisher_ce_setup(ctx, inputs, output):
    _A1, A2, want_grad, _G, keep = inputs
    ctx.a_shape = A2.shape
    ctx.have_grad = want_grad
    ctx.save_for_backward(output[1], keep)


def _fisher_ce_backward(ctx, g_ce, _g_grad):
    grad, keep = ctx.saved_tensors
    if not ctx.have_grad:
        raise RuntimeError("semiuhpe_b200::fisher_ce was called with want_grad=False but A2 requires grad")
    return None, torch.ops.semiuhpe_b200.scale_rows(grad, g_ce, keep).view(ctx.a_shape), None, None, None


fisher_ce.register_autograd(_fisher_ce_backward, setup_context=_fisher_ce_setup)


# --------------------------------------------------------------------------- rotation Laplace
@torch.library.custom_op(f"{_NS}::laplace_nll", mutates_args=(), device_types="cuda")
def laplace_nll(pred: Tensor, gt: Tensor, grids: Tensor, want_grad: bool,
                keep: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    out = _ops.laplace_nll(pred, gt, grids, grad=want_grad, mode=True)
    return out["nll"], out["mode"], (out["grad"] if want_grad else _empty_rows(out["nll"], 0, 9))


@laplace_nll.register_fake
def _(pred, gt, grids, want_grad, keep=None):
    n = _rows(pred)
    return _empty_rows(pred, n), _empty_rows(pred, n, 3, 3), _empty_rows(pred, n if want_grad else 0, 9)


def _laplace_setup(ctx, inputs, output):
    pred, _gt, _grids, want_grad, keep = inputs
    ctx.p_shape = pred.shape
    ctx.have_grad = want_grad
    ctx.save_for_backward(output[2], keep)


def _laplace_backward(ctx, g_nll, _g_mode, _g_grad):
    grad, keep = ctx.saved_tensors
    if not ctx.have_grad:
        raise RuntimeError("semiuhpe_b200::laplace_nll was called with want_grad=False but pred requires grad")
    return torch.ops.semiuhpe_b200.scale_rows(grad, g_nll, keep).view(ctx.p_shape), None, None, None, None


laplace_nll.register_autograd(_laplace_backward, setup_context=_laplace_setup)


# --------------------------------------------------------------------------- metrics
@torch.library.custom_op(f"{_NS}::geodesic_deg", mutates_args=(), device_types="cuda")
def geodesic_deg(pred: Tensor, gt: Tensor) -> Tensor:
    return _ops.so3_metrics(pred, gt, geo=True)["geo"]


@geodesic_deg.register_fake
def _(pred, gt):
    return _empty_rows(pred, _rows(pred))


OPS = ("scale_rows", "fisher_nll", "fisher_entropy", "proper_rotation", "proper_rotation_full", "polar_backward",
       "fisher_ce", "laplace_nll", "geodesic_deg")
