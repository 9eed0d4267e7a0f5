"""Drop-in for the reference's ``src/laplace/rotation_laplace.py`` on CUDA tensors.

``NLL_loss("RLaplace", pred, gt, grids)`` is one launch of K2L
(``csrc/laplace_kernels.cu``): forward, the gradient w.r.t. ``pred`` and the
analytical mode together, without materialising the reference's (b,N,3,3)
broadcast.  TF32 is never used (the reference disables it,
rotation_laplace.py:13).  The grid-Fisher verification path
(``fn_type="RFisher"``, rotation_laplace.py:118-137) and the ``broadcast``
density evaluation are not hot: they are composed from torch ops on the GPU.
"""
import torch

from .. import _ops

EPS = 1e-8  # rotation_laplace.py:11


def delta_R(N):
    """Haar volume element of an N-point SO(3) grid (rotation_laplace.py:15-21)."""
    return 1 / N


class _LaplaceNLL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, gt, grids, keep):
        need_grad = ctx.needs_input_grad[0]
        out = _ops.laplace_nll(pred, gt, grids, grad=need_grad, mode=True)
        if need_grad:
            ctx.save_for_backward(out["grad"])
        ctx.p_shape = pred.shape
        ctx.keep = keep
        ctx.mark_non_differentiable(out["mode"])
        return out["nll"], out["mode"]

    @staticmethod
    def backward(ctx, g_nll, _g_mode):
        (grad,) = ctx.saved_tensors
        return _ops.scale_rows(grad, g_nll, keep=ctx.keep).view(ctx.p_shape), None, None, None


def analytical_mode(pred, fn_type="RLaplace"):
    """(pred_orth (b,3,3), s3sign (b,))  -- rotation_laplace.py:102-115.
    ``s3sign`` is exactly +-1 here (the reference returns det(U V^T) ~ +-1)."""
    out = _ops.proper_svd(pred.reshape(-1, 3, 3), rot=True, S=True, what="analytical_mode")
    sign = torch.where(out["S"][:, 2] < 0, -1.0, 1.0).to(torch.float32)
    return out["rot"], sign


def _power_fisher(A, x):
    return (A * x).sum(dim=(-1, -2))


def _signed_trace(A):
    out = _ops.proper_svd(A.reshape(-1, 3, 3), rot=False, S=True, what="power_fn_sqrtL2_proper")
    return out["S"].sum(-1)


def log_pdf(fn_type, A, x, grids, broadcast=False):
    """rotation_laplace.py:76-99.  (b,) or, with ``broadcast`` / x being the grid, (b,N)."""
    A = A.reshape(-1, 3, 3)
    over_grid = (x.shape[0] == grids.shape[0]) or broadcast
    if fn_type == "RLaplace":
        if not over_grid:
            return -_LaplaceNLL.apply(A, x, grids, None)[0]
        # density over a whole grid (visualisation / checks, not the training path): logF and the signed trace
        # come from the kernels, which do not record autograd history, so the result is returned WITHOUT a
        # grad_fn rather than with a partial one (the reference is differentiable here; nothing in it uses that)
        if A.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError("log_pdf('RLaplace') over a grid / with broadcast=True is not differentiable here; "
                                      "call it under torch.no_grad() (NLL_loss / the per-sample form are differentiable)")
        with torch.no_grad():
            logF = _ops.laplace_nll(A, A.new_zeros(A.shape) + torch.eye(3, device=A.device), grids,
                                    mode=False, logF=True)["logF"]
            T = _signed_trace(A)
            tr = torch.einsum("bij,nij->bn", A, x.reshape(-1, 3, 3))
            power = -torch.sqrt(torch.clamp_min(T[:, None] - tr, EPS))
            return -logF[:, None] + power - torch.log(-power)
    if fn_type == "RFisher":
        tr_grid = torch.einsum("bij,nij->bn", A, grids.reshape(-1, 3, 3))
        c = tr_grid.max(dim=-1)[0]
        logF = c + torch.log(torch.exp(tr_grid - c[:, None]).sum(1) * delta_R(grids.shape[0]))
        if over_grid:
            return -logF[:, None] + torch.einsum("bij,nij->bn", A, x.reshape(-1, 3, 3))
        return -logF + _power_fisher(A, x.reshape(-1, 3, 3))
    raise KeyError(fn_type)


def NLL_loss(fn_type, pred, gt, grids, keep=None):
    """(losses (b,), pred_orth (b,3,3))  -- rotation_laplace.py:24-34.  ``keep``: see fisher_utils.vmf_loss."""
    pred = pred.reshape(-1, 3, 3)
    if fn_type == "RLaplace":
        if torch.compiler.is_compiling():               # traced / compiled step: the registered dispatcher op
            from .. import torch_ops  # noqa: F401
            nll, mode, _ = torch.ops.semiuhpe_b200.laplace_nll(pred, gt, grids, bool(pred.requires_grad and torch.is_grad_enabled()), keep)
            return nll, mode.detach()
        return _LaplaceNLL.apply(pred, gt, grids, keep)
    losses = -log_pdf(fn_type, pred, gt, grids)
    pred_orth, _ = analytical_mode(pred, fn_type)
    return losses, pred_orth
