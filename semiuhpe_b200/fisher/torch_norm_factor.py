"""Drop-in for the reference's ``src/fisher/torch_norm_factor.py``: ``logC_F``.

``logC_F(S)`` is the matrix-Fisher log-normaliser by the reference's 512-node
trapezoid quadrature with the Abramowitz-Stegun Bessel polynomials
(torch_norm_factor.py:66-92); forward value and the gradient ``d logC / dS``
come out of the same CUDA launch (K2 run on given singular values)."""
import torch

from .. import _ops


class _LogCF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, S):
        out = _ops.fisher_from_s(S, logC=True, G=True)
        ctx.save_for_backward(out["G"])
        ctx.in_shape = S.shape
        return out["logC"].view(S.shape[:-1])

    @staticmethod
    def backward(ctx, grad):
        (G,) = ctx.saved_tensors
        return _ops.scale_rows(G, grad).view(ctx.in_shape)


def logC_F(S):
    """(...,3) singular values sorted s1>=s2>=|s3| -> (...) log normaliser
    (reference: ``logC_F = class_logC_F.apply``, torch_norm_factor.py:92)."""
    if S.shape[-1] != 3:
        raise RuntimeError("logC_F expects (..., 3) singular values")
    return _LogCF.apply(S)
