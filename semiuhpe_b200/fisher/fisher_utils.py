"""Drop-in for the reference's ``src/fisher/fisher_utils.py`` on CUDA tensors.

Same names, arguments and return shapes; every function is ONE kernel launch
(K1 or K2 of ``csrc/fisher_kernels.cu``; ``fisher_CE`` is three) instead of the
reference's CPU SVD round trips and hundreds of elementwise ops.  Differences that are deliberate:

* inputs must be CUDA fp32 (the reference moves them to the CPU itself);
* ``fisher_entropy`` returns a plain tensor without a ``grad_fn`` (the agent uses it detached:
  src/agent.py:108,139); ``batch_torch_A_to_R`` is differentiable when its input requires a gradient;
* ``vmf_loss`` / ``KL_Fisher`` / ``fisher_log_pdf`` are differentiable w.r.t. the
  network output exactly like the reference (only the singular values enter
  autograd there, fisher_utils.py:29-31): the backward multiplies the per-sample
  gradient the forward launch already produced.
"""
import torch

from .. import _ops


class _FisherNLL(torch.autograd.Function):
    """nll_i = -<A_i,R_i> + overreg*logC(S_i);  d nll_i/dA_i from the same launch."""

    @staticmethod
    def forward(ctx, A, R, overreg, want_rot, keep):
        need_grad = ctx.needs_input_grad[0]
        out = _ops.fisher_fused(A, R, overreg, nll=True, grad=need_grad, rot=want_rot, what="KL_Fisher")
        if need_grad:
            ctx.save_for_backward(out["grad"])
        ctx.a_shape = A.shape
        ctx.keep = keep
        rot = out.get("rot")
        if rot is None:
            rot = torch.empty(0, device=A.device)
        ctx.mark_non_differentiable(rot)
        return out["nll"], rot

    @staticmethod
    def backward(ctx, g_nll, _g_rot):
        (grad,) = ctx.saved_tensors
        # one launch: saved per-sample gradient x incoming gradient; rows the caller filtered out are exact zeros
        return _ops.scale_rows(grad, g_nll, keep=ctx.keep).view(ctx.a_shape), None, None, None, None


def _compiling():
    return torch.compiler.is_compiling()


def _wants_grad(t):
    return bool(t.requires_grad and torch.is_grad_enabled())


def vmf_loss(net_out, R, overreg=1.05, keep=None):
    """(loss (b,), Rest (b,3,3))  -- reference fisher_utils.py:14-18.
    ``keep`` (extension, optional, (b,) bool): rows a mask filtered out; their gradient is exactly zero whatever
    the row holds (the reference gathers ``pred[mask]`` before the loss: src/agent.py:157)."""
    A = net_out.view(-1, 3, 3)
    if _compiling():                                   # traced / compiled step: the registered dispatcher op
        from .. import torch_ops  # noqa: F401
        loss_v, Rest, _ = torch.ops.semiuhpe_b200.fisher_nll(A, R, float(overreg), True, _wants_grad(A), keep)
        return loss_v, Rest.detach()                   # the projected rotation carries no gradient (as in eager mode)
    loss_v, Rest = _FisherNLL.apply(A, R, float(overreg), True, keep)
    return loss_v, Rest


def KL_Fisher(A, R, overreg=1.05):
    """Matrix-Fisher NLL (b,) -- reference fisher_utils.py:21-36."""
    if _compiling():
        from .. import torch_ops  # noqa: F401
        return torch.ops.semiuhpe_b200.fisher_nll(A, R, float(overreg), False, _wants_grad(A), None)[0]
    loss_v, _ = _FisherNLL.apply(A, R, float(overreg), False, None)
    return loss_v


class _ProperRotation(torch.autograd.Function):
    """R = U V^T of the proper SVD and its gradient (closed form of what autograd gives through torch.svd)."""

    @staticmethod
    def forward(ctx, A):
        out = _ops.proper_svd(A, rot=True, S=True, U=True, V=True, what="batch_torch_A_to_R")
        ctx.save_for_backward(out["U"], out["V"], out["S"])
        ctx.a_shape = A.shape
        return out["rot"]

    @staticmethod
    def backward(ctx, g_rot):
        U, V, S = ctx.saved_tensors
        return _ops.proper_svd_backward(U, V, S, g_rot).view(ctx.a_shape)


def batch_torch_A_to_R(A):
    """Proper-SVD projection onto SO(3), (b,9)|(b,3,3) -> (b,3,3)
    -- reference fisher_utils.py:39-48.  Differentiable like the reference (the agent never back-propagates
    it; the plain K1 launch is used unless ``A`` requires a gradient)."""
    if not _compiling() and _wants_grad(A):
        return _ProperRotation.apply(A)
    if _compiling():
        from .. import torch_ops  # noqa: F401
        if _wants_grad(A):
            return torch.ops.semiuhpe_b200.proper_rotation_full(A)[0]
        return torch.ops.semiuhpe_b200.proper_rotation(A)
    return _ops.proper_svd(A, rot=True, what="batch_torch_A_to_R")["rot"]


def fisher_log_pdf(A, R):
    """<A,R> - logC(S)  -- reference fisher_utils.py:51-67."""
    return -KL_Fisher(A, R, overreg=1.0)


def fisher_entropy(A):
    """(b,9)|(b,3,3) -> (b,) entropy of the matrix-Fisher distribution
    -- reference fisher_utils.py:70-81 (Fisher -> Bingham -> autograd chain, collapsed
    to H = log f(s) + sum_j s_j (1 - g_j))."""
    if _compiling():
        from .. import torch_ops  # noqa: F401
        return torch.ops.semiuhpe_b200.fisher_entropy(A)
    return _ops.fisher_fused(A, None, 1.0, entropy=True, what="fisher_entropy")["entropy"]


def fisher_nll_entropy(net_out, R, overreg=1.05):
    """Fused extra (no reference equivalent): NLL, projected rotation AND entropy of the
    same batch from one launch -> (loss, Rest, entropy)."""
    A = net_out.view(-1, 3, 3)
    out = _ops.fisher_fused(A, R, overreg, nll=True, rot=True, entropy=True, what="fisher_nll_entropy")
    return out["nll"], out["rot"], out["entropy"]


class _FisherCE(torch.autograd.Function):
    """ce_i and d ce_i / d A2_i from the same three launches."""

    @staticmethod
    def forward(ctx, A1, A2, target_G, keep):
        need_grad = ctx.needs_input_grad[1]
        out = _ops.fisher_ce(A1, A2, grad=need_grad, target_G=target_G, keep=keep)
        if need_grad:
            ctx.save_for_backward(out["grad"])
        ctx.a_shape = A2.shape
        ctx.keep = keep
        return out["ce"]

    @staticmethod
    def backward(ctx, g_ce):
        (grad,) = ctx.saved_tensors
        return None, _ops.scale_rows(grad, g_ce, keep=ctx.keep).view(ctx.a_shape), None, None


def fisher_CE(A1, A2, target_G=None, keep=None):
    """Cross entropy h(f1, f2) of two matrix-Fisher densities, A1 the target and A2 the
    prediction, (b,9)|(b,3,3) x2 -> (b,)  -- reference fisher_utils.py:84-99 (the default
    unsupervised loss, src/agent.py:155).  Differentiable w.r.t. A2 like the reference (through the
    SVD, the quaternion frame and logC_F, here in closed form).  A1 is a constant: the agent feeds
    the detached teacher prediction (src/agent.py:107); asking for its gradient is an error rather
    than a silent zero.  NaN/Inf results raise AssertionError as in the reference (:98).
    ``target_G`` (extension, optional): d logC/dS of the target, (b,3), e.g. the ``G`` output of the
    entropy launch on the teacher prediction -- unchanged by the rotate-augmentation adjustment --
    which saves the target's quadrature (one K2 launch instead of two).
    ``keep`` (extension, optional, (b,) bool): the rows the reference would have gathered with
    ``[mask_fisher]`` before the call (src/agent.py:155).  Filtered rows return 0, get an exactly zero
    gradient and are not checked: the assert of :98 looks at the kept rows only, as in the reference."""
    if isinstance(A1, torch.Tensor) and A1.requires_grad and torch.is_grad_enabled():
        raise NotImplementedError("fisher_CE: the gradient w.r.t. the target A1 is not implemented "
                                  "(the reference's training loop detaches it, src/agent.py:107)")
    if _compiling():
        from .. import torch_ops  # noqa: F401
        return torch.ops.semiuhpe_b200.fisher_ce(A1, A2, _wants_grad(A2), target_G, keep)[0]
    return _FisherCE.apply(A1, A2, target_G, keep)
