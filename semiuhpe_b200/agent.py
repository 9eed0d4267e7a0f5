"""Drop-ins for the filter / threshold / metric slice of the reference's ``SSLAgent``
(src/agent.py:148-150, 229-232, 357-417, 420-455).

The reference's agent class itself (networks, optimiser, EMA, checkpoints) is out
of scope and keeps running unchanged; it only needs these functions swapped in
(INTEGRATION.md).  Everything below runs on the GPU: the pool threshold is an
exact radix select over the entropies where they were produced (K3) instead of a
device->host copy per batch plus ``numpy.sort``.
"""
import torch

from . import _ops
from .fisher.fisher_utils import fisher_entropy

_WORKSPACES = {}


def select_workspace(device):
    """The radix-select scratch (32-byte state + two histograms) of the current stream on ``device``: allocated
    once and reused -- a fresh one per filter call costs three allocations and two fills on a 20 us path."""
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, _ops.stream(idx))
    ws = _WORKSPACES.get(key)
    if ws is None:
        ws = _WORKSPACES[key] = _ops.SelectWorkspace(torch.device("cuda", idx))
    return ws


# ------------------------------------------------------------------ a12: filter
def pool_index(n, left_ratio):
    """``index = int(len(entropy_all) * left_ratio)`` (src/agent.py:406): Python float
    product, truncation toward zero.  ``left_ratio >= 1`` raises IndexError like the
    reference's numpy indexing at :407 (negative ratios wrap there; rejected here)."""
    index = int(n * left_ratio)
    if not 0 <= index < n:
        raise IndexError(f"index {index} is out of bounds for axis 0 with size {n}")
    return index


def entropy_threshold(entropies, left_ratio, workspace=None, first_pass_hist=None, sync=True):
    """k-th smallest entropy of the pool, k = int(n*left_ratio), in ``numpy.sort`` order
    (ascending, NaN last) -- the value ``entropy_all.sort(); entropy_all[index]`` of
    src/agent.py:403-407.  Returns a Python float, or with ``sync=False`` the
    :class:`semiuhpe_b200._ops.SelectWorkspace` holding it on the device."""
    e = _ops._entropy_vector(entropies)          # raises for CPU tensors (no CPU path)
    k = pool_index(e.numel(), left_ratio)
    ws = _ops.entropy_threshold_device(e, k, workspace or select_workspace(e.device), first_pass_hist)
    if not sync:
        return ws
    return ws.read()[0]


def entropy_mask(entropy, conf_thres):
    """``mask = entropy < conf_thres`` and ``mask_ratio = mask.sum() / len(mask)``
    (src/agent.py:148-150).  ``conf_thres``: float (yml value or a threshold returned by
    :func:`entropy_threshold`) or a SelectWorkspace (threshold stays on the device)."""
    mask, kept = _ops.entropy_mask(entropy, conf_thres)
    ratio = kept.to(torch.float32) / max(mask.numel(), 1)
    return mask, ratio.reshape(())


def dynamic_entropy_filter(pred_weak, left_ratio, return_threshold=True):
    """Teacher-batch filter of BASELINE config 2: entropies of the unlabeled teacher
    outputs (K2), threshold at rank int(n*left_ratio) (K3), strict-< keep mask -- three
    launches chained on the current stream, no host sync unless the threshold value is
    requested.  Returns (entropy, mask, mask_ratio[, threshold float])."""
    A = pred_weak.reshape(-1, 9)
    n = A.shape[0]
    k = pool_index(n, left_ratio)
    ws = select_workspace(A.device)
    ws.hist[0].zero_()
    ent = _ops.fisher_fused(A, None, 1.0, entropy=True, hist=ws.hist[0], what="fisher_entropy")["entropy"]
    _ops.entropy_threshold_device(ent, k, ws, first_pass_hist=ws.hist[0])
    mask, kept = _ops.entropy_mask(ent, ws)
    ratio = (kept.to(torch.float32) / n).reshape(())
    if return_threshold:
        return ent, mask, ratio, ws.read()[0]
    return ent, mask, ratio


def compute_dynamic_entropy_threshold(agent, ulb_train_bar):
    """Function form of ``SSLAgent.compute_dynamic_entropy_threshold`` (src/agent.py:357-417):
    run the EMA teacher over the unlabeled loader, collect ``fisher_entropy`` of every
    batch ON THE DEVICE, select the ``left_ratio`` percentile and store it in
    ``agent.config.conf_thres``.

    ``config.save_feat`` (src/agent.py:366-401: hooks on the backbone, t-SNE feature / image dumps to
    ``log_dir``) is host-side bookkeeping that stays with the reference: when it is set the call is handed to
    the reference's own method (``install.patch_agent_class`` keeps it as ``_reference_compute_dynamic_entropy_threshold``),
    so a patched agent loses nothing; without a reference method to hand over to, the flag is an error rather
    than a silently skipped dump."""
    if getattr(agent.config, "save_feat", False):
        reference = getattr(type(agent), "_reference_compute_dynamic_entropy_threshold", None)
        if reference is None:
            raise NotImplementedError("config.save_feat: the feature dump of src/agent.py:366-401 is not part of the CUDA "
                                      "path; patch the reference agent with semiuhpe_b200.install (it keeps the "
                                      "reference method for this case) or clear the flag")
        return reference(agent, ulb_train_bar)
    agent.ema_net.eval()
    chunks = []
    with torch.no_grad():
        for ulb_data in ulb_train_bar:
            pred_weak = agent.ema_net(ulb_data.get("img").cuda())
            chunks.append(fisher_entropy(pred_weak))
    entropy_all = torch.cat(chunks, 0)
    thr = entropy_threshold(entropy_all, agent.config.left_ratio)
    print("The best dynamic entropy threshold is:", thr)
    agent.config.conf_thres = thr
    return thr


# --------------------------------------------------- 8f-3: the unsupervised branch, sync-free
def unsupervised_terms(pred_weak, pred_strong, conf_thres, *, type_unsuper="ce", distribution="matrixFisher",
                       grids=None, aug_rot_mat=None, train_labeled="300WLP", ulb_gt=None, overreg=1.025):
    """The unsupervised half of ``SSLAgent.forward`` (src/agent.py:99-192) without its host
    synchronisations: the reference tests ``mask_ratio_fisher > 0`` on the host (:151) and gathers
    ``pred[mask_fisher]`` (dynamic shapes, :152-160); here every sample goes through the loss
    kernel and the mask enters as a weight -- ``mean(l[mask]) * mask_ratio == sum(where(mask, l, 0)) / b``,
    with the same gradients: the mask is handed to the loss (``keep=``), whose backward writes exact zeros for
    filtered rows whatever they hold (a NaN teacher row never reaches the weights, and the NaN assert of
    fisher_utils.py:98 looks at the kept rows only, like the reference's gather) -- so the step is a fixed launch sequence
    (CUDA-graph capturable with ``semiuhpe_b200.set_error_checking(False)``).

    pred_weak: teacher output (b,9) (detached like :107); pred_strong: student output (b,9);
    conf_thres: float or the SelectWorkspace of :func:`entropy_threshold` (``sync=False``).
    Returns device tensors: unsuper_loss (scalar, already multiplied by the mask ratio like :166),
    entropy (b,), mask (b,) bool, mask_ratio, and the three error terms of :169-180 as per-sample
    vectors over the WHOLE batch plus their masked means (the reference returns the masked subsets)."""
    from .fisher.fisher_utils import batch_torch_A_to_R, fisher_CE, vmf_loss
    pred_weak = pred_weak.detach()
    b = pred_weak.reshape(-1, 9).shape[0]
    # :139 (not rotate-adjusted); the same launch yields g = d logC/dS of the teacher, which fisher_CE
    # needs for the target and which the adjustment (a rotation on one side) does not change
    stats = _ops.fisher_fused(pred_weak, None, 1.0, entropy=True, G=(type_unsuper == "ce"), what="fisher_entropy")
    entropy = stats["entropy"]
    mask, mask_ratio = entropy_mask(entropy, conf_thres)                  # :148-150
    adjusted = pred_weak.reshape(-1, 9) if aug_rot_mat is None else rotate_aug_adjust(pred_weak, aug_rot_mat, train_labeled)
    pseudo = batch_torch_A_to_R(adjusted)                                 # :152
    if type_unsuper == "ce":                                              # :155,160 (both distributions)
        losses = fisher_CE(adjusted, pred_strong, target_G=stats["G"], keep=mask)
    elif type_unsuper == "nll" and distribution == "matrixFisher":        # :157
        losses, _ = vmf_loss(pred_strong, pseudo, overreg=overreg, keep=mask)
    elif type_unsuper == "nll" and distribution == "RotationLaplace":     # :162
        from .laplace.rotation_laplace import NLL_loss
        losses, _ = NLL_loss("RLaplace", pred_strong, pseudo, grids, keep=mask)
    else:
        raise ValueError(f"unsupervised_terms: unknown loss {type_unsuper!r} / distribution {distribution!r}")
    zero = torch.zeros((), dtype=losses.dtype, device=losses.device)
    unsuper_loss = torch.where(mask, losses, zero).sum() / b              # == mean(l[mask]) * mask_ratio  (:163,166)
    out = dict(unsuper_loss=unsuper_loss, entropy=entropy, mask=mask, mask_ratio=mask_ratio)
    kept = mask.sum().clamp(min=1)
    masked_mean = lambda v: torch.where(mask, v, zero).sum() / kept
    strong_rot = batch_torch_A_to_R(pred_strong.detach())
    out["err_strongSuper_pseudo"] = compute_err_deg_from_matrices(strong_rot, pseudo)      # :177-180
    out["err_strongSuper_pseudo_mean"] = masked_mean(out["err_strongSuper_pseudo"])
    if ulb_gt is not None:                                                # :169-172
        out["err_weakAll_gt"] = compute_err_deg_from_matrices(pseudo, ulb_gt)
        out["err_weakPseudo_gt_mean"] = masked_mean(out["err_weakAll_gt"])
    return out


# --------------------------------------------------- the validation slice, sync-free
def validation_terms(pred, pred_orth, gt, conf_thres, gt_euler=None):
    """The mask -> masked-error slice of ``SSLAgent.val_func`` (src/agent.py:224-249) without its host
    round trips: the reference reads ``fisher_mask_ratio`` back with ``.item()`` (:230), tests it on the host
    (:232) and gathers ``pred_orth[fisher_mask]`` / ``gt[fisher_mask]`` (dynamic shapes, :239-241).  Here the
    error runs over the whole batch and the mask enters as a weight.

    pred: network output (b,9); pred_orth: its projected rotations (b,3,3); gt: (b,3,3); gt_euler: (b,3) degrees
    or None; conf_thres: float or SelectWorkspace.  Returns device tensors:
      entropy (b,), mask (b,) bool, mask_ratio,
      err_all (b,)        -- compute_err_deg_from_matrices over every row (equals ``err_pseudo_gt`` on the kept rows)
      err_pseudo_gt_sum   -- sum of the kept rows' errors; err_pseudo_gt_mean = sum / max(kept, 1)
    ``err_pseudo_gt`` of the reference (a tensor of kept rows, or None when nothing is kept) is
    ``out["err_all"][out["mask"]]`` -- the one line that needs a sync, left to the caller."""
    entropy = fisher_entropy(pred)                                        # :226
    mask, mask_ratio = entropy_mask(entropy, conf_thres)                  # :229-230
    err_all = compute_err_deg_from_matrices(pred_orth, gt, gt_euler=gt_euler)   # :239-241
    zero = torch.zeros((), dtype=err_all.dtype, device=err_all.device)
    err_sum = torch.where(mask, err_all, zero).sum()
    kept = mask.sum()
    return dict(entropy=entropy, mask=mask, mask_ratio=mask_ratio, err_all=err_all, err_pseudo_gt_sum=err_sum,
                err_pseudo_gt_mean=err_sum / kept.clamp(min=1), kept=kept)


# --------------------------------------------------- the whole loss head of a training step as one call
_SSL_STEPS = {}


def _ssl_handle(device):
    """One handle (two side streams + three events, no data) per device and host thread."""
    import threading
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, threading.get_ident())
    h = _SSL_STEPS.get(key)
    if h is None:
        h = _SSL_STEPS[key] = _ops.SslStep(torch.device("cuda", idx))
    return h


class _SslLoss(torch.autograd.Function):
    """loss_all and its gradients w.r.t. both student outputs from ONE C call (suhpe_ssl_step_f32)."""

    @staticmethod
    def forward(ctx, out_l, pred_strong, gt, pred_weak, aug_rot, opts):
        want_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        b_l = out_l.reshape(-1, 9).shape[0]
        b_u = 0 if pred_weak is None else pred_weak.reshape(-1, 9).shape[0]
        res = _ssl_handle(out_l.device).run(
            out_l, gt, pred_weak, pred_strong, opts["conf_thres"], aug_rot=aug_rot, aug_mode=opts["aug_mode"],
            overreg=opts["overreg"], ssl_lambda=opts["ssl_lambda"], unsup=opts["unsup"], want_grad=want_grad)
        if want_grad:
            ctx.save_for_backward(res["grads"])          # both gradients, contiguous: (b_l + b_u, 9)
        ctx.rows = (b_l, b_u)
        ctx.shapes = (out_l.shape, None if pred_strong is None else pred_strong.shape)
        opts["result"] = res
        return res["losses"][3]

    @staticmethod
    def backward(ctx, g):
        (grads,) = ctx.saved_tensors
        b_l, b_u = ctx.rows
        shape_l, shape_s = ctx.shapes
        scaled = _ops.scale_rows(grads, scalar_weight=g)                  # one launch for both student outputs
        gl = scaled[:b_l].view(shape_l)
        gs = None if shape_s is None or b_u == 0 else scaled[b_l:].view(shape_s)
        return gl, gs, None, None, None, None


def ssl_loss(fisher_out, gt, pred_weak=None, pred_strong=None, conf_thres=0.0, *, SSL_lambda=1.0, type_unsuper="ce",
             aug_rot_mat=None, train_labeled="300WLP", overreg=1.025):
    """``loss_all`` of ``SSLAgent.train_func`` (src/agent.py:194-210) -- ``forward``'s supervised NLL mean (:76-83),
    the unsupervised branch (:99-166: teacher entropy, mask, rotate-augmentation adjustment, fisher_CE or NLL
    against the pseudo labels, masked mean x mask ratio) and ``loss + SSL_lambda * unsuper_loss`` (:203) --
    as ONE C call whose backward is one more launch per student output.  matrixFisher heads only (the
    Laplace head goes through :func:`unsupervised_terms`).  Without ``pred_weak`` it is the supervised step of
    ``train_func_s1`` (:253-270).

    Returns ``(loss_all, info)``: ``loss_all`` is a scalar tensor with autograd history to ``fisher_out`` and
    ``pred_strong``; ``info`` holds detached device tensors -- ``loss``, ``unsuper_loss`` (already x mask ratio),
    ``mask_ratio``, ``pred_orth`` (b,3,3), ``entropy``, ``mask``, ``pseudo_labels`` (every row), ``losses``,
    ``unsuper_losses`` (0 on filtered rows).  No host synchronisation with error checking off, and no allocation
    by the library inside the call (the scratch is a torch allocation): the step works under
    ``torch.cuda.graph`` / ``torch.cuda.make_graphed_callables`` once the first eager call on the device has run."""
    if type_unsuper not in ("ce", "nll"):
        raise ValueError(f"ssl_loss: unknown type_unsuper {type_unsuper!r}")
    if train_labeled not in ("DAD3DHeads", "300WLP"):
        raise ValueError(f"ssl_loss: unknown train_labeled {train_labeled!r}")
    if pred_weak is not None:
        pred_weak = pred_weak.detach()                                    # :107
    opts = dict(conf_thres=conf_thres, aug_mode=0 if train_labeled == "DAD3DHeads" else 1, overreg=float(overreg),
                ssl_lambda=float(SSL_lambda), unsup=type_unsuper)
    loss_all = _SslLoss.apply(fisher_out, pred_strong, gt, pred_weak, aug_rot_mat, opts)
    res = opts["result"]
    L = res["losses"]
    info = dict(loss=L[0], unsuper_loss=L[1], mask_ratio=L[2], pred_orth=res["pred_orth"], entropy=res["entropy"],
                mask=res["mask"], pseudo_labels=res["pseudo"], losses=res["losses_l"], unsuper_losses=res["losses_u"])
    return loss_all, info


# ------------------------------------------------------------ 8f-4: EMA / EMAN teacher update
def _fusable(ema_t, src_t):
    """fp32 and contiguous: what the multi-tensor kernel takes.  Other CUDA tensors (fp16 / bf16 / fp64 entries,
    channels_last weights) keep the reference's own torch expression on the GPU; CPU tensors are an error."""
    if not (ema_t.is_cuda and src_t.is_cuda):
        raise RuntimeError("update_ema_variables: every tensor must live on a CUDA device (semiuhpe_b200 has no CPU path)")
    return (ema_t.dtype == torch.float32 and src_t.dtype == torch.float32
            and ema_t.is_contiguous() and src_t.is_contiguous())


def update_ema_variables(net, ema_net, is_ema, alpha, global_step, eman=False):
    """``SSLAgent.update_ema_variables`` (src/agent.py:277-299) for a student ``net`` and its teacher
    ``ema_net``: the warm-up rule ``alpha = min(1 - 1/(global_step+1), alpha)`` (``alpha = 0`` when
    ``is_ema`` is false), then either the EMAN blend over the whole ``state_dict`` (``num_batches_tracked``
    copied, :290-293) or the plain EMA over ``parameters()`` (:297-298) -- each a handful of multi-tensor
    launches instead of two torch ops per tensor.  Returns the alpha that was applied."""
    alpha = min(1 - 1 / (global_step + 1), alpha) if is_ema else 0
    with torch.no_grad():
        if eman:
            main, ema = net.state_dict(), ema_net.state_dict()
            e_list, s_list = [], []
            for (k_main, v_main), (k_ema, v_ema) in zip(main.items(), ema.items()):
                assert k_main == k_ema, "state_dict names are different!"
                assert v_main.shape == v_ema.shape, "state_dict shapes are different!"
                if "num_batches_tracked" in k_ema:
                    v_ema.copy_(v_main)                                   # integer counters (:290-291)
                elif _fusable(v_ema, v_main):
                    e_list.append(v_ema)
                    s_list.append(v_main)
                else:                                                     # fp16 / bf16 / fp64 / strided entries: the
                    v_ema.copy_(v_ema * alpha + (1. - alpha) * v_main)    # reference's own expression (:293)
            _ops.ema_update(e_list, s_list, alpha, 0)
        else:
            e_list, s_list = [], []
            for ema_param, param in zip(ema_net.parameters(), net.parameters()):
                if _fusable(ema_param.data, param.data):
                    e_list.append(ema_param.data)
                    s_list.append(param.detach())
                else:
                    ema_param.data.mul_(alpha).add_(param.data, alpha=1 - alpha)   # :298
            _ops.ema_update(e_list, s_list, alpha, 1)
    return alpha


# ------------------------------------------------------------ a13..a16: metrics
def rotate_aug_adjust(pred_weak, aug_rot_mat, train_labeled):
    """``pred_weak_adjusted`` of src/agent.py:110-122: the teacher's (b,9) parameters expressed in the
    frame of the rotate-augmented strong view.  ``train_labeled``: "DAD3DHeads" (left product with the
    augmentation rotation) or "300WLP" (the transposed convention, conjugated by Rx(180))."""
    if train_labeled not in ("DAD3DHeads", "300WLP"):
        raise ValueError(f"rotate_aug_adjust: unknown train_labeled {train_labeled!r}")
    return _ops.rotate_adjust(pred_weak, aug_rot_mat, 0 if train_labeled == "DAD3DHeads" else 1)


def compute_err_deg_from_matrices(pred, gt, gt_euler=None):
    """(b,3,3),(b,3,3)[,(b,3) degrees] -> (b,) error in degrees (src/agent.py:447-455):
    geodesic angle via pytorch3d's ``so3_relative_angle`` semantics when ``gt_euler`` is
    None, else the mean absolute (pitch,yaw,roll) error."""
    if gt_euler is None:
        if torch.compiler.is_compiling():
            from . import torch_ops  # noqa: F401
            return torch.ops.semiuhpe_b200.geodesic_deg(pred, gt)
        return _ops.so3_metrics(pred, gt, geo=True)["geo"]
    return _ops.so3_metrics(pred, gt, gt_euler, full_range=False, mae=True)["mae"]


def compute_err_deg_from_quats(pred, gt):
    """src/agent.py:420-424: geodesic error of two real-first quaternion batches."""
    return compute_err_deg_from_matrices(_quat_to_matrix(pred), _quat_to_matrix(gt))


def _quat_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def eval_rotation_metrics(pred, gt, gt_euler=None, dad_trained=False):
    """One K4 launch for the evaluation loop of eval.py:76-98,125-133.

    With ``gt_euler`` (degrees): per-angle absolute errors (n,3) and their means
    ``(pitch, yaw, roll, mean)``; without: geodesic degrees (n,), Frobenius distance
    (n,) and their means.  Means are accumulated in fp64 on the device.  ``dad_trained`` selects
    the Euler convention of models trained on DAD-3DHeads (eval.py:60-74, ``config.train_labeled``)."""
    n = pred.reshape(-1, 9).shape[0]
    if gt_euler is not None:
        out = _ops.so3_metrics(pred, gt, gt_euler, full_range="dad" if dad_trained else False,
                               abs_err=True, mae=True, sums=True)
        s = out["sums"] / max(n, 1)
        return dict(abs_err=out["abs_err"], mae=out["mae"], pitch=s[2], yaw=s[3], roll=s[4], mean=s[5])
    out = _ops.so3_metrics(pred, gt, geo=True, frob=True, sums=True)
    s = out["sums"] / max(n, 1)
    return dict(geodesic_deg=out["geo"], frobenius=out["frob"], geodesic_mean=s[0], frobenius_mean=s[1])
