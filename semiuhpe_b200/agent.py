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
    e = entropies.reshape(-1)
    k = pool_index(e.numel(), left_ratio)
    ws = _ops.entropy_threshold_device(e, k, workspace, first_pass_hist)
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
    ws = _ops.SelectWorkspace(A.device)
    ws.hist.zero_()
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
    ``agent.config.conf_thres``.  (The reference's optional feature dump, ``save_feat``,
    is host-side bookkeeping and stays with the reference agent.)"""
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
    with the same gradients (zero for filtered samples) -- so the step is a fixed launch sequence
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
        losses = fisher_CE(adjusted, pred_strong, target_G=stats["G"])
    elif type_unsuper == "nll" and distribution == "matrixFisher":        # :157
        losses, _ = vmf_loss(pred_strong, pseudo, overreg=overreg)
    elif type_unsuper == "nll" and distribution == "RotationLaplace":     # :162
        from .laplace.rotation_laplace import NLL_loss
        losses, _ = NLL_loss("RLaplace", pred_strong, pseudo, grids)
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


# ------------------------------------------------------------ 8f-4: EMA / EMAN teacher update
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
                if "num_batches_tracked" in k_ema or v_ema.dtype != torch.float32:
                    v_ema.copy_(v_main)                                   # integer counters (:290-291)
                else:
                    e_list.append(v_ema)
                    s_list.append(v_main)
            _ops.ema_update(e_list, s_list, alpha, 0)
        else:
            pairs = list(zip(ema_net.parameters(), net.parameters()))
            _ops.ema_update([e.data for e, _ in pairs], [p.detach() for _, p in pairs], alpha, 1)
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
