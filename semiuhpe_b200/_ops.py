"""Tensor-level wrappers over the C ABI (one CUDA launch each, on the current stream).

These allocate the outputs with torch's caching allocator, pass raw device
pointers through ``ctypes`` and translate the device status word into the
exceptions the reference raises.  Nothing here computes: no torch math, no CPU.
"""
import ctypes

import torch

import semiuhpe_b200 as _pkg
from . import _capi
from ._capi import as_records, check, lib, on_device, ptr, stream

_STATUS = {}


def _status_word(device):
    """One status word per (device, stream): a bit set by a launch on one stream is never reported against
    an op running on another."""
    key = (device.index, stream(device.index))
    t = _STATUS.get(key)
    if t is None:
        t = torch.zeros(1, dtype=torch.int32, device=device)
        _STATUS[key] = t
    return t


def reset_status():
    """Clear every status word (called when error checking is switched back on: bits raised while it was
    off -- graph replays, sync-free steps -- must not be blamed on the next unrelated call)."""
    for t in _STATUS.values():
        t.zero_()


def _raise_from_status(status, what):
    """One 4-byte read (sync), only when error checking is on (reference parity:
    torch.svd raises LinAlgError on NaN/Inf, pytorch3d raises ValueError)."""
    if not _pkg.error_checking():
        return
    bits = int(status.item())
    if bits:
        status.zero_()
        if bits & _capi.STATUS_NONFINITE:
            raise torch.linalg.LinAlgError(
                f"{what}: the input contains non-finite values (torch.svd in the reference fails the same way)")
        if bits & _capi.STATUS_TRACE_RANGE:
            raise ValueError("A matrix has trace outside valid range [-1-eps,3+eps].")
        if bits & _capi.STATUS_NONFINITE_CE:
            raise AssertionError(f"{what}: the cross entropy is NaN or Inf (the reference asserts: fisher_utils.py:98)")


def _cut(cut_bits):
    return _pkg.quadrature_cut_bits() if cut_bits is None else int(cut_bits)


def _keep_vector(keep, n, name="keep"):
    """(n,) bool / uint8 CUDA tensor -> contiguous uint8 view (None passes through)."""
    if keep is None:
        return None
    if not keep.is_cuda or keep.numel() != n or keep.dtype not in (torch.bool, torch.uint8):
        raise RuntimeError(f"{name} must be a CUDA bool/uint8 tensor with {n} elements")
    k = keep.reshape(-1)
    if not k.is_contiguous():
        k = k.contiguous()
    return k.view(torch.uint8) if k.dtype == torch.bool else k


def fisher_fused(A, R=None, overreg=1.0, *, nll=False, grad=False, rot=False, entropy=False,
                 logC=False, S=False, G=False, hist=None, what="fisher", cut_bits=None):
    """K2.  Returns a dict of the requested outputs (keys = argument names)."""
    A9 = as_records(A, "A")
    n = A9.shape[0]
    R9 = None
    if R is not None:
        R9 = as_records(R, "R")
        if R9.shape[0] != n:
            raise RuntimeError(f"shape mismatch: A has {n} matrices, R has {R9.shape[0]}")
    dev = A9.device
    new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    out = {}
    if nll: out["nll"] = new(n)
    if grad: out["grad"] = new(n, 9)
    if rot: out["rot"] = new(n, 3, 3)
    if entropy: out["entropy"] = new(n)
    if logC: out["logC"] = new(n)
    if S: out["S"] = new(n, 3)
    if G: out["G"] = new(n, 3)
    if n == 0:
        return out
    with on_device(dev) as idx:
        status = _status_word(dev)
        check(lib().suhpe_fisher_fused_f32(
            ptr(A9), ptr(R9), n, float(overreg), _cut(cut_bits), ptr(out.get("nll")), ptr(out.get("grad")),
            ptr(out.get("rot")), ptr(out.get("entropy")), ptr(out.get("logC")), ptr(out.get("S")),
            ptr(out.get("G")), ptr(hist), ptr(status), stream(idx)), what)
    _raise_from_status(status, what)
    return out


def fisher_ce(A1, A2, *, grad=False, target_G=None, keep=None, cut_bits=None):
    """fisher_CE value (n,) and, on request, d ce_i / d A2_i (n,9): two K2 launches + the closing kernel
    (one K2 launch when ``target_G`` = d logC/dS of the target (n,3) is supplied).  ``keep`` (n, bool):
    rows the caller's mask filtered out -- they get ce = 0, a zero gradient and raise nothing."""
    T9, P9 = as_records(A1, "A1"), as_records(A2, "A2")
    n = P9.shape[0]
    if T9.shape[0] != n:
        raise RuntimeError(f"shape mismatch: A1 has {T9.shape[0]} matrices, A2 has {n}")
    dev = P9.device
    out = {"ce": torch.empty(n, dtype=torch.float32, device=dev)}
    if grad: out["grad"] = torch.empty((n, 9), dtype=torch.float32, device=dev)
    if n == 0:
        return out
    K = _keep_vector(keep, n)
    work = torch.empty(_capi.FISHER_CE_WORKSPACE_FLOATS * n, dtype=torch.float32, device=dev)
    with on_device(dev) as idx:
        status = _status_word(dev)
        if target_G is None:
            check(lib().suhpe_fisher_ce_f32(ptr(T9), ptr(P9), n, _cut(cut_bits), ptr(K), ptr(out["ce"]), ptr(out.get("grad")),
                                            ptr(work), ptr(status), stream(idx)), "fisher_CE")
        else:
            G3 = as_records(target_G, "target_G", 3)
            if G3.shape[0] != n:
                raise RuntimeError(f"shape mismatch: target_G has {G3.shape[0]} rows, A2 has {n}")
            check(lib().suhpe_fisher_ce_with_g1_f32(ptr(T9), ptr(G3), ptr(P9), n, _cut(cut_bits), ptr(K), ptr(out["ce"]),
                                                    ptr(out.get("grad")), ptr(work), ptr(status), stream(idx)), "fisher_CE")
    _raise_from_status(status, "fisher_CE")
    return out


def scale_rows(rows, row_weight=None, scalar_weight=None, keep=None):
    """``rows * row_weight[:, None] * scalar_weight`` with filtered rows (``keep`` false) written as exact zeros:
    the backward of the loss mirrors (per-sample gradient x incoming gradient) as one launch."""
    if not rows.is_cuda or rows.dtype != torch.float32:
        raise RuntimeError("scale_rows: rows must be a CUDA float32 tensor")
    r = rows if rows.is_contiguous() else rows.contiguous()
    n = r.shape[0] if r.dim() > 1 else r.numel()
    width = r.numel() // max(n, 1) if n else 1
    out = torch.empty_like(r)
    if n == 0:
        return out
    w = None
    if row_weight is not None:
        w = row_weight.detach().reshape(-1)
        if w.numel() == 1 and scalar_weight is None and n != 1:
            scalar_weight, w = w, None                  # a broadcast scalar (e.g. the gradient of a .sum())
        else:
            if w.numel() == 1:
                w = w.expand(n)
            if w.numel() != n:
                raise RuntimeError(f"scale_rows: {w.numel()} weights for {n} rows")
            w = w.to(torch.float32)
            w = w if w.is_contiguous() else w.contiguous()
    sw = None
    if scalar_weight is not None:
        sw = scalar_weight.detach().reshape(-1).to(torch.float32)
    K = _keep_vector(keep, n)
    with on_device(r.device) as idx:
        check(lib().suhpe_scale_rows_f32(ptr(r), n, width, ptr(w), ptr(sw), ptr(K), ptr(out), stream(idx)), "scale_rows")
    return out


def rotate_adjust(pred, aug_rot, mode):
    """(n,9) teacher parameters moved into the rotate-augmented view (src/agent.py:110-119)."""
    P9, R9 = as_records(pred, "pred_weak"), as_records(aug_rot, "aug_rot_mat")
    n = P9.shape[0]
    if R9.shape[0] != n:
        raise RuntimeError(f"shape mismatch: pred_weak has {n} matrices, aug_rot_mat has {R9.shape[0]}")
    out = torch.empty((n, 9), dtype=torch.float32, device=P9.device)
    if n:
        with on_device(P9.device) as idx:
            check(lib().suhpe_rotate_adjust_f32(ptr(P9), ptr(R9), n, int(mode), ptr(out), stream(idx)), "rotate_adjust")
    return out


def ema_update(ema_tensors, src_tensors, alpha, mode):
    """In-place EMA blend of a list of fp32 CUDA tensors (src/agent.py:293,298) in ceil(count/48)-ish
    launches.  ``alpha`` is a Python float, rounded to fp32 here exactly like ATen rounds a Python scalar
    multiplied into an fp32 tensor; so is ``1 - alpha`` (computed in double first, as the reference does)."""
    ema_tensors, src_tensors = list(ema_tensors), list(src_tensors)
    if len(ema_tensors) != len(src_tensors):
        raise RuntimeError(f"ema_update: {len(ema_tensors)} teacher tensors against {len(src_tensors)} student tensors")
    if not ema_tensors:
        return
    dev = ema_tensors[0].device
    for e, s in zip(ema_tensors, src_tensors):
        if not (e.is_cuda and s.is_cuda) or e.device != dev or s.device != dev:
            raise RuntimeError("ema_update: every tensor must live on the same CUDA device (no CPU path)")
        if e.dtype != torch.float32 or s.dtype != torch.float32:
            raise TypeError("ema_update: fp32 tensors only (copy integer buffers with copy_)")
        if e.shape != s.shape:
            raise RuntimeError(f"ema_update: shape mismatch {tuple(e.shape)} vs {tuple(s.shape)}")
        if not (e.is_contiguous() and s.is_contiguous()):
            raise RuntimeError("ema_update: tensors must be contiguous (parameters and buffers are)")
    count = len(ema_tensors)
    PtrArr, NumArr = ctypes.c_void_p * count, ctypes.c_int64 * count
    e_arr = PtrArr(*[e.data_ptr() for e in ema_tensors])
    s_arr = PtrArr(*[s.data_ptr() for s in src_tensors])
    n_arr = NumArr(*[e.numel() for e in ema_tensors])
    a32 = ctypes.c_float(float(alpha)).value
    oma32 = ctypes.c_float(1.0 - float(alpha)).value
    with on_device(dev) as idx:
        check(lib().suhpe_ema_update_f32(e_arr, s_arr, n_arr, count, a32, oma32, int(mode), stream(idx)), "ema_update")


def fisher_from_s(S, *, logC=True, G=False, entropy=False, cut_bits=None):
    """K2 on given singular values (logC_F)."""
    S3 = as_records(S, "S", 3)
    n = S3.shape[0]
    dev = S3.device
    out = {}
    if logC: out["logC"] = torch.empty(n, dtype=torch.float32, device=dev)
    if G: out["G"] = torch.empty((n, 3), dtype=torch.float32, device=dev)
    if entropy: out["entropy"] = torch.empty(n, dtype=torch.float32, device=dev)
    if n:
        with on_device(dev) as idx:
            check(lib().suhpe_fisher_from_s_f32(ptr(S3), n, _cut(cut_bits), ptr(out.get("logC")), ptr(out.get("G")),
                                                ptr(out.get("entropy")), None, stream(idx)), "logC_F")
    return out


def proper_svd(A, *, rot=True, S=False, U=False, V=False, what="svd"):
    """K1."""
    A9 = as_records(A, "A")
    n = A9.shape[0]
    dev = A9.device
    new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    out = {}
    if rot: out["rot"] = new(n, 3, 3)
    if S: out["S"] = new(n, 3)
    if U: out["U"] = new(n, 3, 3)
    if V: out["V"] = new(n, 3, 3)
    if n == 0:
        return out
    status = _status_word(dev)
    with on_device(dev) as idx:
        check(lib().suhpe_proper_svd_f32(ptr(A9), n, ptr(out.get("rot")), ptr(out.get("S")),
                                         ptr(out.get("U")), ptr(out.get("V")), ptr(status), stream(idx)), what)
    _raise_from_status(status, what)
    return out


def proper_svd_backward(U, V, S, grad_R):
    """d L / d A (n,9) from d L / d R for R = U V^T of the proper SVD (K1's U, V, S)."""
    U9, V9, G9 = as_records(U, "U"), as_records(V, "V"), as_records(grad_R, "grad_R")
    S3 = as_records(S, "S", 3)
    n = U9.shape[0]
    out = torch.empty((n, 9), dtype=torch.float32, device=U9.device)
    if n:
        with on_device(U9.device) as idx:
            check(lib().suhpe_proper_svd_backward_f32(ptr(U9), ptr(V9), ptr(S3), ptr(G9), n, ptr(out), stream(idx)),
                  "proper_svd_backward")
    return out


def laplace_nll(A, R, grids, *, grad=False, mode=True, logF=False):
    """K2L."""
    A9 = as_records(A, "pred")
    R9 = as_records(R, "gt")
    g9 = as_records(grids, "grids")
    n, N = A9.shape[0], g9.shape[0]
    if R9.shape[0] != n:
        raise RuntimeError(f"shape mismatch: pred has {n} matrices, gt has {R9.shape[0]}")
    dev = A9.device
    new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    out = {"nll": new(n)}
    if grad: out["grad"] = new(n, 9)
    if mode: out["mode"] = new(n, 3, 3)
    if logF: out["logF"] = new(n)
    if n == 0:
        return out
    status = _status_word(dev)
    with on_device(dev) as idx:
        check(lib().suhpe_laplace_nll_f32(ptr(A9), ptr(R9), n, ptr(g9), N, ptr(out["nll"]), ptr(out.get("grad")),
                                          ptr(out.get("mode")), ptr(out.get("logF")), ptr(status), stream(idx)),
              "laplace_nll")
    _raise_from_status(status, "laplace_nll")
    return out


def so3_metrics(Rp, Rg=None, gt_euler=None, *, full_range=False, geo=False, frob=False, euler=False,
                abs_err=False, mae=False, sums=False):
    """K4."""
    P9 = as_records(Rp, "pred")
    n = P9.shape[0]
    G9 = as_records(Rg, "gt") if Rg is not None else None
    E3 = as_records(gt_euler, "gt_euler", 3) if gt_euler is not None else None
    for name, t in (("gt", G9), ("gt_euler", E3)):
        if t is not None and t.shape[0] != n:
            raise RuntimeError(f"shape mismatch: pred has {n} rows, {name} has {t.shape[0]}")
    dev = P9.device
    new = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    out = {}
    if geo: out["geo"] = new(n)
    if frob: out["frob"] = new(n)
    if euler: out["euler"] = new(n, 3)
    if abs_err: out["abs_err"] = new(n, 3)
    if mae: out["mae"] = new(n)
    if sums: out["sums"] = torch.zeros(8, dtype=torch.float64, device=dev)
    if n == 0:
        return out
    status = _status_word(dev)
    with on_device(dev) as idx:
        mode = 2 if full_range == "dad" else int(bool(full_range))
        check(lib().suhpe_so3_metrics_f32(ptr(P9), ptr(G9), ptr(E3), n, mode,
                                          ptr(out.get("geo")), ptr(out.get("frob")), ptr(out.get("euler")),
                                          ptr(out.get("abs_err")), ptr(out.get("mae")), ptr(out.get("sums")),
                                          ptr(status), stream(idx)), "so3_metrics")
    if geo or sums:
        _raise_from_status(status, "so3_relative_angle")
    return out


class SelectWorkspace:
    """Device scratch of the radix select: 32-byte state + two 2048-bin uint64 histograms."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.state = torch.zeros(_capi.SELECT_STATE_BYTES // 8, dtype=torch.int64, device=self.device)
        self.hist = torch.zeros((2, _capi.HIST_BINS), dtype=torch.int64, device=self.device)

    def threshold_ptr(self):
        return ctypes.c_void_p(lib().suhpe_select_threshold_ptr(ptr(self.state)))

    def read(self):
        thr, key, kept = ctypes.c_float(), ctypes.c_uint32(), ctypes.c_uint64()
        with on_device(self.device) as idx:
            check(lib().suhpe_select_read(ptr(self.state), ctypes.byref(thr), ctypes.byref(key),
                                          ctypes.byref(kept), stream(idx)), "select_read")
        return thr.value, key.value, kept.value


def _entropy_vector(e):
    if not e.is_cuda:
        raise RuntimeError("entropy must be a CUDA tensor: semiuhpe_b200 has no CPU path")
    if e.dtype != torch.float32:
        raise TypeError("entropy must be float32")
    e = e.detach().reshape(-1)
    return e if e.is_contiguous() else e.contiguous()


def entropy_threshold_device(entropy, k, ws=None, first_pass_hist=None):
    """K3 on one GPU: leaves the k-th smallest entropy in ``ws.state`` (no host sync)."""
    e = _entropy_vector(entropy)
    n = e.numel()
    if not 0 <= k < n:
        raise IndexError(f"index {k} is out of bounds for axis 0 with size {n}")
    ws = ws or SelectWorkspace(e.device)
    with on_device(e.device) as idx:
        check(lib().suhpe_entropy_threshold_f32(ptr(e), n, k, ptr(ws.state), ptr(ws.hist[1]),
                                                ptr(first_pass_hist), stream(idx)), "entropy_threshold")
    return ws


def entropy_mask(entropy, thr, ws=None, want_mask=True):
    """mask = entropy < thr (strict).  ``thr``: python float or a SelectWorkspace
    (threshold read on the device).  Returns (mask bool tensor | None, kept-count tensor int64[1])."""
    e = _entropy_vector(entropy)
    n = e.numel()
    mask = torch.empty(n, dtype=torch.bool, device=e.device) if want_mask else None
    kept = torch.zeros(1, dtype=torch.int64, device=e.device)
    if n:
        with on_device(e.device) as idx:
            if isinstance(thr, SelectWorkspace):
                check(lib().suhpe_entropy_mask_f32(ptr(e), n, thr.threshold_ptr(), 0.0, ptr(mask), ptr(kept),
                                                   stream(idx)), "entropy_mask")
            else:
                check(lib().suhpe_entropy_mask_f32(ptr(e), n, None, float(thr), ptr(mask), ptr(kept),
                                                   stream(idx)), "entropy_mask")
    return mask, kept


class SslStep:
    """Handle of ``suhpe_ssl_step_f32``: the loss head of a whole semi-supervised step -- supervised Fisher NLL,
    teacher entropy, mask, rotate-augmentation adjustment, fisher_CE (or NLL) against the pseudo labels, the
    means, ``loss_all`` and its gradients -- as one C call (a fixed launch sequence over three forked streams).
    The handle holds two side streams and three events, no data: the scratch of a call is a torch allocation, so
    nothing is allocated by the library inside the call and the step can be issued during CUDA-graph capture."""

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("SslStep needs a CUDA device: semiuhpe_b200 has no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._h = ctypes.c_void_p()
        with on_device(self.device):
            check(lib().suhpe_ssl_step_create(ctypes.byref(self._h)), "ssl_step_create")

    def close(self):
        if self._h:
            lib().suhpe_ssl_step_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, out_l, gt_l, pred_weak=None, pred_strong=None, conf_thres=0.0, *, aug_rot=None, aug_mode=0,
            overreg=1.025, ssl_lambda=1.0, unsup="ce", want_grad=True, cut_bits=None, extras=True):
        """Raw tensors in, dict of raw tensors out (no autograd: see ``semiuhpe_b200.agent.ssl_loss``)."""
        L9, G9 = as_records(out_l, "fisher_out"), as_records(gt_l, "gt")
        b_l = L9.shape[0]
        if G9.shape[0] != b_l:
            raise RuntimeError(f"shape mismatch: {b_l} labeled outputs, {G9.shape[0]} labels")
        b_u = 0
        W9 = S9 = R9 = None
        if pred_weak is not None:
            W9, S9 = as_records(pred_weak, "pred_weak"), as_records(pred_strong, "pred_strong")
            b_u = W9.shape[0]
            if S9.shape[0] != b_u:
                raise RuntimeError(f"shape mismatch: {b_u} teacher outputs, {S9.shape[0]} student outputs")
            if aug_rot is not None:
                R9 = as_records(aug_rot, "aug_rot_mat")
                if R9.shape[0] != b_u:
                    raise RuntimeError(f"shape mismatch: {b_u} teacher outputs, {R9.shape[0]} augmentation rotations")
        if b_l == 0:
            raise ValueError("ssl step without labeled rows")
        if unsup not in ("ce", "nll"):
            raise ValueError(f"unknown unsupervised loss {unsup!r}")
        dev = L9.device
        # one allocation for every fp32 output
        sizes = [4, b_l * 9 if want_grad else 0, b_u * 9 if want_grad else 0]
        if extras:
            sizes += [b_l * 9, b_u, b_u * 9, b_l, b_u]
        n_work = int(lib().suhpe_ssl_step_workspace_floats(b_l, b_u))
        tail = (-sum(sizes)) % 4                                   # the workspace starts on a 16-byte boundary
        flat = torch.empty(sum(sizes) + tail + n_work, dtype=torch.float32, device=dev)
        work = flat[sum(sizes) + tail:]
        parts, off = [], 0
        for sz in sizes:
            parts.append(flat[off:off + sz] if sz else None)
            off += sz
        losses, grad_l, grad_s = parts[:3]
        rest = ent = pseudo = nll_l = loss_u = mask = None
        if extras:
            rest, ent, pseudo, nll_l, loss_u = parts[3:]
            mask = torch.empty(b_u, dtype=torch.bool, device=dev) if b_u else None
        thr_dev, thr_host = None, 0.0
        if isinstance(conf_thres, SelectWorkspace):
            thr_dev = conf_thres.threshold_ptr()
        elif isinstance(conf_thres, torch.Tensor):
            thr_dev = ptr(conf_thres.detach().reshape(-1).to(torch.float32))
        else:
            thr_host = float(conf_thres)
        with on_device(dev) as idx:
            status = _status_word(dev)
            check(lib().suhpe_ssl_step_f32(
                self._h, ptr(L9), ptr(G9), b_l, ptr(W9), ptr(S9), b_u, ptr(R9), int(aug_mode), thr_dev, thr_host,
                float(overreg), float(ssl_lambda), 0 if unsup == "ce" else 1, _cut(cut_bits), ptr(work),
                ptr(losses), ptr(grad_l), ptr(grad_s), ptr(rest), ptr(ent), ptr(mask), ptr(pseudo), ptr(nll_l), ptr(loss_u),
                ptr(status), stream(idx)), "ssl_step")
        _raise_from_status(status, "ssl_step")
        out = dict(losses=losses, grad_l=None if grad_l is None else grad_l.view(b_l, 9),
                   grad_strong=None if grad_s is None else grad_s.view(b_u, 9),
                   grads=flat[4:4 + (b_l + b_u) * 9].view(b_l + b_u, 9) if want_grad else None)
        if extras:
            out.update(pred_orth=rest.view(b_l, 3, 3), entropy=ent, mask=mask, pseudo=None if pseudo is None else pseudo.view(b_u, 3, 3),
                       losses_l=nll_l, losses_u=loss_u)
        return out
