"""Host-buffer entry of the C ABI: the teacher-side filter step over a pool that lives
in HOST memory (``suhpe_fisher_filter_host``).

This is what a non-torch host (or ``bench.py``'s end-to-end leg) calls: pinned host
arrays in, pinned host arrays out; chunks move through three in-order queues
(H2D copies, the fused Fisher kernel + first radix histogram, D2H copies) over a
ring of chunk buffers, then the percentile threshold and the keep-mask are
produced on the device and copied back.  torch is used only to own the buffers.

With a process group (one rank per GPU, each owning a shard of the pool in its host
memory) ``run(..., group=...)`` uses the two-phase entry ``suhpe_fisher_pool_host``:
the shard's entropies stay on the device, the global k-th smallest entropy is found by
the all-gathered radix select of ``semiuhpe_b200.distributed`` and every rank emits the
mask of its shard -- the threshold is the one a single GPU would find on the whole pool.
"""
import ctypes

import torch

from . import _capi
from .agent import pool_index


class FisherFilterPipeline:
    def __init__(self, max_n, chunk=1 << 20, device=0, cut_bits=None):
        if not torch.cuda.is_available():
            raise RuntimeError("FisherFilterPipeline needs a CUDA device: semiuhpe_b200 has no CPU path")
        self.device = torch.device("cuda", device)
        self.max_n, self.chunk = int(max_n), int(chunk)
        self._h = ctypes.c_void_p()
        import semiuhpe_b200 as _pkg
        bits = _pkg.quadrature_cut_bits() if cut_bits is None else int(cut_bits)
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib().suhpe_pipeline_create(ctypes.byref(self._h), self.max_n, self.chunk, bits),
                        "pipeline_create")
        self._out = None

    def close(self):
        if self._h:
            _capi.lib().suhpe_pipeline_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _buffers(self, n, want_grad):
        if self._out is None or self._out["n"] != n or (want_grad and self._out["grad"] is None):
            self._out = dict(
                n=n,
                nll=torch.empty(n, dtype=torch.float32).pin_memory(),
                grad=torch.empty((n, 9), dtype=torch.float32).pin_memory() if want_grad else None,
                entropy=torch.empty(n, dtype=torch.float32).pin_memory(),
                mask=torch.empty(n, dtype=torch.bool).pin_memory())
        return self._out

    def run(self, A_host, R_host, overreg=1.025, left_ratio=0.95, want_grad=True, group=None):
        """A_host, R_host: CPU fp32 (n,9) tensors (pinned for full PCIe rate).
        Returns dict(nll, grad, entropy, mask (CPU, pinned), threshold: float, kept: int).
        ``group``: a torch.distributed group (or True for the default group) whose ranks each
        hold a shard; the threshold is then global and ``kept`` counts this rank's shard."""
        for name, t in (("A_host", A_host), ("R_host", R_host)):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise TypeError(f"{name} must be a contiguous CPU float32 tensor")
        n = A_host.reshape(-1, 9).shape[0]
        if n > self.max_n:
            raise ValueError(f"pool of {n} exceeds the pipeline capacity {self.max_n}")
        out = dict(self._buffers(n, want_grad))
        if not want_grad:
            out["grad"] = None                      # a cached gradient buffer of an earlier call is not filled (nor copied)
        if group is not None:
            return self._run_sharded(A_host, R_host, n, overreg, left_ratio, want_grad, out,
                                     None if group is True else group)
        k = pool_index(n, left_ratio)
        thr, kept = ctypes.c_float(), ctypes.c_uint64()
        P = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(self.device):
            code = _capi.check(_capi.lib().suhpe_fisher_filter_host(
                self._h, P(A_host), P(R_host), n, float(overreg), k, P(out["nll"]), P(out["grad"]),
                P(out["entropy"]), P(out["mask"]), ctypes.byref(thr), ctypes.byref(kept)), "fisher_filter_host")
        if code & _capi.STATUS_NONFINITE:
            raise torch.linalg.LinAlgError("fisher_filter_host: the input contains non-finite values")
        res = dict(out)
        res.update(threshold=thr.value, kept=int(kept.value),
                   h2d_bytes=n * 72, d2h_bytes=n * (4 + 4 + 1 + (36 if want_grad else 0)))
        return res

    def _run_sharded(self, A_host, R_host, n, overreg, left_ratio, want_grad, out, group):
        from . import _ops
        from .distributed import global_entropy_threshold
        dev = self.device
        P = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(dev):
            if getattr(self, "_dev", None) is None or self._dev["ent"].numel() != n:
                self._dev = dict(ent=torch.empty(n, dtype=torch.float32, device=dev),
                                 hist=torch.empty(_capi.HIST_BINS, dtype=torch.int64, device=dev),
                                 status=torch.empty(1, dtype=torch.int32, device=dev))
            d = self._dev
            d["hist"].zero_()
            d["status"].zero_()
            _capi.check(_capi.lib().suhpe_fisher_pool_host(
                self._h, P(A_host), P(R_host), n, float(overreg), P(out["nll"]), P(out["grad"]), P(out["entropy"]),
                _capi.ptr(d["ent"]), _capi.ptr(d["hist"]), _capi.ptr(d["status"]), _capi.stream()), "fisher_pool_host")
            # global k-th smallest entropy: 3 x {local histogram, all-gather, identical scan} on the current stream
            backend = global_entropy_threshold(d["ent"], left_ratio, group=group, first_pass_hist=d["hist"], sync=False)
            mask, kept = _ops.entropy_mask(d["ent"], backend.ws)
            out["mask"].copy_(mask, non_blocking=True)
            thr = backend.result()                       # host read: synchronises the current stream
            _capi.check(_capi.lib().suhpe_pipeline_sync(self._h), "pipeline_sync")
            torch.cuda.current_stream().synchronize()
            if int(d["status"].item()) & _capi.STATUS_NONFINITE:
                raise torch.linalg.LinAlgError("fisher_pool_host: the input contains non-finite values")
        res = dict(out)
        res.update(threshold=thr, kept=int(kept.item()),
                   h2d_bytes=n * 72, d2h_bytes=n * (4 + 4 + 1 + (36 if want_grad else 0)))
        return res
