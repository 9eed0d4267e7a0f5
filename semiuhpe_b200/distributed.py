"""Multi-GPU form of the dynamic-entropy threshold (BASELINE config 5).

The reference is single-process.  Here every rank owns a shard
of the unlabeled pool; samples are independent, so NLL/entropy/metrics need no
communication.  The only exchange is the global k-th smallest entropy: per radix
pass each rank histograms its shard (K3), the 2048 uint64 counters are
ALL-GATHERED over NCCL (16 KB per rank -- latency-bound, NVLink bandwidth is
irrelevant), and every rank runs the same integer scan over the gathered block,
so all ranks hold the bit-identical threshold that a single GPU (or
``numpy.sort`` on the concatenated pool) would produce.

``HistogramBackend`` separates the per-rank device work from the collective
logic so the latter is testable with ``gloo`` on CPU (tests use a numpy backend
there; the product backend is CUDA-only).
"""
import torch
import torch.distributed as dist

from . import _capi, _ops
from .agent import pool_index


class CudaHistogramBackend:
    """K3 kernels on this rank's shard."""

    def __init__(self, entropy):
        self.e = _ops._entropy_vector(entropy)
        self.ws = _ops.SelectWorkspace(self.e.device)
        self.device = self.e.device

    def size(self):
        return self.e.numel()

    def init(self, k):
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib().suhpe_select_init(_capi.ptr(self.ws.state), k, _capi.stream()), "select_init")

    def local_hist(self, pass_no, first_pass_hist=None):
        if pass_no == 1 and first_pass_hist is not None:
            return first_pass_hist
        h = self.ws.hist[1]
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib().suhpe_select_hist_f32(_capi.ptr(self.e), self.e.numel(), pass_no,
                                                          _capi.ptr(self.ws.state), _capi.ptr(h), _capi.stream()),
                        "select_hist")
        return h

    def scan(self, gathered, pass_no):
        parts = gathered.shape[0]
        with torch.cuda.device(self.device):
            _capi.check(_capi.lib().suhpe_select_scan(_capi.ptr(gathered), parts, pass_no,
                                                      _capi.ptr(self.ws.state), _capi.stream()), "select_scan")

    def result(self):
        return self.ws.read()[0]


def global_entropy_threshold(entropy_shard, left_ratio, group=None, backend=None, first_pass_hist=None,
                             n_total=None, sync=True):
    """Threshold over the union of all ranks' shards; identical on every rank.

    ``k = int(n_total * left_ratio)`` exactly as src/agent.py:406 on the concatenated pool
    (``n_total`` is all-reduced when the caller does not know it).  Then three rounds of
    {local histogram -> all-gather (world, 2048) int64 -> identical scan on every rank}."""
    backend = backend or CudaHistogramBackend(entropy_shard)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    device = backend.device
    if n_total is None:
        n_total = backend.size()
        if world > 1:
            sizes = torch.tensor([n_total], dtype=torch.int64, device=device)
            dist.all_reduce(sizes, group=group)
            n_total = int(sizes.item())
    backend.init(pool_index(n_total, left_ratio))
    for pass_no in (1, 2, 3):
        local = backend.local_hist(pass_no, first_pass_hist)
        if world > 1:
            gathered = torch.empty((world, _capi.HIST_BINS), dtype=torch.int64, device=device)
            dist.all_gather_into_tensor(gathered, local.reshape(1, -1).contiguous(), group=group)
        else:
            gathered = local.reshape(1, -1)
        backend.scan(gathered, pass_no)
    return backend.result() if sync else backend


def sharded_mean(values, group=None, weights=None):
    """``losses.mean()`` (src/agent.py:83,163) when the batch is sharded over the ranks of ``group``: the mean of
    the CONCATENATED per-sample values, identical on every rank -- one all-reduce of (sum, count) in float64, so
    the integer part (the count) is exact and the sum is the float64 sum of the ranks' float64 partial sums.
    ``weights`` (same shape, e.g. a keep mask): returns ``sum(values * weights) / total_count`` -- the
    masked-mean-times-mask-ratio form of the unsupervised loss (:163-166).  Differentiable: each rank's
    gradient is ``weights / total_count``, which is what the single-process mean gives its slice."""
    v = values.reshape(-1)
    w = None if weights is None else weights.reshape(-1).to(v.dtype)
    local = (v if w is None else torch.where(w != 0, v, torch.zeros((), dtype=v.dtype, device=v.device)) * w)
    stats = torch.stack((local.detach().to(torch.float64).sum(),
                         torch.tensor(float(v.numel()), dtype=torch.float64, device=v.device)))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, group=group)
    total = stats[1].clamp(min=1.0)
    mean = (stats[0] / total).to(v.dtype)
    # value from the exact global statistics, gradient through this rank's slice only
    return mean.detach() + (local.sum() - local.sum().detach()) / total.to(v.dtype)
