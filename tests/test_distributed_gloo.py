"""World-size-2 (and 3) gloo run of the global-threshold exchange on CPU.

The collective logic of semiuhpe_b200.distributed (all-gather of per-rank radix
histograms -> identical scan on every rank) is exercised with a numpy histogram
backend standing in for the K3 kernels (which need a GPU); the product backend is
CUDA-only and is covered by the `gpu` tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _keys(e):
    e = np.where(e == 0, np.float32(0.0), e).astype(np.float32)
    u = e.view(np.uint32)
    k = np.where(u & 0x80000000, ~u, u | np.uint32(0x80000000)).astype(np.uint32)
    k[np.isnan(e)] = 0xFFFFFFFF
    return k


class NumpyHistogramBackend:
    """Test double of CudaHistogramBackend (same interface, host arithmetic)."""

    def __init__(self, entropy):
        self.k = _keys(np.asarray(entropy, np.float32))
        self.device = torch.device("cpu")
        self.prefix, self.rank_left, self.key = 0, 0, None

    def size(self):
        return len(self.k)

    def init(self, k):
        self.rank_left, self.prefix = int(k), 0

    def local_hist(self, pass_no, first_pass_hist=None):
        k = self.k
        if pass_no == 1:
            d = k >> 21
        elif pass_no == 2:
            d = ((k >> 10) & 2047)[(k >> 21) == self.prefix]
        else:
            d = (k & 1023)[(k >> 10) == self.prefix]
        return torch.from_numpy(np.bincount(d, minlength=2048).astype(np.int64))

    def scan(self, gathered, pass_no):
        tot = gathered.sum(0).numpy()
        cum = np.cumsum(tot)
        b = int(np.searchsorted(cum, self.rank_left, side="right"))
        self.rank_left -= int(cum[b - 1]) if b else 0
        bits = 10 if pass_no == 3 else 11
        self.prefix = b if pass_no == 1 else (self.prefix << bits) | b
        if pass_no == 3:
            self.key = self.prefix

    def result(self):
        key = np.uint32(self.key)
        if key == 0xFFFFFFFF:
            return float("nan")
        u = (key & np.uint32(0x7FFFFFFF)) if key & 0x80000000 else ~key
        return float(np.array([u], np.uint32).view(np.float32)[0])


def _worker(rank, world, port, pool, ratios, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from semiuhpe_b200.distributed import global_entropy_threshold
    bounds = np.linspace(0, len(pool), world + 1).astype(int)
    bounds[1:-1] += np.arange(1, world) * 7 % 5          # ragged shards
    shard = pool[bounds[rank]:bounds[rank + 1]]
    res = [global_entropy_threshold(None, float(r), backend=NumpyHistogramBackend(shard)) for r in ratios]
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_global_threshold_equals_single_sort(world, golden):
    g = golden("select")
    ratios = [0.95, 0.75, 0.5, 0.0, 0.05, 0.999]
    for pool in (g["entropy"], g["ties"]):
        mgr = mp.Manager()
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), pool, ratios, out), nprocs=world, join=True)
        want = []
        for r in ratios:
            srt = np.sort(pool)
            want.append(float(srt[int(len(pool) * r)]))
        for rank in range(world):
            got = out[rank]
            for a, b in zip(got, want):
                assert (np.isnan(a) and np.isnan(b)) or a == b, (rank, got, want)


def test_single_process_path_without_process_group(golden):
    from semiuhpe_b200.distributed import global_entropy_threshold
    pool = golden("select")["entropy"]
    got = global_entropy_threshold(None, 0.95, backend=NumpyHistogramBackend(pool))
    assert got == float(np.sort(pool)[int(len(pool) * 0.95)])


def _mean_worker(rank, world, port, values, weights, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from semiuhpe_b200.distributed import sharded_mean
    bounds = np.linspace(0, len(values), world + 1).astype(int)
    bounds[1:-1] += 3                                        # ragged shards
    lo, hi = bounds[rank], bounds[rank + 1]
    v = torch.from_numpy(values[lo:hi].copy()).requires_grad_(True)
    w = torch.from_numpy(weights[lo:hi].copy())
    plain = sharded_mean(v)
    masked = sharded_mean(v, weights=w)
    (plain + 2 * masked).backward()
    out[rank] = (float(plain), float(masked), v.grad.numpy().copy(), lo, hi)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_mean_equals_single_process_mean(world):
    """losses.mean() (src/agent.py:83) and the masked unsupervised mean x mask ratio (:163-166) over a batch that is
    sharded across ranks: the value every rank gets and the gradient its slice gets are the single-process ones."""
    rng = np.random.default_rng(5)
    n = 1000
    values = rng.normal(size=n).astype(np.float32) * 7
    weights = (rng.random(n) < 0.8).astype(np.float32)
    values[weights == 0] = np.nan                            # filtered rows may hold anything (0 * NaN must not form)
    out = mp.Manager().dict()
    mp.spawn(_mean_worker, args=(world, _free_port(), values, weights, out), nprocs=world, join=True)
    kept = weights != 0
    want_masked = float(values[kept].astype(np.float64).sum() / n)
    ref = torch.from_numpy(np.where(kept, values, 0).astype(np.float32)).requires_grad_(True)
    for rank in range(world):
        plain, masked, grad, lo, hi = out[rank]
        assert np.isnan(plain)                               # the plain mean of a pool with NaN rows is NaN, as in torch
        assert abs(masked - want_masked) <= 1e-6 * abs(want_masked)
        assert np.allclose(grad[kept[lo:hi]], 1.0 / n + 2.0 / n, rtol=1e-6)
        assert masked == out[0][1]                           # bit-identical on every rank
    # without NaN rows the plain mean is the global mean
    values2 = np.where(kept, values, 1.5).astype(np.float32)
    out2 = mp.Manager().dict()
    mp.spawn(_mean_worker, args=(world, _free_port(), values2, weights, out2), nprocs=world, join=True)
    for rank in range(world):
        assert abs(out2[rank][0] - float(values2.astype(np.float64).mean())) <= 1e-6
        assert out2[rank][0] == out2[0][0]
