"""Host-buffer pipeline of the C ABI (suhpe_fisher_filter_host): identical results to the
device-resident path, including ragged chunking."""
import ctypes

import numpy as np
import pytest
import torch

from helpers import random_rotations

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,chunk", [(1000, 1 << 20), (100003, 4096), (300000, 65536)])
def test_pipeline_matches_device_path(cuda, n, chunk):
    from semiuhpe_b200 import _capi, _ops
    from semiuhpe_b200.agent import pool_index
    from semiuhpe_b200.host_pipeline import FisherFilterPipeline
    gen = torch.Generator().manual_seed(n)
    A = (10 * torch.randn(n, 9, generator=gen)).pin_memory()
    R = random_rotations(n, gen).reshape(n, 9).pin_memory()
    pipe = FisherFilterPipeline(max_n=n, chunk=chunk)
    res = pipe.run(A, R, overreg=1.025, left_ratio=0.95)
    dev = _ops.fisher_fused(A.to(cuda), R.to(cuda), 1.025, nll=True, grad=True, entropy=True)
    assert torch.equal(res["nll"], dev["nll"].cpu())
    assert torch.equal(res["grad"], dev["grad"].cpu())
    assert torch.equal(res["entropy"], dev["entropy"].cpu())
    e = res["entropy"].numpy()
    k = pool_index(n, 0.95)
    thr = np.sort(e)[k]
    assert res["threshold"] == float(thr)
    assert np.array_equal(res["mask"].numpy(), e < thr)
    assert res["kept"] == int((e < thr).sum())
    pipe.close()


def test_two_phase_pipeline_matches_single_call(cuda):
    """suhpe_fisher_pool_host + select on the caller's stream (the form the sharded pool uses)
    gives the results of the one-call entry; with one rank the 'global' threshold is the local one."""
    from semiuhpe_b200.host_pipeline import FisherFilterPipeline
    n = 200001
    gen = torch.Generator().manual_seed(5)
    A = (10 * torch.randn(n, 9, generator=gen)).pin_memory()
    R = random_rotations(n, gen).reshape(n, 9).pin_memory()
    pipe = FisherFilterPipeline(max_n=n, chunk=8192)
    one = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in pipe.run(A, R, 1.025, 0.95).items()}
    for _ in range(2):                                   # twice: buffers and events are reused
        two = pipe.run(A, R, 1.025, 0.95, group=True)
        for key in ("nll", "grad", "entropy", "mask"):
            assert torch.equal(one[key], two[key]), key
        assert one["threshold"] == two["threshold"] and one["kept"] == two["kept"]
    pipe.close()


@pytest.mark.parametrize("n", [1, 33, 4097])
@pytest.mark.parametrize("dataset", ["DAD3DHeads", "300WLP"])
def test_rotate_aug_adjust(cuda, n, dataset):
    """src/agent.py:110-119: exact up to the rounding of a 3-term dot product."""
    from oracle import so3_oracle as orc
    from semiuhpe_b200.agent import rotate_aug_adjust
    gen = torch.Generator().manual_seed(n)
    P = 10 * torch.randn(n, 9, generator=gen)
    Raug = random_rotations(n, gen)
    ours = rotate_aug_adjust(P.to(cuda), Raug.to(cuda), dataset).cpu()
    ref = orc.rotate_aug_adjust(P.double(), Raug.double(), dataset)
    assert ours.shape == (n, 9)
    np.testing.assert_allclose(ours.numpy(), ref.numpy(), rtol=0, atol=4e-6)
    with pytest.raises(ValueError):
        rotate_aug_adjust(P.to(cuda), Raug.to(cuda), "BIWI")
