"""K3: k-th smallest entropy + keep mask, bit-exact against numpy.sort semantics."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def np_threshold(e, ratio):
    s = np.sort(e)
    return s[int(len(e) * ratio)]


def same(a, b):
    return (np.isnan(a) and np.isnan(b)) or float(a) == float(b)


def test_golden_thresholds_bit_exact(cuda, golden):
    from semiuhpe_b200.agent import entropy_threshold, entropy_mask
    g = golden("select")
    for pool, thr in ((g["entropy"], g["thresholds"]), (g["ties"], g["ties_thresholds"])):
        dev = torch.from_numpy(pool).to(cuda)
        for r, t in zip(g["ratios"], thr):
            got = entropy_threshold(dev, float(r))
            assert same(got, t), (r, got, t)
            mask, ratio = entropy_mask(dev, got)
            want = pool < t
            assert np.array_equal(mask.cpu().numpy(), want)
            assert abs(ratio.item() - want.sum() / len(pool)) < 1e-7
    with pytest.raises(IndexError):
        entropy_threshold(torch.from_numpy(g["entropy"]).to(cuda), 1.0)


def test_teacher_batch_filter_config2(cuda, golden):
    """BASELINE config 2: 128 unlabeled teacher outputs, left_ratio=0.95 -> k=121 kept."""
    from semiuhpe_b200.agent import dynamic_entropy_filter
    from oracle import so3_oracle as orc
    g = golden("select")
    A = torch.from_numpy(g["A"][:128]).to(cuda)
    ent, mask, ratio, thr = dynamic_entropy_filter(A.reshape(-1, 9), 0.95)
    e = ent.cpu().numpy()
    assert same(thr, np_threshold(e, 0.95))                       # exact on our own entropies
    assert np.array_equal(mask.cpu().numpy(), e < thr) and int(mask.sum()) == 121
    assert abs(ratio.item() - 121 / 128) < 1e-7
    # and the decision agrees with the reference's entropies (no near-ties in this batch)
    assert np.array_equal(mask.cpu().numpy(), g["mask128"])
    np.testing.assert_allclose(thr, float(g["thr128"]), rtol=1e-5)


@pytest.mark.parametrize("n", [1, 2, 3, 5, 127, 1000, 4097, 65537, 1 << 20])
def test_random_pools_all_ranks(cuda, n):
    from semiuhpe_b200.agent import entropy_threshold
    rng = np.random.default_rng(n)
    e = rng.normal(-5, 1.5, n).astype(np.float32)
    if n > 100:
        e[rng.integers(0, n, n // 50)] = e[0]                    # heavy ties
    dev = torch.from_numpy(e).to(cuda)
    s = np.sort(e)
    for ratio in (0.0, 0.25, 0.5, 0.95, 0.999999):
        k = int(n * ratio)
        assert entropy_threshold(dev, ratio) == float(s[k])
    if n > 8:
        off = torch.from_numpy(np.concatenate([[0.0], e]).astype(np.float32)).to(cuda)[1:]   # 4-byte aligned only
        assert entropy_threshold(off, 0.5) == float(s[int(n * 0.5)])


def test_special_values(cuda):
    from semiuhpe_b200.agent import entropy_threshold, entropy_mask
    e = np.array([np.nan, -np.inf, np.inf, -0.0, 0.0, 1e-45, -1e-45, 3.0, -3.0, np.nan], np.float32)
    dev = torch.from_numpy(e).to(cuda)
    s = np.sort(e)
    for k in range(len(e)):
        got = entropy_threshold(dev, (k + 0.5) / len(e))
        assert same(got, s[k]), (k, got, s[k])
    mask, _ = entropy_mask(dev, float("nan"))
    assert not mask.any()
    mask, _ = entropy_mask(dev, 0.0)
    assert np.array_equal(mask.cpu().numpy(), e < 0.0)


def test_full_pool_2pow26(cuda):
    """BASELINE config 5 pool size on one GPU: threshold index and mask exact vs a GPU sort."""
    from semiuhpe_b200.agent import entropy_threshold, entropy_mask, pool_index
    n = 1 << 26
    gen = torch.Generator(device=cuda).manual_seed(0)
    e = torch.randn(n, device=cuda, generator=gen) * 1.3 - 5.0
    k = pool_index(n, 0.95)
    assert k == 63753420
    thr = entropy_threshold(e, 0.95)
    want = torch.sort(e)[0][k].item()
    assert thr == want
    mask, ratio = entropy_mask(e, thr)
    assert torch.equal(mask, e < want)
    assert int(mask.sum()) <= k                                   # == k when there are no ties at the threshold
    # idempotence: the threshold of the kept set at ratio->max is below thr
    assert e[mask].max().item() < thr


def test_single_process_distributed_entry(cuda, golden):
    from semiuhpe_b200.distributed import global_entropy_threshold
    pool = golden("select")["entropy"]
    got = global_entropy_threshold(torch.from_numpy(pool).to(cuda), 0.95)
    assert got == float(np_threshold(pool, 0.95))
