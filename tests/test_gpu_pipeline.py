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
    np.testing.assert_allclose(ours.numpy(), ref.numpy(), rtol=5e-7, atol=2e-6)     # fp32 rounding of a 3-term dot product
    with pytest.raises(ValueError):
        rotate_aug_adjust(P.to(cuda), Raug.to(cuda), "BIWI")


@pytest.mark.parametrize("kind", ["ce", "nll"])
def test_unsupervised_terms_match_masked_gather(cuda, kind):
    """the sync-free form (mask as a weight) reproduces the reference's
    gather-then-mean-times-ratio value and gradient (src/agent.py:148-166), here composed from the
    oracle's restatements on the CPU."""
    from oracle import so3_oracle as orc
    from semiuhpe_b200.agent import unsupervised_terms
    gen = torch.Generator().manual_seed(11)
    b = 96
    weak = 12 * torch.randn(b, 9, generator=gen)
    strong = weak + 2 * torch.randn(b, 9, generator=gen)
    aug = random_rotations(b, gen)
    gt = random_rotations(b, gen)
    thres = -3.2
    # reference composition
    leaf = strong.clone().requires_grad_(True)
    ent = orc.fisher_entropy(weak)
    m = ent < thres
    ratio = m.sum() / len(m)
    adj = orc.rotate_aug_adjust(weak, aug, "300WLP")
    pseudo = orc.a_to_r(adj[m])
    if kind == "ce":
        ref_losses = orc.fisher_ce(adj[m], leaf[m])
    else:
        ref_losses, _ = orc.vmf_loss(leaf[m], pseudo, overreg=1.025)
    ref_loss = ref_losses.mean() * ratio
    ref_loss.backward()
    assert 0 < int(m.sum()) < b
    dev = strong.to(cuda).requires_grad_(True)
    out = unsupervised_terms(weak.to(cuda), dev, thres, type_unsuper=kind, aug_rot_mat=aug.to(cuda),
                             train_labeled="300WLP", ulb_gt=gt.to(cuda))
    out["unsuper_loss"].backward()
    assert torch.equal(out["mask"].cpu(), m)
    np.testing.assert_allclose(out["mask_ratio"].item(), ratio.item(), rtol=1e-6)
    np.testing.assert_allclose(out["unsuper_loss"].item(), ref_loss.item(), rtol=2e-5, atol=1e-5)
    g_ref = leaf.grad.numpy()
    scale = np.abs(g_ref).max()
    assert np.abs(dev.grad.cpu().numpy() - g_ref).max() < 1e-4 * scale
    assert float(dev.grad[~out["mask"]].abs().max()) == 0.0
    err_ref = orc.geodesic_deg(orc.a_to_r(leaf.detach()[m]), pseudo)
    np.testing.assert_allclose(out["err_strongSuper_pseudo"].cpu().numpy()[m.numpy()], err_ref.numpy(), rtol=1e-4, atol=5e-3)


def _toy_net(seed):
    """Conv + BatchNorm + a few odd-sized Linear layers: fp32 parameters, BN running statistics and an
    int64 ``num_batches_tracked`` counter, tensor sizes from 1 to > one CTA chunk, unaligned storage offsets."""
    torch.manual_seed(seed)
    net = torch.nn.Sequential(
        torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Flatten(),
        torch.nn.Linear(8, 257), torch.nn.Linear(257, 301), torch.nn.Linear(301, 9), torch.nn.Linear(9, 1))
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_()
                m.running_var.uniform_(0.5, 2.0)
                m.num_batches_tracked.fill_(seed + 7)
    return net


@pytest.mark.gpu
@pytest.mark.parametrize("eman", [False, True])
@pytest.mark.parametrize("is_ema,step", [(True, 0), (True, 3), (True, 100000), (False, 50)])
def test_update_ema_variables(cuda, eman, is_ema, step):
    """SSLAgent.update_ema_variables (src/agent.py:277-299): both branches, the warm-up rule
    and the is_ema=False reset, against the line-by-line torch restatement.  The blend is two products and
    a sum (EMAN) or a product and a fused multiply-add (parameters); ATen's CPU and CUDA kernels fuse the
    same way, so the results agree to the last bit on FMA hardware -- asserted to 1 ulp."""
    import copy
    from oracle import so3_oracle as orc
    from semiuhpe_b200.agent import update_ema_variables
    net, ema = _toy_net(1), _toy_net(2)
    ref_net, ref_ema = copy.deepcopy(net), copy.deepcopy(ema)
    a_ref = orc.update_ema_variables(ref_net, ref_ema, is_ema, 0.999, step, eman=eman)
    net_d, ema_d = copy.deepcopy(net).to(cuda), copy.deepcopy(ema).to(cuda)
    a_ours = update_ema_variables(net_d, ema_d, is_ema, 0.999, step, eman=eman)
    assert a_ours == a_ref
    for (k, v_ref), (k2, v) in zip(ref_ema.state_dict().items(), ema_d.state_dict().items()):
        assert k == k2
        if v_ref.dtype == torch.float32:
            np.testing.assert_allclose(v.cpu().numpy(), v_ref.numpy(), rtol=1.2e-7, atol=1e-38, err_msg=k)
        else:
            assert torch.equal(v.cpu(), v_ref), k
    # the student is never touched
    for (k, v_ref), (_, v) in zip(net.state_dict().items(), net_d.state_dict().items()):
        assert torch.equal(v.cpu(), v_ref), k


@pytest.mark.gpu
def test_ema_update_many_tensors_and_big_tensor(cuda):
    """More tensors than one launch carries (48), more CTAs than one launch carries (320 chunks of 32768), a tensor
    that straddles launches, empty tensors, an unaligned view."""
    from semiuhpe_b200 import _ops
    gen = torch.Generator(device=cuda).manual_seed(3)
    sizes = [0, 1, 3, 5, 32768, 32769, 100003] + [17 + i for i in range(60)] + [320 * 32768 + 12345]
    base_e = [torch.randn(n + 1, device=cuda, generator=gen) for n in sizes]
    base_s = [torch.randn(n + 1, device=cuda, generator=gen) for n in sizes]
    ema = [b[1:] if i % 2 else b[:-1] for i, b in enumerate(base_e)]          # odd ones start 4 bytes off a 16-byte boundary
    src = [b[1:] if i % 3 == 0 else b[:-1] for i, b in enumerate(base_s)]
    alpha = 0.97
    a32, o32 = np.float32(alpha), np.float32(1.0 - alpha)
    for mode in (0, 1):
        want = []
        for e, s in zip(ema, src):
            e64, s64 = e.double(), s.double()
            prod = (e64 * float(a32)).float().double()                         # fl32(ema * alpha)
            if mode == 0:
                want.append((prod + (s64 * float(o32)).float().double()).float())   # fl32(fl32 + fl32): exact in double before rounding
            else:
                want.append((prod + s64 * float(o32)).float())                  # fma: one rounding (double holds the exact product)
        _ops.ema_update(ema, src, alpha, mode)
        for i, (e, w) in enumerate(zip(ema, want)):
            # bit-identical up to the (2^-29 per element) double-rounding cases of the float64 emulation above
            assert int((e != w).sum()) <= 2 and torch.allclose(e, w, rtol=1.2e-7, atol=0), (mode, sizes[i])
    with pytest.raises(TypeError):
        _ops.ema_update([ema[1].double()], [src[1].double()], 0.5, 0)
    with pytest.raises(RuntimeError):
        _ops.ema_update(ema[:2], src[:1], 0.5, 0)
