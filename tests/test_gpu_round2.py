"""Round-2 additions on the GPU: the one-call SSL loss head (suhpe_ssl_step_f32), masked backward, row scaling,
the sync-free validation slice, CUDA-graph capture, per-call cut_bits, status-word hygiene, back-to-back host
pipeline calls and the all-gathered radix select over NCCL."""
import os

import numpy as np
import pytest
import torch

from helpers import ATOL, RTOL, assert_close, grad_rel_err, random_rotations

pytestmark = pytest.mark.gpu


def _batch(gen, b_l=32, b_u=128, scale=10.0):
    A_l = scale * torch.randn(b_l, 9, generator=gen)
    R_l = random_rotations(b_l, gen)
    W = scale * torch.randn(b_u, 9, generator=gen)
    S = W + 0.1 * scale * torch.randn(b_u, 9, generator=gen)
    return A_l, R_l, W, S


@pytest.mark.parametrize("unsup", ["ce", "nll"])
@pytest.mark.parametrize("aug", [None, "DAD3DHeads", "300WLP"])
def test_ssl_loss_matches_reference_forward_backward(cuda, unsup, aug):
    """ONE C call (supervised NLL + teacher entropy + mask + adjustment + CE/NLL + means + gradients) against the
    oracle's restatement of SSLAgent.forward / train_func (src/agent.py:76-83,99-166,203) with boolean gathers."""
    from oracle import so3_oracle as orc
    from semiuhpe_b200.agent import ssl_loss
    gen = torch.Generator().manual_seed(11)
    A_l, R_l, W, S = _batch(gen)
    aug_rot = random_rotations(len(W), gen) if aug else None
    lam, thr = 0.7, -3.9
    # ---- reference semantics on the CPU (oracle)
    ref_l = A_l.clone().requires_grad_(True)
    ref_s = S.clone().requires_grad_(True)
    losses, pred_orth = orc.vmf_loss(ref_l, R_l, overreg=1.025)
    ent = orc.fisher_entropy(W)
    mask, ratio = orc.keep_mask(ent, thr)
    adjusted = W if aug is None else orc.rotate_aug_adjust(W, aug_rot, aug)
    assert 0 < int(mask.sum()) < len(W)
    if unsup == "ce":
        ul = orc.fisher_ce(adjusted[mask], ref_s[mask])
    else:
        ul, _ = orc.vmf_loss(ref_s[mask], orc.a_to_r(adjusted[mask]), overreg=1.025)
    ref_all = losses.mean() + lam * (ul.mean() * ratio)
    ref_all.backward()
    # ---- ours
    out_l = A_l.to(cuda).requires_grad_(True)
    out_s = S.to(cuda).requires_grad_(True)
    loss_all, info = ssl_loss(out_l, R_l.to(cuda), W.to(cuda), out_s, thr, SSL_lambda=lam, type_unsuper=unsup,
                              aug_rot_mat=None if aug is None else aug_rot.to(cuda), train_labeled=aug or "300WLP")
    loss_all.backward()
    assert_close(info["entropy"].cpu().numpy(), ent.numpy(), RTOL, 2 * ATOL, "entropy")
    assert torch.equal(info["mask"].cpu(), info["entropy"].cpu() < thr)
    assert abs(info["mask_ratio"].item() - float(info["mask"].float().mean())) < 1e-7
    assert_close(info["losses"].cpu().numpy(), losses.detach().numpy(), RTOL, ATOL, "supervised losses")
    assert_close([info["loss"].item()], [losses.mean().item()], RTOL, ATOL, "loss")
    assert_close([loss_all.item()], [ref_all.item()], 3 * RTOL, 3 * ATOL, "loss_all")
    assert torch.allclose(info["pred_orth"].cpu(), pred_orth, atol=2e-5)
    assert torch.allclose(info["pseudo_labels"].cpu(), orc.a_to_r(adjusted), atol=2e-5)
    if torch.equal(info["mask"].cpu(), mask):
        kept = mask.numpy()
        assert_close(info["unsuper_losses"].cpu().numpy()[kept], ul.detach().numpy(), 5 * RTOL, 5 * ATOL, "unsupervised losses")
        assert float(info["unsuper_losses"].cpu()[~mask].abs().max()) == 0.0
        assert grad_rel_err(out_l.grad.cpu().numpy(), ref_l.grad.numpy()).max() < 1e-4
        gs, rs = out_s.grad.cpu().numpy(), ref_s.grad.numpy()
        assert np.all(gs[~kept] == 0.0)
        scale = np.abs(rs).max()
        assert np.abs(gs - rs).max() <= 2e-4 * scale, np.abs(gs - rs).max() / scale


def test_ssl_loss_equals_the_mirror_composition(cuda):
    """Same numbers as the per-function mirrors (vmf_loss + unsupervised_terms), to rounding of the final sums."""
    from semiuhpe_b200.agent import ssl_loss, unsupervised_terms
    from semiuhpe_b200.fisher.fisher_utils import vmf_loss
    gen = torch.Generator().manual_seed(3)
    A_l, R_l, W, S = (t.to(cuda) for t in _batch(gen))
    l1, s1 = A_l.clone().requires_grad_(True), S.clone().requires_grad_(True)
    loss, _ = vmf_loss(l1, R_l, overreg=1.025)
    un = unsupervised_terms(W, s1, -4.0, type_unsuper="ce")
    (loss.mean() + 0.5 * un["unsuper_loss"]).backward()
    l2, s2 = A_l.clone().requires_grad_(True), S.clone().requires_grad_(True)
    loss_all, info = ssl_loss(l2, R_l, W, s2, -4.0, SSL_lambda=0.5)
    loss_all.backward()
    assert torch.equal(info["mask"], un["mask"]) and torch.equal(info["entropy"], un["entropy"])
    assert torch.allclose(loss_all, loss.mean() + 0.5 * un["unsuper_loss"], rtol=1e-6, atol=1e-6)
    assert torch.allclose(l2.grad, l1.grad, rtol=1e-6, atol=1e-9)
    assert torch.allclose(s2.grad, s1.grad, rtol=1e-5, atol=1e-9)


def test_supervised_only_step_and_no_grad(cuda):
    """train_func_s1 (src/agent.py:253-270): no unlabeled batch; and a no_grad call (validation) skips the gradients."""
    from oracle import so3_oracle as orc
    from semiuhpe_b200.agent import ssl_loss
    gen = torch.Generator().manual_seed(4)
    A_l, R_l, _, _ = _batch(gen, b_l=57)
    leaf = A_l.to(cuda).requires_grad_(True)
    loss_all, info = ssl_loss(leaf, R_l.to(cuda))
    loss_all.backward()
    ref = A_l.clone().requires_grad_(True)
    rl, _ = orc.vmf_loss(ref, R_l, overreg=1.025)
    rl.mean().backward()
    assert_close([loss_all.item()], [rl.mean().item()], RTOL, ATOL, "loss")
    assert grad_rel_err(leaf.grad.cpu().numpy(), ref.grad.numpy()).max() < 1e-4
    assert info["entropy"] is None or info["entropy"].numel() == 0
    with torch.no_grad():
        l2, _ = ssl_loss(A_l.to(cuda), R_l.to(cuda))
    assert torch.allclose(l2, loss_all.detach(), rtol=1e-6, atol=1e-6)


def test_filtered_nan_rows_never_reach_the_gradient(cuda):
    """a filtered row holding NaN/Inf must give an exactly zero gradient (0 * NaN never forms) and must
    not trip the NaN assert, which the reference applies to the gathered rows only (fisher_utils.py:98)."""
    import semiuhpe_b200
    from semiuhpe_b200.agent import unsupervised_terms, ssl_loss
    from semiuhpe_b200.fisher.fisher_utils import fisher_CE
    gen = torch.Generator().manual_seed(8)
    _, _, W, S = _batch(gen)
    keep = torch.ones(len(W), dtype=torch.bool)
    keep[[3, 77]] = False
    S_bad = S.clone()
    S_bad[3] = float("nan")
    S_bad[77, 4] = float("inf")
    leaf = S_bad.to(cuda).requires_grad_(True)
    ce = fisher_CE(W.to(cuda), leaf, keep=keep.to(cuda))          # error checking ON: must not raise
    ce.sum().backward()
    assert float(ce[3]) == 0.0 and float(ce[77]) == 0.0 and bool(torch.isfinite(ce).all())
    g = leaf.grad.cpu()
    assert bool(torch.isfinite(g).all()) and float(g[3].abs().max()) == 0.0 and float(g[77].abs().max()) == 0.0
    good = S.to(cuda).requires_grad_(True)
    fisher_CE(W.to(cuda), good, keep=keep.to(cuda)).sum().backward()
    assert torch.equal(good.grad.cpu()[keep], g[keep])
    with pytest.raises(torch.linalg.LinAlgError):                 # an UNfiltered NaN row still raises like the reference's torch.svd
        fisher_CE(W.to(cuda), S_bad.to(cuda))
    # the sync-free step: a teacher row with NaN is filtered by its own NaN entropy (NaN < thr is false)
    semiuhpe_b200.set_error_checking(False)
    try:
        W_bad = W.clone()
        W_bad[5] = float("nan")
        s = S.to(cuda).requires_grad_(True)
        out = unsupervised_terms(W_bad.to(cuda), s, -3.0, type_unsuper="ce")
        out["unsuper_loss"].backward()
        assert not bool(out["mask"][5]) and bool(torch.isfinite(s.grad).all()) and bool(torch.isfinite(out["unsuper_loss"]))
        s2 = S.to(cuda).requires_grad_(True)
        la, info = ssl_loss(10 * torch.randn(8, 9, device=cuda), random_rotations(8).to(cuda), W_bad.to(cuda), s2, -3.0)
        la.backward()
        assert not bool(info["mask"][5]) and bool(torch.isfinite(s2.grad).all()) and bool(torch.isfinite(la))
        torch.cuda.synchronize()
    finally:
        semiuhpe_b200.set_error_checking(True)                    # clears the bits the unchecked launches left behind
    from semiuhpe_b200.fisher.fisher_utils import fisher_entropy
    fisher_entropy(W.to(cuda))                                    # must not be blamed for the NaN rows above


def test_scale_rows(cuda):
    from semiuhpe_b200 import _ops
    gen = torch.Generator().manual_seed(1)
    for n, w in ((1, 9), (1000, 9), (4099, 3)):
        rows = torch.randn(n, w, generator=gen).to(cuda)
        rw = torch.randn(n, generator=gen).to(cuda)
        sw = torch.tensor(0.37, device=cuda)
        keep = (torch.rand(n, generator=gen) < 0.7).to(cuda)
        rows_nan = rows.clone()
        rows_nan[~keep] = float("nan")
        out = _ops.scale_rows(rows_nan, rw, sw, keep)
        ref = torch.where(keep[:, None], rows * rw[:, None] * sw, torch.zeros((), device=cuda))
        assert torch.equal(out, ref)
        assert torch.equal(_ops.scale_rows(rows, rw), rows * rw[:, None])
        assert torch.equal(_ops.scale_rows(rows, scalar_weight=sw), rows * sw)
        assert torch.equal(_ops.scale_rows(rows, torch.tensor([2.0], device=cuda)), rows * 2.0)


def test_validation_terms_match_gather_path(cuda):
    """src/agent.py:224-249: mask as a weight == gather-then-mean."""
    from semiuhpe_b200.agent import validation_terms, compute_err_deg_from_matrices
    from semiuhpe_b200.fisher.fisher_utils import batch_torch_A_to_R, fisher_entropy
    gen = torch.Generator().manual_seed(21)
    b = 300
    pred = (10 * torch.randn(b, 9, generator=gen)).to(cuda)
    gt = random_rotations(b, gen).to(cuda)
    gt_euler = ((torch.rand(b, 3, generator=gen) * 2 - 1) * 80).to(cuda)
    orth = batch_torch_A_to_R(pred)
    for ge in (None, gt_euler):
        out = validation_terms(pred, orth, gt, -4.0, gt_euler=ge)
        ent = fisher_entropy(pred)
        mask = ent < -4.0
        assert torch.equal(out["mask"], mask) and 0 < int(mask.sum()) < b
        ref = compute_err_deg_from_matrices(orth[mask], gt[mask], gt_euler=None if ge is None else ge[mask])
        assert torch.equal(out["err_all"][mask], ref)
        assert torch.allclose(out["err_pseudo_gt_mean"], ref.mean(), rtol=1e-6)
        assert abs(out["mask_ratio"].item() - float(mask.float().mean())) < 1e-7


def test_cuda_graph_capture_of_the_ssl_step(cuda):
    """The sync-free steps are fixed launch sequences: capture once, replay on new data, same numbers as eager."""
    import semiuhpe_b200
    from semiuhpe_b200.agent import ssl_loss, unsupervised_terms
    gen = torch.Generator().manual_seed(31)
    A_l, R_l, W, S = (t.to(cuda) for t in _batch(gen))
    sA, sR, sW, sS = (torch.empty_like(t) for t in (A_l, R_l, W, S))      # static graph inputs
    leaf_l, leaf_s = sA.requires_grad_(True), sS.requires_grad_(True)
    outs = {}

    def step_one_call():
        leaf_l.grad = None
        leaf_s.grad = None
        la, info = ssl_loss(leaf_l, sR, sW, leaf_s, -4.0, SSL_lambda=0.3)
        la.backward()
        outs["one"] = (la.detach(), info["mask"])

    def step_mirrors():
        leaf_s.grad = None
        o = unsupervised_terms(sW, leaf_s, -4.0, type_unsuper="ce")
        o["unsuper_loss"].backward()
        outs["mir"] = (o["unsuper_loss"].detach(), o["mask"], o["err_strongSuper_pseudo_mean"])

    semiuhpe_b200.set_error_checking(False)
    try:
        for name, step in (("one", step_one_call), ("mir", step_mirrors)):
            with torch.no_grad():
                for dst, src in ((sA, A_l), (sR, R_l), (sW, W), (sS, S)):
                    dst.copy_(src)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step()                                                # warm-up on the capture stream
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            eager = [t.clone() for t in outs[name]] + [leaf_s.grad.clone()]
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                step()
            graph_out, graph_grad = outs[name], leaf_s.grad
            g.replay()
            torch.cuda.synchronize()
            for a, b in zip(eager, list(graph_out) + [graph_grad]):
                assert torch.equal(a, b), name
            # new data through the same graph
            with torch.no_grad():
                sW.copy_(W.flip(0)); sS.copy_(S.flip(0))
            g.replay()
            torch.cuda.synchronize()
            replayed = [t.clone() for t in graph_out] + [graph_grad.clone()]
            step()
            torch.cuda.synchronize()
            for a, b in zip(replayed, list(outs[name]) + [leaf_s.grad]):
                assert torch.equal(a, b), name
    finally:
        semiuhpe_b200.set_error_checking(True)


def test_cut_bits_is_a_per_call_argument(cuda):
    """No process-wide setting in the library: two calls with different cut_bits do not interact."""
    import semiuhpe_b200
    from semiuhpe_b200 import _ops
    A = (10 * torch.randn(5000, 9, generator=torch.Generator().manual_seed(2))).to(cuda)
    d26 = _ops.fisher_fused(A, None, 1.0, nll=True, entropy=True, cut_bits=26)
    d0 = _ops.fisher_fused(A, None, 1.0, nll=True, entropy=True, cut_bits=0)
    again = _ops.fisher_fused(A, None, 1.0, nll=True, entropy=True, cut_bits=26)
    assert torch.equal(d26["nll"], again["nll"]) and torch.equal(d26["entropy"], again["entropy"])
    assert_close(d26["nll"].cpu().numpy(), d0["nll"].cpu().numpy(), 2e-6, 2e-6, "cut on/off")
    assert semiuhpe_b200.quadrature_cut_bits() == 26
    prev = semiuhpe_b200.set_quadrature_cut_bits(0)
    try:
        assert torch.equal(_ops.fisher_fused(A, None, 1.0, nll=True)["nll"], _ops.fisher_fused(A, None, 1.0, nll=True, cut_bits=0)["nll"])
    finally:
        semiuhpe_b200.set_quadrature_cut_bits(prev)


def test_misshaped_records_are_rejected(cuda):
    from semiuhpe_b200.fisher.fisher_utils import fisher_entropy, batch_torch_A_to_R
    with pytest.raises(RuntimeError, match="trailing dimensions"):
        fisher_entropy(torch.randn(9, 3, device=cuda))               # (n,3) with n % 3 == 0 must not become 3 matrices
    with pytest.raises(RuntimeError, match="trailing dimensions"):
        batch_torch_A_to_R(torch.randn(18, device=cuda))
    assert fisher_entropy(torch.randn(2, 2, 3, 3, device=cuda)).shape == (4,)


def test_status_bits_of_unchecked_launches_are_not_blamed_later(cuda):
    import semiuhpe_b200
    from semiuhpe_b200.fisher.fisher_utils import fisher_entropy
    bad = torch.full((4, 9), float("nan"), device=cuda)
    semiuhpe_b200.set_error_checking(False)
    fisher_entropy(bad)
    torch.cuda.synchronize()
    semiuhpe_b200.set_error_checking(True)
    fisher_entropy(torch.randn(4, 9, device=cuda))                   # clean input: no stale LinAlgError
    with pytest.raises(torch.linalg.LinAlgError):
        fisher_entropy(bad)
    # a second stream has its own word
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        fisher_entropy(torch.randn(4, 9, device=cuda))


def test_pipeline_calls_back_to_back_without_sync(cuda):
    """suhpe_fisher_pool_host twice on one pipeline with no suhpe_pipeline_sync in between -- the second
    call must not overwrite chunk buffers the first call's kernels / copies still use."""
    import ctypes
    from semiuhpe_b200 import _capi, _ops
    n, chunk = 300000, 8192
    gen = torch.Generator().manual_seed(9)
    hA = [(10 * torch.randn(n, 9, generator=gen)).pin_memory() for _ in range(2)]
    hR = [random_rotations(n, gen).reshape(n, 9).pin_memory() for _ in range(2)]
    lib = _capi.lib()
    h = ctypes.c_void_p()
    _capi.check(lib.suhpe_pipeline_create(ctypes.byref(h), n, chunk, 26), "create")
    outs = [dict(nll=torch.empty(n).pin_memory(), grad=torch.empty(n, 9).pin_memory(), ent=torch.empty(n).pin_memory()) for _ in range(2)]
    dent = [torch.empty(n, device=cuda) for _ in range(2)]
    hist = torch.zeros(2048, dtype=torch.int64, device=cuda)
    status = torch.zeros(1, dtype=torch.int32, device=cuda)
    P = lambda t: t.data_ptr()
    for i in range(2):
        _capi.check(lib.suhpe_fisher_pool_host(h, P(hA[i]), P(hR[i]), n, 1.025, P(outs[i]["nll"]), P(outs[i]["grad"]), P(outs[i]["ent"]),
                                               P(dent[i]), P(hist), P(status), _capi.stream()), "pool")
    _capi.check(lib.suhpe_pipeline_sync(h), "sync")
    torch.cuda.synchronize()
    for i in range(2):
        ref = _ops.fisher_fused(hA[i].to(cuda), hR[i].to(cuda), 1.025, nll=True, grad=True, entropy=True)
        assert torch.equal(outs[i]["nll"], ref["nll"].cpu()), i
        assert torch.equal(outs[i]["grad"], ref["grad"].cpu()), i
        assert torch.equal(outs[i]["ent"], ref["entropy"].cpu()) and torch.equal(dent[i], ref["entropy"]), i
    lib.suhpe_pipeline_destroy(h)


def test_fused_first_histogram_is_counted_in_the_kernel(cuda):
    """K2's `hist` output == the stand-alone first radix pass over the entropies it wrote (ragged n, NaN rows)."""
    from semiuhpe_b200 import _capi, _ops
    import semiuhpe_b200
    n = 100003
    A = (10 * torch.randn(n, 9, generator=torch.Generator().manual_seed(6))).to(cuda)
    A[17] = float("nan")
    semiuhpe_b200.set_error_checking(False)
    try:
        hist = torch.zeros(2048, dtype=torch.int64, device=cuda)
        ent = _ops.fisher_fused(A, None, 1.0, entropy=True, hist=hist)["entropy"]
        ws = _ops.SelectWorkspace(cuda)
        lib = _capi.lib()
        _capi.check(lib.suhpe_select_init(_capi.ptr(ws.state), 0, _capi.stream()), "init")
        _capi.check(lib.suhpe_select_hist_f32(_capi.ptr(ent), n, 1, _capi.ptr(ws.state), _capi.ptr(ws.hist[1]), _capi.stream()), "hist")
        assert torch.equal(hist, ws.hist[1]) and int(hist.sum()) == n
        only_hist = torch.zeros(2048, dtype=torch.int64, device=cuda)       # hist without the entropy vector
        _capi.check(lib.suhpe_fisher_fused_f32(_capi.ptr(A), None, n, 1.0, 26, None, None, None, None, None, None, None,
                                               _capi.ptr(only_hist), None, _capi.stream()), "fused")
        assert torch.equal(only_hist, hist)
        torch.cuda.synchronize()
    finally:
        semiuhpe_b200.set_error_checking(True)


def _nccl_worker(rank, world, port, n_per, ratio, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import semiuhpe_b200
        from semiuhpe_b200 import _ops
        from semiuhpe_b200.agent import entropy_mask, pool_index
        from semiuhpe_b200.distributed import global_entropy_threshold, sharded_mean
        gen = torch.Generator(device=dev).manual_seed(100 + rank)
        A = 10 * torch.randn(n_per + 17 * rank, 9, device=dev, generator=gen)            # ragged shards
        ent = _ops.fisher_fused(A, None, 1.0, entropy=True)["entropy"]
        thr = global_entropy_threshold(ent, ratio)
        mask, _ = entropy_mask(ent, thr)
        sizes = [n_per + 17 * r for r in range(world)]
        pool = [torch.empty(s, device=dev) for s in sizes]
        dist.all_gather(pool, ent)
        flat = torch.cat(pool)
        k = pool_index(flat.numel(), ratio)
        ref = torch.sort(flat).values[k]
        kept = torch.tensor([int(mask.sum())], device=dev)
        dist.all_reduce(kept)
        mean = sharded_mean(ent)
        ok = (thr == float(ref)) and int(kept) == int((flat < ref).sum()) and \
            abs(float(mean) - float(flat.double().mean())) < 1e-6 * abs(float(flat.double().mean()))
        ret[rank] = (bool(ok), thr, float(ref))
    finally:
        dist.destroy_process_group()


def test_global_threshold_over_nccl_matches_sort_of_the_concatenated_pool(cuda):
    """src/agent.py:403-407 on a pool sharded over 2 GPUs: CUDA histograms + NCCL all-gather + identical scans give
    sort(concat)[k] on every rank, the masks keep count(e < thr) rows, the sharded mean equals the global mean."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the driver's single-GPU tier skips this; run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_nccl_worker, args=(2, port, 1 << 18, 0.95, ret), nprocs=2, join=True)
    assert len(ret) == 2 and all(v[0] for v in ret.values()), dict(ret)
    assert ret[0][1] == ret[1][1]


def test_err_deg_from_quats_and_formula_known_answers(cuda):
    """src/agent.py:420-424 (real-first quaternions -> geodesic degrees) against the oracle, and K4's geodesic angle
    against the formula-derived known answers of tests/golden/pytorch3d_kat.json (acos extension at +-(1-1e-4), the
    0.40514 deg identity, the trace check).  a16 stays parity-unpinned: these are not pytorch3d outputs."""
    import json
    from oracle import pytorch3d_restated as p3d, so3_oracle as orc
    from semiuhpe_b200.agent import compute_err_deg_from_quats, compute_err_deg_from_matrices
    gen = torch.Generator().manual_seed(13)
    q1 = torch.randn(500, 4, generator=gen)
    q2 = torch.randn(500, 4, generator=gen) * 3.0                     # non-unit: two_s = 2/|q|^2 handles it
    want = orc.geodesic_deg(p3d.quaternion_to_matrix(q1), p3d.quaternion_to_matrix(q2))
    got = compute_err_deg_from_quats(q1.to(cuda), q2.to(cuda)).cpu()
    assert torch.allclose(got, want, rtol=1e-4, atol=2e-3)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pytorch3d_kat.json")) as f:
        kat = json.load(f)
    rows = kat["so3_relative_angle_about_z"]
    R1 = torch.tensor([r["R1"] for r in rows], dtype=torch.float32).to(cuda)
    I = torch.eye(3).expand(len(rows), 3, 3).contiguous().to(cuda)
    deg = compute_err_deg_from_matrices(R1, I).cpu().double()
    want = torch.tensor([np.degrees(r["angle"]) for r in rows])
    # fp32 next to the bound: d acos/dx ~ 70, x carries ~6e-8 of rounding -> ~3e-4 deg
    assert (deg - want).abs().max() < 2e-3, (deg - want).abs().max()
    assert abs(float(deg[0]) - kat["identity_angle_deg"]) < 1e-4
    for M in kat["trace_out_of_range"]:
        with pytest.raises(ValueError):
            compute_err_deg_from_matrices(torch.tensor(M, dtype=torch.float32)[None].to(cuda), torch.eye(3)[None].to(cuda))
    for M in kat["trace_in_range_edge"]:
        compute_err_deg_from_matrices(torch.tensor(M, dtype=torch.float32)[None].to(cuda), torch.eye(3)[None].to(cuda))


def test_batch_torch_A_to_R_is_differentiable_like_the_reference(cuda):
    """src/fisher/fisher_utils.py:39-48 back-propagates through torch.svd; ours: the closed-form polar gradient."""
    from oracle import so3_oracle as orc
    from semiuhpe_b200.fisher.fisher_utils import batch_torch_A_to_R
    gen = torch.Generator().manual_seed(17)
    A = torch.randn(400, 3, 3, generator=gen) * torch.tensor([0.3, 3.0, 10.0, 30.0]).repeat_interleave(100)[:, None, None]
    W = torch.randn(400, 3, 3, generator=gen)
    ref = A.double().clone().requires_grad_(True)
    (orc.a_to_r(ref) * W.double()).sum().backward()
    leaf = A.to(cuda).requires_grad_(True)
    R = batch_torch_A_to_R(leaf)
    assert R.requires_grad
    (R * W.to(cuda)).sum().backward()
    g, r = leaf.grad.cpu().double(), ref.grad
    scale = r.abs().amax(dim=(1, 2)).clamp(min=1e-12)
    rel = (g - r).abs().amax(dim=(1, 2)) / scale
    # 1/(s_i + s_j) amplifies fp32 rounding when s2 + s3 is small against s1 (det < 0 matrices): compare per class
    assert rel.median() < 1e-5 and rel.quantile(0.95) < 2e-4, (rel.median(), rel.quantile(0.95), rel.max())
    with torch.no_grad():
        assert not batch_torch_A_to_R(leaf).requires_grad


def test_ema_update_keeps_reference_semantics_for_non_fp32_and_strided_entries(cuda):
    """the EMAN blend must blend EVERY state_dict entry except num_batches_tracked (src/agent.py:290-293) --
    bf16 / fp64 / channels_last tensors take the reference's own torch expression on the GPU instead of being copied."""
    from oracle import so3_oracle as orc
    from semiuhpe_b200.agent import update_ema_variables
    import copy

    def make():
        torch.manual_seed(3)
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3), torch.nn.BatchNorm2d(8), torch.nn.Conv2d(8, 4, 3))
        net[2].weight.data = net[2].weight.data.to(memory_format=torch.channels_last)       # strided parameter
        net.register_buffer("half_buf", torch.randn(7).to(torch.bfloat16))
        net.register_buffer("dbl_buf", torch.randn(5, dtype=torch.float64))
        return net

    for eman in (True, False):
        student, teacher = make().to(cuda), make().to(cuda)
        with torch.no_grad():
            for p in student.parameters():
                p.add_(0.5)
            student.half_buf.add_(1.0); student.dbl_buf.add_(1.0)
            student[1].num_batches_tracked.fill_(11)
        ref_student, ref_teacher = copy.deepcopy(student).cpu(), copy.deepcopy(teacher).cpu()
        a1 = update_ema_variables(student, teacher, True, 0.999, 100, eman=eman)
        a2 = orc.update_ema_variables(ref_student, ref_teacher, True, 0.999, 100, eman=eman)
        assert a1 == a2
        for (k, v), (_, r) in zip(teacher.state_dict().items(), ref_teacher.state_dict().items()):
            if v.dtype.is_floating_point:
                assert torch.allclose(v.cpu().double(), r.double(), rtol=1e-6 if v.dtype != torch.bfloat16 else 1e-2, atol=1e-7), k
            else:
                assert torch.equal(v.cpu(), r), k
        if eman:
            assert int(teacher[1].num_batches_tracked) == 11
            assert not torch.equal(teacher.half_buf, student.half_buf)            # blended, not copied


def test_two_streams_do_not_interact(cuda):
    """Re-entrancy: the same entry points on two streams at once (each with its own status word, select workspace and
    SSL-step handle) give the results of running them one after the other."""
    import semiuhpe_b200
    from semiuhpe_b200.agent import dynamic_entropy_filter, ssl_loss
    from semiuhpe_b200 import _ops
    gen = torch.Generator().manual_seed(41)
    A = [(10 * torch.randn(200000, 9, generator=gen)).to(cuda) for _ in range(2)]
    A_l, R_l, W, S = (t.to(cuda) for t in _batch(gen))
    serial = []
    for a in A:
        ent, mask, ratio, thr = dynamic_entropy_filter(a, 0.9)
        serial.append((ent.clone(), mask.clone(), thr))
    ref_loss = ssl_loss(A_l, R_l, W, S, -4.0)[0].clone()
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    out = [None, None]
    semiuhpe_b200.set_error_checking(False)
    try:
        for rep in range(3):
            for i, st in enumerate(streams):
                st.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(st):
                    ent, mask, ratio = dynamic_entropy_filter(A[i], 0.9, return_threshold=False)
                    loss = ssl_loss(A_l, R_l, W, S, -4.0)[0]
                    out[i] = (ent, mask, loss)
        torch.cuda.synchronize()
    finally:
        semiuhpe_b200.set_error_checking(True)
    for i in range(2):
        assert torch.equal(out[i][0], serial[i][0]) and torch.equal(out[i][1], serial[i][1])
        assert torch.equal(out[i][2], ref_loss)


def test_extreme_scales_stay_finite_and_proper(cuda):
    """Parameter matrices from 1e-30 to 1e+30 (the SVD prescales by an exact power of two): rotations stay proper and
    finite, entropies of tiny matrices tend to 0 like the reference's, nothing hangs or raises."""
    from semiuhpe_b200.fisher.fisher_utils import batch_torch_A_to_R, fisher_entropy, vmf_loss
    gen = torch.Generator().manual_seed(43)
    base = torch.randn(64, 9, generator=gen)
    for scale in (1e-30, 1e-12, 1e-3, 1.0, 1e3, 1e12, 1e30):
        A = (base * scale).to(cuda)
        R = batch_torch_A_to_R(A)
        assert bool(torch.isfinite(R).all())
        eye = torch.eye(3, device=cuda).expand(64, 3, 3)
        assert float((R @ R.transpose(1, 2) - eye).abs().max()) < 1e-5
        assert float((torch.det(R) - 1).abs().max()) < 1e-5
        if scale <= 1e-3:
            ent = fisher_entropy(A)
            assert bool(torch.isfinite(ent).all()) and float(ent.abs().max()) < 1e-4     # uniform density: H -> 0
            loss, _ = vmf_loss(A, R, overreg=1.025)
            assert bool(torch.isfinite(loss).all())


def test_make_graphed_callables_on_the_ssl_step(cuda):
    """The stock PyTorch route to a graph-replayed training step (torch.cuda.make_graphed_callables: warm-up on a
    side stream, forward and backward captured separately) works on the one-call loss head: the library allocates
    nothing inside the call and its handle is per device, not per stream."""
    import semiuhpe_b200
    from semiuhpe_b200.agent import ssl_loss
    gen = torch.Generator().manual_seed(51)
    A_l, R_l, W, S = (t.to(cuda) for t in _batch(gen))

    def head(out_l, strong, gt, weak):
        return ssl_loss(out_l, gt, weak, strong, -4.0, SSL_lambda=0.3)[0]

    semiuhpe_b200.set_error_checking(False)
    try:
        sample = (A_l.clone().requires_grad_(True), S.clone().requires_grad_(True), R_l.clone(), W.clone())
        graphed = torch.cuda.make_graphed_callables(head, sample)
        for shift in (0.0, 0.25):
            l1, s1 = (A_l + shift).requires_grad_(True), (S - shift).requires_grad_(True)
            out = graphed(l1, s1, R_l, W)
            out.backward()
            l2, s2 = (A_l + shift).requires_grad_(True), (S - shift).requires_grad_(True)
            ref = head(l2, s2, R_l, W)
            ref.backward()
            torch.cuda.synchronize()
            assert torch.equal(out.detach(), ref.detach())
            assert torch.equal(l1.grad, l2.grad) and torch.equal(s1.grad, s2.grad)
    finally:
        semiuhpe_b200.set_error_checking(True)
