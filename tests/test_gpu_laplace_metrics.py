"""K2L (rotation-Laplace NLL fwd+bwd) and K4 (error metrics) parity."""
import numpy as np
import pytest
import torch

from helpers import assert_close, grad_rel_err, no_worse_than_reference, random_rotations
from oracle import so3_oracle as orc

pytestmark = pytest.mark.gpu

# The reference's own fp32 noise on this loss is 1.7e-5 relative for the NLL and ~6e-4 for the
# gradient (cancellation in sum(s) - tr(A^T R)), so parity with its fp32
# output is checked at 5e-5 / 2e-3 AND we must be no further from exact arithmetic than it is.
LAP_RTOL, LAP_ATOL, LAP_GRAD = 5e-5, 5e-5, 2e-3


def test_laplace_golden_small_batch_path(cuda, golden):
    from semiuhpe_b200.laplace.rotation_laplace import NLL_loss, analytical_mode, log_pdf
    g = golden("laplace")
    A, R, grids = (torch.from_numpy(g[k]).to(cuda) for k in ("A", "R", "grids"))
    leaf = A.clone().requires_grad_(True)
    losses, mode = NLL_loss("RLaplace", leaf, R, grids)
    losses.sum().backward()
    assert losses.shape == (len(A),) and mode.shape == (len(A), 3, 3)
    assert_close(losses.detach().cpu().numpy(), g["nll"], LAP_RTOL, LAP_ATOL, "laplace nll")
    assert grad_rel_err(leaf.grad.cpu().numpy(), g["grad"]).max() < LAP_GRAD
    A64, R64, g64 = (torch.from_numpy(g[k]).double() for k in ("A", "R", "grids"))
    l64 = A64.clone().requires_grad_(True)
    nll64, _ = orc.laplace_nll("RLaplace", l64, R64, g64)
    nll64.sum().backward()
    assert no_worse_than_reference(losses.detach().cpu().numpy(), g["nll"], nll64.detach().numpy(), 2.0, 5e-6).all()
    # gradient against exact arithmetic: over the batch we must be at least as close as the reference's
    # own fp32 result, and every row within the reference's noise level (2e-4 of the row's scale)
    g64 = l64.grad.numpy().reshape(-1, 9)
    ours_rel, ref_rel = grad_rel_err(leaf.grad.cpu().numpy(), g64), grad_rel_err(g["grad"], g64)
    assert ours_rel.max() <= ref_rel.max() and np.median(ours_rel) <= 1.5 * np.median(ref_rel)
    assert (ours_rel <= 2 * ref_rel + 2e-4).all()
    well = (g["A"].reshape(-1, 9).std(1) > 0.5)
    assert np.abs(mode.cpu().numpy() - g["mode"])[well].max() < 2e-5
    m2, s3 = analytical_mode(A, "RLaplace")
    # K2L derives its mode from the fp64 per-sample set-up, K1 (analytical_mode) works in fp32: same rotation to rounding
    assert (m2 - mode).abs().max().item() < 1e-6 and torch.equal(s3.cpu(), torch.sign(torch.from_numpy(g["s3sign"])))
    assert_close(log_pdf("RFisher", A, R, grids).cpu().numpy(), g["rfisher_logpdf"], 2e-5, 2e-5, "RFisher grid pdf")
    dens = log_pdf("RLaplace", A[:4], grids, grids)
    want = orc.grid_log_pdf("RLaplace", torch.from_numpy(g["A"][:4]), torch.from_numpy(g["grids"]), torch.from_numpy(g["grids"]))
    assert dens.shape == (4, grids.shape[0])
    assert_close(dens.cpu().numpy(), want.numpy(), 2e-4, 2e-4, "broadcast density")


def test_laplace_large_batch_path_matches_small(cuda, golden):
    """>= 148*256 samples switch to the thread-per-sample decomposition; both decompositions
    and a multi-chunk grid (N > 4608) must agree with each other and with the oracle."""
    from semiuhpe_b200 import _ops
    g = golden("laplace")
    grids = torch.from_numpy(g["grids"]).to(cuda)
    n = 148 * 256 + 77
    gen = torch.Generator().manual_seed(2)
    A = (5 * torch.randn(n, 3, 3, generator=gen)).to(cuda)
    R = random_rotations(n, gen).to(cuda)
    big = _ops.laplace_nll(A, R, grids, grad=True, mode=True)
    small = _ops.laplace_nll(A[:300], R[:300], grids, grad=True, mode=True)
    assert_close(big["nll"][:300].cpu().numpy(), small["nll"].cpu().numpy(), 1e-5, 1e-5, "decompositions")
    # forward-only launch (no gradient requested: the gradient sums are not accumulated): the same NLL up to the
    # last bit of the normaliser sum (without the gradient's second use of each weight, ptxas fuses the
    # weight's multiply into the accumulate)
    fwd = _ops.laplace_nll(A, R, grids, grad=False, mode=True)
    assert_close(fwd["nll"].cpu().numpy(), big["nll"].cpu().numpy(), 5e-7, 2e-6, "forward-only")
    assert torch.equal(fwd["mode"], big["mode"])
    assert grad_rel_err(big["grad"][:300].cpu().numpy(), small["grad"].cpu().numpy()).max() < 2e-4
    assert (big["mode"][:300] - small["mode"]).abs().max().item() < 1e-6
    idx = torch.arange(n - 64, n)
    ref, _ = orc.laplace_nll("RLaplace", A[idx].cpu(), R[idx].cpu(), torch.from_numpy(g["grids"]))
    assert_close(big["nll"][idx].cpu().numpy(), ref.numpy(), LAP_RTOL, LAP_ATOL, "tail rows")
    # multi-chunk grid: 3x the points = 3 shared-memory chunks; logF shifts by exactly log(1) (duplicates)
    grid3 = torch.cat([grids, grids, grids])
    tri = _ops.laplace_nll(A[:300], R[:300], grid3)
    assert_close(tri["nll"].cpu().numpy(), small["nll"].cpu().numpy(), 1e-5, 1e-5, "chunked grid")


@pytest.mark.parametrize("batch", ["one_round", "1500", "8k", "33k", "90k"])
def test_laplace_sample_packed_path(cuda, golden, batch):
    """Large batches switch to the kernel that keeps two samples per thread (launch_laplace picks it when its rounds of
    1024 samples per SM come out cheaper than the other decompositions'; odd n: the last thread's second sample is a
    dummy).  Mid-sized batches run it as thread-block CLUSTERS: the CTAs of a cluster share a tile of samples, take a
    slice of the grid each and merge their totals through distributed shared memory (8k: 8 slices, 33k: 4, 90k: 2 on a
    148-SM part); 1,501 rotations take the warp-per-sample kernel (grid staged in shared memory), and the 200-300 row
    calls they are all compared with take the CTA-per-sample kernel that training-sized batches use (the one the golden
    vectors pin).  Every form must agree with that one and the oracle, on grids whose size is
    not a multiple of 4 (trailing points), smaller than one trip (no clusters then), and larger than two shared-memory
    chunks, and on the clamp edge (every grid point within eps of the mode: A = 0 and A ~ 1e-10)."""
    from semiuhpe_b200 import _ops
    g = golden("laplace")
    grids = torch.from_numpy(g["grids"]).to(cuda)
    sms = torch.cuda.get_device_properties(cuda).multi_processor_count
    n = {"one_round": sms * 1024 - 77, "1500": 1501, "8k": 8191, "33k": 33001, "90k": 90001}[batch]
    gen = torch.Generator().manual_seed(12)
    A = (5 * torch.randn(n, 3, 3, generator=gen))
    A[:8] = 0.0
    A[8:16] *= 2e-10
    A[16:32] *= 10.0                        # sharp: the offset moves several times over the grid
    A[32:48] *= 300.0                       # q up to several hundred: most terms underflow against the offset
    A = A.to(cuda)
    R = random_rotations(n, gen).to(cuda)
    big = _ops.laplace_nll(A, R, grids, grad=True, mode=True)
    small = _ops.laplace_nll(A[:300], R[:300], grids, grad=True, mode=True)
    assert torch.isfinite(big["nll"]).all() and torch.isfinite(big["grad"]).all()
    assert_close(big["nll"][:300].cpu().numpy(), small["nll"].cpu().numpy(), 1e-5, 1e-5, "decompositions")
    assert grad_rel_err(big["grad"][16:300].cpu().numpy(), small["grad"][16:].cpu().numpy()).max() < 2e-4
    assert (big["grad"][:16] - small["grad"][:16]).abs().max().item() < 1e-3 * small["grad"][:16].abs().max().item() + 1e-6
    assert (big["mode"][16:300] - small["mode"][16:]).abs().max().item() < 1e-6
    fwd = _ops.laplace_nll(A, R, grids, grad=False, mode=True)
    assert_close(fwd["nll"].cpu().numpy(), big["nll"].cpu().numpy(), 5e-7, 2e-6, "forward-only")
    idx = torch.cat([torch.arange(0, 32), torch.arange(n - 33, n)])
    ref, _ = orc.laplace_nll("RLaplace", A[idx].cpu(), R[idx].cpu(), torch.from_numpy(g["grids"]))
    assert_close(big["nll"][idx].cpu().numpy(), ref.numpy(), LAP_RTOL, LAP_ATOL, "edge and tail rows")
    for N in (3, 5, 130, 4607, 4608 * 3 + 2):
        sub = torch.cat([grids] * 4)[:N].contiguous()
        a = _ops.laplace_nll(A, R, sub, grad=True)
        b = _ops.laplace_nll(A[:200], R[:200], sub, grad=True)
        assert_close(a["nll"][:200].cpu().numpy(), b["nll"].cpu().numpy(), 1e-5, 1e-5, f"grid of {N} points")
        assert grad_rel_err(a["grad"][16:200].cpu().numpy(), b["grad"][16:].cpu().numpy()).max() < 2e-4
        last, _ = orc.laplace_nll("RLaplace", A[-7:].cpu(), R[-7:].cpu(), sub.cpu())    # 7 rows: no grid has 7 points
        assert_close(a["nll"][-7:].cpu().numpy(), last.numpy(), LAP_RTOL, LAP_ATOL, f"tail rows, grid of {N} points")


def test_laplace_launches_in_a_cuda_graph(cuda, golden):
    """Every K2L form is a plain stream-ordered launch (the cluster forms through cudaLaunchKernelEx): captured in a
    CUDA graph after one eager call, replayed on new inputs, bit-identical to the eager result."""
    import semiuhpe_b200
    from semiuhpe_b200.laplace.rotation_laplace import NLL_loss
    g = golden("laplace")
    grids = torch.from_numpy(g["grids"]).to(cuda)
    gen = torch.Generator().manual_seed(77)
    semiuhpe_b200.set_error_checking(False)
    try:
        for n in (160, 1501, 8191, 40001):           # block / warp / clusters of 8 / clusters of 2-4
            A = (5 * torch.randn(n, 3, 3, generator=gen)).to(cuda)
            R = random_rotations(n, gen).to(cuda)
            sA, sR = torch.empty_like(A), torch.empty_like(R)
            leaf = sA.requires_grad_(True)
            out = {}

            def step():
                leaf.grad = None
                losses, mode = NLL_loss("RLaplace", leaf, sR, grids)
                losses.sum().backward()
                out["nll"], out["mode"] = losses.detach(), mode

            with torch.no_grad():
                sA.copy_(A + 1.0); sR.copy_(R)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            with torch.no_grad():
                sA.copy_(A); sR.copy_(R)
            graph.replay()
            torch.cuda.synchronize()
            got = (out["nll"].clone(), out["mode"].clone(), leaf.grad.clone())
            l2 = A.clone().requires_grad_(True)
            ref, mode = NLL_loss("RLaplace", l2, R, grids)
            ref.sum().backward()
            assert torch.equal(got[0], ref.detach()) and torch.equal(got[1], mode) and torch.equal(got[2], l2.grad), n
    finally:
        semiuhpe_b200.set_error_checking(True)


def test_metrics_golden(cuda, golden):
    from semiuhpe_b200.agent import compute_err_deg_from_matrices, eval_rotation_metrics
    from semiuhpe_b200.utils import compute_euler_angles_from_rotation_matrices as euler
    g = golden("metrics")
    Rp, Rg, Rf, ge = (torch.from_numpy(g[k]).to(cuda) for k in ("R_pd", "R_gt", "R_full", "gt_euler"))
    # atan2f/acosf on the device are ~2 ulp; 1e-6 rad absolute on angles in [-pi, pi]
    np.testing.assert_allclose(euler(Rp).cpu().numpy(), g["euler_pd"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(euler(Rf, full_range=False).cpu().numpy(), g["euler_full_false"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(euler(Rf, full_range=True).cpu().numpy(), g["euler_full_true"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(compute_err_deg_from_matrices(Rp, Rg, ge).cpu().numpy(), g["mae"], rtol=1e-5, atol=2e-4)
    # geodesic: acos amplifies fp32 rounding of the trace near 0 deg: d(angle)/d(trace) ~ 1/(2 sin)
    np.testing.assert_allclose(compute_err_deg_from_matrices(Rp, Rg).cpu().numpy(), g["geodesic_deg"], rtol=1e-4, atol=2e-3)
    np.testing.assert_allclose(compute_err_deg_from_matrices(Rf, Rg).cpu().numpy(), g["geodesic_deg_full"], rtol=1e-5, atol=2e-4)
    ev = eval_rotation_metrics(Rp, Rg, ge)
    abs_ref = np.abs(g["euler_pd"] * 180 / np.pi - g["gt_euler"])
    np.testing.assert_allclose(ev["abs_err"].cpu().numpy(), abs_ref, rtol=1e-5, atol=2e-4)
    p, y, r, m = orc.euler_mae_summary(abs_ref)
    np.testing.assert_allclose([ev["pitch"].item(), ev["yaw"].item(), ev["roll"].item(), ev["mean"].item()], [p, y, r, m], rtol=1e-5)
    ev2 = eval_rotation_metrics(Rf, Rg)
    np.testing.assert_allclose(ev2["frobenius"].cpu().numpy(), g["frob_full"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ev2["frobenius_mean"].item(), g["frob_full"].mean(), rtol=1e-6)
    np.testing.assert_allclose(ev2["geodesic_mean"].item(), g["geodesic_deg_full"].astype(np.float64).mean(), rtol=1e-5)
    with pytest.raises(ValueError):
        compute_err_deg_from_matrices(3 * torch.eye(3, device=cuda)[None], torch.eye(3, device=cuda)[None])


@pytest.mark.parametrize("n", [1, 255, 256, 257, 10007])
def test_metrics_ragged(cuda, n):
    from semiuhpe_b200 import _ops
    gen = torch.Generator().manual_seed(n)
    Rp, Rg = random_rotations(n, gen), random_rotations(n, gen)
    ge = (torch.rand(n, 3, generator=gen) * 2 - 1) * 90
    out = _ops.so3_metrics(Rp.to(cuda), Rg.to(cuda), ge.to(cuda), geo=True, frob=True, euler=True, abs_err=True, mae=True, sums=True)
    np.testing.assert_allclose(out["euler"].cpu().numpy(), orc.euler_from_matrices(Rp).numpy(), atol=2e-6)
    np.testing.assert_allclose(out["mae"].cpu().numpy(), orc.err_deg_from_matrices(Rp, Rg, ge).numpy(), rtol=1e-5, atol=2e-4)
    np.testing.assert_allclose(out["geo"].cpu().numpy(), orc.geodesic_deg(Rp, Rg).numpy(), rtol=1e-5, atol=2e-3)
    np.testing.assert_allclose(out["frob"].cpu().numpy(), orc.frobenius_identity_distance(Rp, Rg), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["sums"][0].item(), out["geo"].double().sum().item(), rtol=1e-12)
    off = torch.cat([torch.zeros(1, 9), Rp.reshape(-1, 9)]).to(cuda)[1:]
    assert torch.equal(_ops.so3_metrics(off, Rg.to(cuda), geo=True)["geo"], out["geo"])


def test_metrics_full_size_properties(cuda):
    """10 M pairs (BASELINE config 4): Euler angles of Rz Ry Rx(gt) recover gt; the geodesic
    angle of a pair built with a known relative rotation recovers that angle."""
    from semiuhpe_b200 import _ops
    n = 10_000_000
    gen = torch.Generator(device=cuda).manual_seed(4)
    e = (torch.rand(n, 3, device=cuda, generator=gen) * 2 - 1) * 89.0
    r = torch.deg2rad(e)
    cx, sx, cy, sy, cz, sz = r[:, 0].cos(), r[:, 0].sin(), r[:, 1].cos(), r[:, 1].sin(), r[:, 2].cos(), r[:, 2].sin()
    R = torch.stack([cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx,
                     sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx,
                     -sy, cy * sx, cy * cx], 1)
    out = _ops.so3_metrics(R, R, e, geo=True, mae=True, euler=True, sums=True)
    assert out["mae"].max().item() < 2e-3                        # degrees
    assert (out["geo"] - 0.40507).abs().max().item() < 5e-3      # pytorch3d's acos extension at identity
    theta = torch.rand(n, device=cuda, generator=gen) * 170 + 5
    t = torch.deg2rad(theta)
    Rz = torch.zeros(n, 9, device=cuda)
    Rz[:, 0], Rz[:, 1], Rz[:, 3], Rz[:, 4], Rz[:, 8] = t.cos(), -t.sin(), t.sin(), t.cos(), 1.0
    R2 = (R.view(n, 3, 3) @ Rz.view(n, 3, 3)).reshape(n, 9)
    geo = _ops.so3_metrics(R2, R, geo=True, frob=True)
    assert (geo["geo"] - theta).abs().max().item() < 5e-3
    assert (geo["frob"] - 2 * (2 ** 0.5) * (t / 2).sin()).abs().max().item() < 1e-5


def test_dad_euler_convention(cuda, golden):
    """K4 mode 2 (eval.py:66-74): Euler angles of a DAD-trained model in degrees, the per-angle errors
    and their means against the scipy-generated fixture."""
    from semiuhpe_b200.agent import eval_rotation_metrics
    from semiuhpe_b200.utils import euler_dad_degrees
    g = golden("dad_euler")
    R = torch.from_numpy(g["R"]).to(cuda)
    ours = euler_dad_degrees(R).cpu().numpy()
    d = np.abs(ours.astype(np.float64) - g["euler_deg"])
    d = np.minimum(d, np.abs(d - 360.0))
    assert d.max() < 1e-3
    assert d[g["names"] == "near_frontal"].max() < 1e-4
    keep = g["names"] == "near_frontal"                       # away from the +-180 wrap: errors are plain differences
    gen = torch.Generator().manual_seed(1)
    gt = torch.from_numpy(g["euler_deg"][keep]).float() + 3 * torch.randn(int(keep.sum()), 3, generator=gen)
    ev = eval_rotation_metrics(R[torch.from_numpy(keep).to(cuda)], None, gt.to(cuda), dad_trained=True)
    ref_err = np.abs(gt.numpy().astype(np.float64) - g["euler_deg"][keep])
    np.testing.assert_allclose(ev["abs_err"].cpu().numpy(), ref_err, rtol=0, atol=2e-4)
    np.testing.assert_allclose([ev["pitch"].item(), ev["yaw"].item(), ev["roll"].item()], ref_err.mean(0), rtol=1e-5)
