"""Parity of K1/K2 (through the reference-shaped Python API -> C ABI -> CUDA) with the
oracle and the golden vectors.  Tolerances: see tests/helpers.py."""
import numpy as np
import pytest
import torch

from helpers import ATOL, RTOL, assert_close, grad_rel_err, no_worse_than_reference, random_rotations
from oracle import so3_oracle as orc

pytestmark = pytest.mark.gpu


def test_vmf_loss_golden(cuda, golden):
    from semiuhpe_b200.fisher.fisher_utils import vmf_loss, KL_Fisher, fisher_log_pdf, batch_torch_A_to_R
    g = golden("fisher")
    names = g["names"]
    A = torch.from_numpy(g["A"]).to(cuda)
    R = torch.from_numpy(g["R"]).to(cuda)
    leaf = A.reshape(-1, 9).clone().requires_grad_(True)
    loss, Rest = vmf_loss(leaf, R, overreg=float(g["overreg"]))
    loss.sum().backward()
    assert loss.shape == (len(names),) and Rest.shape == (len(names), 3, 3) and leaf.grad.shape == leaf.shape
    regime = np.isin(names, ["generic1", "generic10", "generic30", "realistic", "neardegenerate"])
    assert_close(loss.detach().cpu().numpy()[regime], g["nll"][regime], RTOL, ATOL, "nll")
    grad = leaf.grad.cpu().numpy()
    well = np.isin(names, ["generic1", "generic10", "generic30", "realistic"])
    assert grad_rel_err(grad[well], g["grad"][well]).max() < 1e-5
    assert grad_rel_err(grad[names == "neardegenerate"], g["grad"][names == "neardegenerate"]).max() < 1e-4
    assert np.abs(Rest.cpu().numpy()[well] - g["Rest"][well]).max() < 1e-5
    # edge rows (huge / tiny singular values): no worse than the reference against exact arithmetic
    A64, R64 = torch.from_numpy(g["A"]).double(), torch.from_numpy(g["R"]).double()
    nll64 = orc.kl_fisher(A64, R64, float(g["overreg"])).numpy()
    edge = ~regime
    assert no_worse_than_reference(loss.detach().cpu().numpy()[edge], g["nll"][edge], nll64[edge], 2.0, 2e-6).all()
    # the other entry points share the kernel
    assert_close(KL_Fisher(A, R, overreg=float(g["overreg"])).cpu().numpy()[regime], g["nll"][regime], RTOL, ATOL, "KL_Fisher")
    assert_close(fisher_log_pdf(A, R).cpu().numpy()[regime], g["logpdf"][regime], RTOL, ATOL, "fisher_log_pdf")
    assert np.abs(batch_torch_A_to_R(A).cpu().numpy()[well] - g["Rest"][well]).max() < 1e-5
    zero = list(names).index("zero")
    assert loss[zero].item() == 0.0
    assert torch.equal(Rest[zero].cpu(), torch.eye(3))
    np.testing.assert_allclose(grad[zero], -np.eye(3).ravel(), atol=1e-6)


def test_fisher_entropy_golden(cuda, golden):
    from semiuhpe_b200.fisher.fisher_utils import fisher_entropy
    g = golden("fisher")
    names = g["names"]
    ent = fisher_entropy(torch.from_numpy(g["A"]).to(cuda).reshape(-1, 9)).cpu().numpy()
    regime = np.isin(names, ["generic1", "generic10", "generic30", "realistic", "neardegenerate"])
    assert_close(ent[regime], g["entropy"][regime], RTOL, ATOL, "entropy")
    ent64 = orc.fisher_entropy_closed_form(torch.from_numpy(g["A"]).double()).numpy()
    edge_ok = no_worse_than_reference(ent[~regime], g["entropy"][~regime], ent64[~regime], 2.0, 2e-6)
    edge_ok |= np.abs(ent[~regime] - g["entropy"][~regime]) <= 3e-5 * np.abs(g["entropy"][~regime])   # s ~ 300: quadrature breakdown regime
    assert edge_ok.all()
    row = lambda n: list(names).index(n)
    assert ent[row("zero")] == 0.0
    np.testing.assert_allclose(ent[row("diag10_5_0")], -3.309231, atol=3e-5)
    np.testing.assert_allclose(ent[row("diag300_200_100")], 7.026253, rtol=2e-5)      # quadrature breakdown reproduced
    for k, want in ((1, -0.6814), (5, -3.4806), (10, -4.5669), (30, -6.2500), (100, -8.1494), (500, -11.0276)):
        np.testing.assert_allclose(ent[row(f"eye{k}")], want, atol=3e-4)


def test_logC_F_forward_backward(cuda, golden):
    from semiuhpe_b200.fisher.torch_norm_factor import logC_F
    g = golden("fisher")
    regime = np.isin(g["names"], ["generic1", "generic10", "generic30", "realistic"])
    S = torch.from_numpy(g["S"][regime]).to(cuda).requires_grad_(True)
    out = logC_F(S)
    out.sum().backward()
    assert_close(out.detach().cpu().numpy(), g["logC"][regime], RTOL, ATOL, "logC")
    np.testing.assert_allclose(S.grad.cpu().numpy(), g["dlogC"][regime], rtol=2e-5, atol=2e-6)
    assert logC_F(torch.zeros(2, 5, 3, device=cuda)).shape == (2, 5)
    assert logC_F(torch.zeros(1, 3, device=cuda)).item() == 0.0


@pytest.mark.parametrize("b", [1, 5, 31, 32, 33, 160, 1000, 4099])
def test_batch_sizes_and_mean_backward(cuda, b):
    """Ragged sizes exercise the small-batch (few samples per warp) and tail-tile paths;
    the loss is reduced with .mean() like src/agent.py:83."""
    from semiuhpe_b200.fisher.fisher_utils import vmf_loss, fisher_entropy
    gen = torch.Generator().manual_seed(b)
    A = 10 * torch.randn(b, 9, generator=gen)
    R = random_rotations(b, gen)
    leaf = A.to(cuda).requires_grad_(True)
    loss, Rest = vmf_loss(leaf, R.to(cuda), overreg=1.025)
    loss.mean().backward()
    ref_leaf = A.clone().requires_grad_(True)
    ref_loss, ref_R = orc.vmf_loss(ref_leaf, R, overreg=1.025)
    ref_loss.mean().backward()
    assert_close(loss.detach().cpu().numpy(), ref_loss.detach().numpy(), RTOL, ATOL, "nll")
    assert grad_rel_err(leaf.grad.cpu().numpy(), ref_leaf.grad.numpy()).max() < 1e-4
    assert_close(fisher_entropy(A.to(cuda)).cpu().numpy(), orc.fisher_entropy(A).numpy(), RTOL, ATOL, "entropy")
    cond = torch.linalg.svdvals(A.view(-1, 3, 3))
    ok = ((cond[:, 1] - cond[:, 2]) / cond[:, 0] > 1e-2).numpy() | (torch.det(A.view(-1, 3, 3)) > 0).numpy()
    assert np.abs(Rest.cpu().numpy() - ref_R.detach().numpy())[ok].max() < 2e-5


def test_unaligned_and_noncontiguous_inputs(cuda):
    from semiuhpe_b200.fisher.fisher_utils import fisher_entropy, batch_torch_A_to_R
    gen = torch.Generator().manual_seed(5)
    big = (10 * torch.randn(5000, 9, generator=gen)).to(cuda)
    want = fisher_entropy(big)
    got = fisher_entropy(big[1:])            # base pointer only 4-byte aligned -> scalar tile path
    assert torch.equal(got, want[1:])
    wide = torch.zeros(5000, 12, device=cuda)
    wide[:, :9] = big
    assert torch.equal(fisher_entropy(wide[:, :9]), want)      # non-contiguous rows are packed first
    assert torch.equal(batch_torch_A_to_R(big[3:]), batch_torch_A_to_R(big)[3:])
    assert fisher_entropy(big[:0]).shape == (0,) and batch_torch_A_to_R(big[:0]).shape == (0, 3, 3)


def test_sign_convention_and_rotation_properties(cuda):
    """Proper rotation on every input: det(R)=+1, R^T R=I; sign(S3) == sign(det A) exactly
    as the reference's det(U V^T) (bit-exact convention), over 2M random matrices."""
    from semiuhpe_b200 import _ops
    gen = torch.Generator(device=cuda).manual_seed(1)
    A = torch.randn(1 << 21, 3, 3, device=cuda, generator=gen) * 5
    out = _ops.proper_svd(A, rot=True, S=True, U=True, V=True)
    R, S = out["rot"], out["S"]
    assert (torch.det(R.double()) - 1).abs().max() < 1e-5
    assert (R.transpose(1, 2) @ R - torch.eye(3, device=cuda)).abs().max() < 1e-5
    detA = torch.det(A.double())
    clear = detA.abs() > 1e-3
    assert torch.equal(torch.sign(S[clear, 2]).double(), torch.sign(detA[clear]))
    rec = out["U"] @ torch.diag_embed(S) @ out["V"].transpose(1, 2)
    assert ((rec - A).abs().amax((1, 2)) / A.abs().amax((1, 2))).max() < 3e-6
    assert (S[:, 0] >= S[:, 1] * (1 - 1e-6)).all() and (S[:, 1] >= S[:, 2].abs() * (1 - 1e-6)).all()
    sv = torch.linalg.svdvals(A[:200000].double())
    assert (S[:200000].abs().double() - sv).abs().max() < 2e-5


def test_nonfinite_input_raises_like_torch_svd(cuda):
    import semiuhpe_b200
    from semiuhpe_b200.fisher.fisher_utils import vmf_loss, fisher_entropy
    A = torch.randn(64, 9, device=cuda)
    A[17, 4] = float("nan")
    R = torch.eye(3, device=cuda).repeat(64, 1, 1)
    with pytest.raises(torch.linalg.LinAlgError):
        vmf_loss(A, R)
    with pytest.raises(torch.linalg.LinAlgError):
        fisher_entropy(A)
    semiuhpe_b200.set_error_checking(False)
    try:
        ent = fisher_entropy(A)
        assert torch.isnan(ent[17]) and torch.isfinite(ent[:17]).all()
    finally:
        semiuhpe_b200.set_error_checking(True)
        from semiuhpe_b200 import _ops
        _ops._status_word(A.device).zero_()
    with pytest.raises(RuntimeError):
        vmf_loss(A[:10], R[:9])


def test_full_size_properties(cuda):
    """BASELINE-size batch (2^23 per GPU): size-independent properties of the path.
    (1) entropy and singular values are invariant under A -> Q1 A Q2 (Q in SO(3));
    (2) nll(A,R1) - nll(A,R2) = -<A, R1-R2>;  (3) grad + R_gt does not depend on R_gt;
    (4) a random sub-sample agrees with the oracle."""
    from semiuhpe_b200 import _ops
    n = 1 << 23
    gen = torch.Generator(device=cuda).manual_seed(7)
    A = 10 * torch.randn(n, 3, 3, device=cuda, generator=gen)
    qa = torch.randn(n, 4, device=cuda, generator=gen); qb = torch.randn(n, 4, device=cuda, generator=gen)
    from semiuhpe_b200.agent import _quat_to_matrix
    R1, R2 = _quat_to_matrix(qa / qa.norm(dim=1, keepdim=True)), _quat_to_matrix(qb / qb.norm(dim=1, keepdim=True))
    o1 = _ops.fisher_fused(A, R1, 1.025, nll=True, grad=True, entropy=True, S=True)
    o2 = _ops.fisher_fused(A, R2, 1.025, nll=True, grad=True)
    lhs = (o1["nll"] - o2["nll"]).double()
    rhs = -(A.reshape(n, 9).double() * (R1 - R2).reshape(n, 9).double()).sum(1)
    assert (lhs - rhs).abs().max() < 2e-4                       # fp32 rounding of values ~1e2
    assert ((o1["grad"] + R1.reshape(n, 9)) - (o2["grad"] + R2.reshape(n, 9))).abs().max() < 1e-6
    m = 1 << 20
    rot = _ops.fisher_fused(R1[:m] @ A[:m] @ R2[:m], None, 1.0, entropy=True, S=True)
    assert (rot["S"] - o1["S"][:m]).abs().max() < 2e-4
    assert (rot["entropy"] - o1["entropy"][:m]).abs().max() < 2e-4
    idx = torch.randint(0, n, (2048,), generator=torch.Generator().manual_seed(3))
    sub_A, sub_R = A[idx.to(cuda)].cpu(), R1[idx.to(cuda)].cpu().contiguous()
    leaf = sub_A.clone().requires_grad_(True)
    ref = orc.kl_fisher(leaf, sub_R, 1.025)
    ref.sum().backward()
    assert_close(o1["nll"][idx.to(cuda)].cpu().numpy(), ref.detach().numpy(), RTOL, ATOL, "nll sample")
    assert grad_rel_err(o1["grad"][idx.to(cuda)].cpu().numpy(), leaf.grad.numpy()).max() < 1e-4
    assert_close(o1["entropy"][idx.to(cuda)].cpu().numpy(), orc.fisher_entropy(sub_A).numpy(), RTOL, ATOL, "entropy sample")


def test_negligible_node_cut_on_off(cuda, golden):
    """K2 skips the provably negligible node prefix by default (cut_bits = 26, a per-call argument of the C ABI);
    evaluating all 512 nodes (bits=0) must give the same numbers to fp32 rounding, and both
    settings must hold the golden parity."""
    import semiuhpe_b200
    from semiuhpe_b200 import _ops
    g = golden("fisher")
    A = torch.from_numpy(g["A"]).to(cuda).reshape(-1, 9)
    R = torch.from_numpy(g["R"]).to(cuda)
    regime = np.isin(g["names"], ["generic1", "generic10", "generic30", "realistic", "neardegenerate"])
    gen = torch.Generator().manual_seed(11)
    scales = torch.tensor([0.1, 1.0, 5.0, 10.0, 30.0, 100.0])[torch.randint(0, 6, (20000, 1), generator=gen)]
    Ar = (torch.randn(20000, 9, generator=gen) * scales).to(cuda)
    Rr = random_rotations(20000, gen).to(cuda)
    outs = {}
    prev = semiuhpe_b200.set_quadrature_cut_bits(0)
    try:
        for bits in (0, 26):
            semiuhpe_b200.set_quadrature_cut_bits(bits)
            o = _ops.fisher_fused(A, R, float(g["overreg"]), nll=True, entropy=True, grad=True)
            assert_close(o["nll"].cpu().numpy()[regime], g["nll"][regime], RTOL, ATOL, f"nll bits={bits}")
            assert_close(o["entropy"].cpu().numpy()[regime], g["entropy"][regime], RTOL, ATOL, f"entropy bits={bits}")
            outs[bits] = _ops.fisher_fused(Ar, Rr, 1.025, nll=True, entropy=True, grad=True, logC=True, G=True, S=True)
    finally:
        semiuhpe_b200.set_quadrature_cut_bits(prev)
    full, cut = outs[0], outs[26]
    assert_close(cut["logC"].cpu().numpy(), full["logC"].cpu().numpy(), 3e-7, 3e-7, "logC")
    assert (cut["G"] - full["G"]).abs().max().item() < 6e-7
    S1 = full["S"].abs().sum(1)
    assert bool(((cut["entropy"] - full["entropy"]).abs() <= 2e-6 + 6e-7 * S1).all())
    assert grad_rel_err(cut["grad"].cpu().numpy(), full["grad"].cpu().numpy()).max() < 2e-6


def test_scale_sweep_against_oracle(cuda):
    """6,000 random parameter matrices over the scales the heads are trained at (singular values from
    ~0.3 to ~60): every combination of run types, odd / even type boundaries (edge nodes), cut and uncut
    families -- NLL, entropy, gradient and rotation against the oracle's fp32 restatement of the reference."""
    from semiuhpe_b200.fisher.fisher_utils import vmf_loss, fisher_entropy
    gen = torch.Generator().manual_seed(2024)
    scales = torch.tensor([0.3, 1.0, 3.0, 5.0, 10.0, 20.0]).repeat_interleave(1000)
    A = torch.randn(len(scales), 9, generator=gen) * scales[:, None]
    R = random_rotations(len(scales), gen)
    leaf = A.to(cuda).requires_grad_(True)
    loss, Rest = vmf_loss(leaf, R.to(cuda), overreg=1.025)
    loss.sum().backward()
    ent = fisher_entropy(A.to(cuda))
    ref_leaf = A.clone().requires_grad_(True)
    ref_loss, ref_R = orc.vmf_loss(ref_leaf, R, overreg=1.025)
    ref_loss.sum().backward()
    ref_ent = orc.fisher_entropy(A)
    assert_close(loss.detach().cpu().numpy(), ref_loss.detach().numpy(), RTOL, ATOL, "nll")
    assert_close(ent.cpu().numpy(), ref_ent.numpy(), RTOL, 2 * ATOL, "entropy")
    # forward-only launches (no gradient requested: one family of the quadrature) give the same NLL bit for bit
    from semiuhpe_b200.fisher.fisher_utils import KL_Fisher
    with torch.no_grad():
        assert torch.equal(KL_Fisher(A.to(cuda), R.to(cuda), overreg=1.025), loss.detach())
    # gradients: 1e-5 away from singular-value degeneracies, 1e-4 near them (BASELINE north_star)
    S = torch.linalg.svdvals(A.reshape(-1, 3, 3).double())
    gap = torch.minimum(S[:, 0] - S[:, 1], S[:, 1] - S[:, 2]) / S[:, 0]
    err = grad_rel_err(leaf.grad.cpu().numpy(), ref_leaf.grad.numpy())
    assert err[(gap > 0.05).numpy()].max() < 1e-5 and err.max() < 1e-4
    # rotation: conditioning ~ eps * s1 / (s2 + s3)
    sgn = torch.sign(torch.linalg.det(A.reshape(-1, 3, 3).double()))
    cond = (S[:, 0] / (S[:, 1] + sgn * S[:, 2]).clamp_min(1e-9)).numpy()
    dR = np.abs(Rest.cpu().numpy() - ref_R.detach().numpy()).reshape(len(scales), -1).max(1)
    assert (dR <= 2e-6 * np.maximum(cond, 1.0) + 2e-6).all()
