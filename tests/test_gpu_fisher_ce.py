"""fisher_CE (the default unsupervised loss): two K2 launches + the closing kernel
against the golden vectors generated from the live reference, the oracle, and size-independent
properties at large n."""
import numpy as np
import pytest
import torch

from helpers import ATOL, RTOL, grad_rel_err, no_worse_than_reference
from oracle import so3_oracle as orc

pytestmark = pytest.mark.gpu


def test_fisher_ce_golden(cuda, golden):
    from semiuhpe_b200.fisher.fisher_utils import fisher_CE
    g = golden("fisher_ce")
    A1 = torch.from_numpy(g["A1"]).to(cuda)
    A2 = torch.from_numpy(g["A2"]).to(cuda).requires_grad_(True)
    ce = fisher_CE(A1, A2)
    assert ce.shape == (len(g["A1"]),)
    ce.sum().backward()
    ours, grad = ce.detach().cpu().numpy(), A2.grad.cpu().numpy()
    ok = np.abs(ours - g["ce"]) <= ATOL + RTOL * np.abs(g["ce"])
    assert (ok | no_worse_than_reference(ours, g["ce"], g["ce64"])).all()
    err64, ref_err64 = grad_rel_err(grad, g["grad64"]), grad_rel_err(g["grad"], g["grad64"])
    stable = g["names"] != "student_neardegenerate"
    assert (err64[stable] <= np.maximum(2 * ref_err64[stable], 2e-5)).all(), err64[stable].max()
    # near-degenerate students: 1/(s_i - s_j) amplifies fp32 rounding sample by sample; compare the class
    assert err64[~stable].max() <= max(2 * ref_err64[~stable].max(), 1e-4), err64[~stable].max()


@pytest.mark.parametrize("n", [1, 31, 33, 1000])
def test_fisher_ce_ragged_and_layouts(cuda, n):
    """(n,9) and (n,3,3) inputs, ragged tile tails, unaligned base pointers; the training-loop form
    loss.mean().backward() (src/agent.py:155-163)."""
    from semiuhpe_b200.fisher.fisher_utils import fisher_CE
    gen = torch.Generator().manual_seed(n)
    A1 = 10 * torch.randn(n, 9, generator=gen)
    A2 = A1 + 2 * torch.randn(n, 9, generator=gen)
    leaf = A2.double().requires_grad_(True)
    ref = orc.fisher_ce(A1.double(), leaf)                 # fp64 restatement + autograd: the exact anchor
    ref.mean().backward()
    dev = A2.to(cuda).requires_grad_(True)
    ce = fisher_CE(A1.to(cuda).view(n, 3, 3), dev)
    ce.mean().backward()
    np.testing.assert_allclose(ce.detach().cpu().numpy(), ref.detach().numpy(), rtol=2e-5, atol=5e-5)
    assert dev.grad.shape == (n, 9)
    # gradients: 1e-4 where the student's singular values are separated, looser next to a degeneracy
    # (s_i = +-s_j), where 1/(s_i -+ s_j) amplifies fp32 rounding for the reference and for us alike
    s = torch.linalg.svdvals(A2.view(n, 3, 3).double())
    gap = torch.minimum(s[:, 0] - s[:, 1], s[:, 1] - s[:, 2]) / s[:, 0]
    err = grad_rel_err(dev.grad.cpu().numpy(), leaf.grad.numpy())
    separated = (gap > 2e-2).numpy()
    assert err[separated].max() < 1e-4, err[separated].max()
    assert err.max() < 2e-2, err.max()
    # unaligned base pointers take the scalar tile path: identical results
    pad1, pad2 = torch.zeros(n * 9 + 1, device=cuda), torch.zeros(n * 9 + 1, device=cuda)
    pad1[1:] = A1.to(cuda).reshape(-1); pad2[1:] = A2.to(cuda).reshape(-1)
    off = fisher_CE(pad1[1:].view(n, 9), pad2[1:].view(n, 9))
    assert torch.equal(off, ce.detach())


def test_fisher_ce_with_supplied_target_statistics(cuda):
    """target_G from the entropy launch on the UNADJUSTED teacher output equals what fisher_CE computes
    itself for the rotate-adjusted target (singular values are invariant): same result, one K2 launch less."""
    from semiuhpe_b200 import _ops
    from semiuhpe_b200.agent import rotate_aug_adjust
    from semiuhpe_b200.fisher.fisher_utils import fisher_CE
    from helpers import random_rotations
    gen = torch.Generator().manual_seed(2)
    n = 5000
    weak = (10 * torch.randn(n, 9, generator=gen)).to(cuda)
    strong = (weak + 2 * torch.randn(n, 9, generator=gen).to(cuda)).requires_grad_(True)
    adj = rotate_aug_adjust(weak, random_rotations(n, gen).to(cuda), "300WLP")
    G = _ops.fisher_fused(weak, None, 1.0, entropy=True, G=True)["G"]
    a = fisher_CE(adj, strong)
    a.sum().backward()
    ga = strong.grad.clone(); strong.grad = None
    b = fisher_CE(adj, strong, target_G=G)
    b.sum().backward()
    assert torch.allclose(a, b, rtol=2e-6, atol=5e-6)
    # G from the unadjusted teacher differs from G of the adjusted one by fp32 rounding of the singular values
    assert grad_rel_err(strong.grad.cpu().numpy(), ga.cpu().numpy()).max() < 1e-4


def test_fisher_ce_errors(cuda):
    from semiuhpe_b200.fisher.fisher_utils import fisher_CE
    A = 5 * torch.randn(8, 9, device=cuda)
    with pytest.raises(NotImplementedError):
        fisher_CE(A.clone().requires_grad_(True), A)
    bad = A.clone(); bad[3, 4] = float("nan")
    with pytest.raises((torch.linalg.LinAlgError, AssertionError)):
        fisher_CE(A, bad)
    with pytest.raises(RuntimeError):
        fisher_CE(A[:4], A)
    assert fisher_CE(A[:0], A[:0]).shape == (0,)


def test_fisher_ce_full_size_properties(cuda):
    """2^20 pairs: h(f, f) equals the entropy of f (K2's own entropy output) -- with coinciding
    frames the row/column quirk of bingham_utils.py:27 is invisible -- and the analytic gradient
    agrees with central finite differences of the value along random directions (subsample).
    (Gibbs' inequality is NOT a property of the reference's formula because of that quirk.)"""
    from semiuhpe_b200.fisher.fisher_utils import fisher_CE, fisher_entropy
    n = 1 << 20
    gen = torch.Generator(device=cuda).manual_seed(3)
    A1 = 10 * torch.randn(n, 9, device=cuda, generator=gen)
    same = fisher_CE(A1, A1.clone())
    ent = fisher_entropy(A1)
    assert torch.allclose(same, ent, rtol=1e-5, atol=3e-5), (same - ent).abs().max().item()
    A2 = (A1 + 0.5 * torch.randn(n, 9, device=cuda, generator=gen)).requires_grad_(True)
    ce = fisher_CE(A1, A2)
    assert bool(torch.isfinite(ce).all())
    ce.sum().backward()
    m = 4096
    d = torch.randn(m, 9, device=cuda, generator=gen)
    h = 5e-2                                   # fp32 values: rounding noise ~5e-6 / 2h against O(h^2) truncation
    with torch.no_grad():
        fd = (fisher_CE(A1[:m], A2[:m] + h * d) - fisher_CE(A1[:m], A2[:m] - h * d)) / (2 * h)
    an = (A2.grad[:m] * d).sum(1)
    scale = A2.grad[:m].abs().max(1)[0] * d.abs().max(1)[0]
    assert float(((fd - an).abs() / scale).median()) < 1e-2
