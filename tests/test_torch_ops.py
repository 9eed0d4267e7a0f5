"""torch.library registration (semiuhpe_b200.torch_ops): schema, fake-tensor implementations on the meta device
(no GPU needed), and -- on the GPU -- opcheck plus a compiled (aot_eager) step that matches the eager mirrors."""
import pytest
import torch


def test_ops_are_registered_with_fake_impls(built):
    from semiuhpe_b200 import torch_ops
    ns = torch.ops.semiuhpe_b200
    for name in torch_ops.OPS:
        assert hasattr(ns, name), name
    A = torch.empty(7, 9, device="meta")
    R = torch.empty(7, 3, 3, device="meta")
    grids = torch.empty(4608, 3, 3, device="meta")
    nll, rot, grad = ns.fisher_nll(A, R, 1.025, True, True, None)
    assert nll.shape == (7,) and rot.shape == (7, 3, 3) and grad.shape == (7, 9)
    nll, rot, grad = ns.fisher_nll(A.view(7, 3, 3), R, 1.025, False, False, None)
    assert nll.shape == (7,) and rot.shape == (0, 3, 3) and grad.shape == (0, 9)
    assert ns.fisher_entropy(A).shape == (7,) and ns.proper_rotation(A).shape == (7, 3, 3)
    ce, g = ns.fisher_ce(A, A, True, None, None)
    assert ce.shape == (7,) and g.shape == (7, 9)
    nll, mode, g = ns.laplace_nll(A, R, grids, True, None)
    assert nll.shape == (7,) and mode.shape == (7, 3, 3) and g.shape == (7, 9)
    assert ns.geodesic_deg(R, R).shape == (7,)
    assert ns.scale_rows(A, torch.empty(7, device="meta"), None).shape == (7, 9)
    schema = str(ns.fisher_nll.default._schema)
    assert "Tensor? keep" in schema and "float overreg" in schema


@pytest.mark.gpu
def test_opcheck_and_compiled_step_match_eager(cuda):
    from helpers import random_rotations
    from semiuhpe_b200 import torch_ops  # noqa: F401
    from semiuhpe_b200.fisher.fisher_utils import vmf_loss, fisher_entropy, fisher_CE, batch_torch_A_to_R
    from semiuhpe_b200.agent import compute_err_deg_from_matrices
    gen = torch.Generator().manual_seed(2)
    A = (10 * torch.randn(40, 9, generator=gen)).to(cuda)
    R = random_rotations(40, gen).to(cuda)
    ns = torch.ops.semiuhpe_b200
    torch.library.opcheck(ns.fisher_entropy.default, (A,))
    torch.library.opcheck(ns.proper_rotation.default, (A,))
    torch.library.opcheck(ns.geodesic_deg.default, (R, R.flip(0).contiguous()))
    torch.library.opcheck(ns.scale_rows.default, (A, torch.randn(40, device=cuda), None))
    torch.library.opcheck(ns.fisher_nll.default, (A.clone().requires_grad_(True), R, 1.025, True, True, None),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))
    torch.library.opcheck(ns.fisher_ce.default, (A, (A + 1).requires_grad_(True), True, None, None),
                          test_utils=("test_schema", "test_faketensor", "test_autograd_registration"))

    def step(net_out, strong):
        loss, rest = vmf_loss(net_out, R, overreg=1.025)
        ent = fisher_entropy(net_out.detach())
        mask = ent < -4.0
        ce = fisher_CE(net_out.detach(), strong, keep=mask)
        err = compute_err_deg_from_matrices(batch_torch_A_to_R(strong.detach()), rest)
        return loss.mean() + 0.5 * ce.sum() / ce.numel(), err

    a1, s1 = A.clone().requires_grad_(True), (A + 0.3).requires_grad_(True)
    total, err = step(a1, s1)
    total.backward()
    compiled = torch.compile(step, backend="aot_eager", fullgraph=True)
    a2, s2 = A.clone().requires_grad_(True), (A + 0.3).requires_grad_(True)
    total2, err2 = compiled(a2, s2)
    total2.backward()
    assert torch.equal(total, total2) and torch.equal(err, err2)
    assert torch.equal(a1.grad, a2.grad) and torch.equal(s1.grad, s2.grad)
