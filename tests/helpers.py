"""Shared test helpers (input generators, tolerance rules)."""
import numpy as np
import torch

# Tolerances (BASELINE.json north_star): NLL / entropy / gradient within 1e-5 relative in
# fp32, 1e-4 for gradients near singular-value degeneracies; an absolute floor is needed
# because NLL and entropy cross zero.
RTOL = 1e-5
ATOL = 1e-5
GRAD_RTOL = 1e-5
GRAD_RTOL_DEGENERATE = 1e-4


def random_rotations(n, gen=None):
    q, _ = torch.linalg.qr(torch.randn(n, 3, 3, generator=gen))
    q[:, :, 2] *= torch.det(q)[:, None]
    return q.contiguous()


def close(ours, ref, rtol=RTOL, atol=ATOL):
    ours, ref = np.asarray(ours, np.float64), np.asarray(ref, np.float64)
    err = np.abs(ours - ref)
    bound = atol + rtol * np.abs(ref)
    bad = ~(err <= bound) & ~(np.isnan(ours) & np.isnan(ref))
    return not bad.any(), float((err / bound)[np.isfinite(err / bound)].max()) if err.size else 0.0


def assert_close(ours, ref, rtol=RTOL, atol=ATOL, what=""):
    ok, worst = close(ours, ref, rtol, atol)
    assert ok, f"{what}: worst error is {worst:.2f}x the bound (rtol={rtol}, atol={atol})"


def grad_rel_err(ours, ref):
    """per-sample max-norm relative error of a (n,9) gradient"""
    ours, ref = np.asarray(ours, np.float64).reshape(-1, 9), np.asarray(ref, np.float64).reshape(-1, 9)
    return np.abs(ours - ref).max(1) / np.maximum(np.abs(ref).max(1), 1e-12)


def no_worse_than_reference(ours, ref32, ref64, slack=2.0, floor=1e-6):
    """|ours - exact| <= slack*|ref32 - exact| + floor*(1+|exact|): the rule for quantities where the
    reference's own fp32 noise exceeds 1e-5 (Laplace, large-s entropy)."""
    ours, ref32, ref64 = (np.asarray(a, np.float64) for a in (ours, ref32, ref64))
    return np.abs(ours - ref64) <= slack * np.abs(ref32 - ref64) + floor * (1 + np.abs(ref64))
