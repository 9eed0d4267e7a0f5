"""Restatement of the three ``pytorch3d.transforms`` functions on the hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference pins ``pytorch3d==0.7.2`` in its README (README.md:48) but does
not vendor it, and the wheel is not in this image, so these are restated from
the published algorithm of that release.  **PARITY UNPINNED**: no reference
test or fixture covers them; the anchors are the reference's call sites:

* ``so3_relative_angle``    src/agent.py:450, src/agent.py:422, eval.py:88
* ``matrix_to_quaternion``  src/fisher/between_bingham_fisher.py:15
* ``quaternion_to_matrix``  src/fisher/between_bingham_fisher.py:10

Published behaviour restated here
---------------------------------
``so3_relative_angle(R1, R2)``: ``c = (trace(R1 R2^T) - 1) / 2``; a trace
outside ``[-1-eps, 3+eps]`` (eps=1e-4) raises ``ValueError``; the angle is
``acos(c)`` for ``|c| < 1-1e-4`` and the first-order Taylor extension of acos
about ``+-(1-1e-4)`` outside, ``acos(b) + (x - b) * (-1/sqrt(1 - b^2))`` -- so identical rotations give
``acos(1-1e-4) - 1e-4/sqrt(1-(1-1e-4)^2)`` = 0.0141422 - 0.0070711 = **0.0070711 rad = 0.40514 deg**, not 0.

Known answers derived from these formulas in plain float64 (``tests/golden/make_pytorch3d_kat.py`` ->
``tests/golden/pytorch3d_kat.json``, independent of this file) are checked in ``tests/test_oracle_golden.py``;
they pin the restatement to the documented algorithm, not to pytorch3d's own output -- the row stays
"parity unpinned" until a real pytorch3d result has been compared.
"""
import math

import torch

COS_BOUND = 1e-4
TRACE_EPS = 1e-4


def acos_linear_extrapolation(x, bound=1.0 - COS_BOUND):
    """acos on (-bound, bound), tangent-line continuation elsewhere."""
    slope_hi = -1.0 / math.sqrt(1.0 - bound * bound)
    slope_lo = -1.0 / math.sqrt(1.0 - bound * bound)
    hi = x >= bound
    lo = x <= -bound
    inner = torch.acos(torch.where(hi | lo, torch.zeros_like(x), x))
    ext_hi = (x - bound) * slope_hi + math.acos(bound)
    ext_lo = (x + bound) * slope_lo + math.acos(-bound)
    return torch.where(hi, ext_hi, torch.where(lo, ext_lo, inner))


def so3_rotation_angle(R, eps=TRACE_EPS, cos_angle=False, cos_bound=COS_BOUND):
    if R.dim() != 3 or R.shape[1:] != (3, 3):
        raise ValueError("Input has to be a batch of 3x3 Tensors.")
    tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    if ((tr < -1.0 - eps) | (tr > 3.0 + eps)).any():
        raise ValueError("A matrix has trace outside valid range [-1-eps,3+eps].")
    c = (tr - 1.0) * 0.5
    if cos_angle:
        return c
    if cos_bound > 0.0:
        return acos_linear_extrapolation(c, 1.0 - cos_bound)
    return torch.acos(c)


def so3_relative_angle(R1, R2, cos_angle=False, cos_bound=COS_BOUND, eps=TRACE_EPS):
    R12 = torch.bmm(R1, R2.permute(0, 2, 1))
    return so3_rotation_angle(R12, eps=eps, cos_angle=cos_angle, cos_bound=cos_bound)


def matrix_to_quaternion(matrix):
    """Real-first quaternion; picks the best-conditioned of four candidates."""
    if matrix.shape[-2:] != (3, 3):
        raise ValueError(f"Invalid rotation matrix shape {matrix.shape}.")
    lead = matrix.shape[:-2]
    m = matrix.reshape(lead + (9,))
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m, -1)
    radicand = torch.stack(
        (1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22,
         1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22), dim=-1)
    q_abs = torch.sqrt(torch.clamp_min(radicand, 0.0))
    cand = torch.stack((
        torch.stack((q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01), -1),
        torch.stack((m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20), -1),
        torch.stack((m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21), -1),
        torch.stack((m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2), -1),
    ), dim=-2)
    floor = torch.tensor(0.1, dtype=q_abs.dtype, device=q_abs.device)
    cand = cand / (2.0 * torch.maximum(q_abs[..., None], floor))
    pick = q_abs.argmax(dim=-1)
    idx = pick[..., None, None].expand(lead + (1, 4))
    return torch.gather(cand, -2, idx).squeeze(-2)


def quaternion_to_matrix(quaternions):
    r, i, j, k = torch.unbind(quaternions, -1)
    two_s = 2.0 / (quaternions * quaternions).sum(-1)
    o = torch.stack((
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
    ), -1)
    return o.reshape(quaternions.shape[:-1] + (3, 3))
