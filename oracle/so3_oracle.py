"""torch-CPU restatement of SemiUHPE's rotation-distribution hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): the checker for the CUDA
kernels and the timed CPU baseline of ``bench.py``; never on a product path.

It restates, function by function, what the reference computes (citations are
``/root/reference``-relative ``file:line``) with the same primitive operations
in the same order, so that in fp32 it reproduces the reference to rounding
(pinned by ``tests/golden/*.npz``, generated from the live reference by
``tests/golden/make_golden.py``).  Every function is dtype-generic: fed fp64
it is the exact-arithmetic anchor used for the "no worse than the reference
against fp64" tolerance rule.

The reference bounces every tensor through the host (``.cpu()`` ... ``.cuda()``);
this restatement simply stays on the CPU.
"""
import math

import numpy as np
import torch

from . import pytorch3d_restated as p3d

QUAD_NODES = 512  # src/fisher/torch_norm_factor.py:69,83
LAPLACE_EPS = 1e-8  # src/laplace/rotation_laplace.py:11

# Abramowitz & Stegun 9.8.1 / 9.8.2, highest power first
# (src/fisher/torch_norm_factor.py:10-11 stores them reversed the same way).
_I0_SMALL = (0.45813e-2, 0.360768e-1, 0.2659732, 1.2067492, 3.0899424, 3.5156229, 1.0)
_I0_LARGE = (0.392377e-2, -0.1647633e-1, 0.2635537e-1, -0.2057706e-1, 0.916281e-2,
             -0.157565e-2, 0.225319e-2, 0.1328592e-1, 0.39894228)


# --------------------------------------------------------------------------- a1
def _poly(coeffs, x):
    """Horner with a separately rounded multiply and add per step
    (src/fisher/torch_norm_factor.py:4-8: ``z.mul_(x).add_(c)``)."""
    acc = torch.full_like(x, coeffs[0])
    for c in coeffs[1:]:
        acc = acc * x
        acc = acc + c
    return acc


def i0e(x):
    """Exponentially scaled Bessel ``exp(-|x|) I0(x)`` by the two A&S
    polynomials, switch at |x| = 3.75; both branches are evaluated everywhere
    and merged by mask (src/fisher/torch_norm_factor.py:12-19)."""
    ax = torch.abs(x)
    small = _poly(_I0_SMALL, (ax / 3.75) ** 2) / torch.exp(ax)
    large = _poly(_I0_LARGE, 3.75 / ax) / torch.sqrt(ax)
    return torch.where(ax <= 3.75, small, large)


# --------------------------------------------------------------------------- a2
def quad_nodes(dtype):
    """x_i = fl(i * fl(2/511)) - 1, trapezoid weights 1/2,1,...,1,1/2
    (src/fisher/torch_norm_factor.py:25-29)."""
    idx = torch.arange(QUAD_NODES, dtype=dtype)
    x = (idx * (2.0 / (QUAD_NODES - 1)) + (-1)).view(1, QUAD_NODES)
    w = torch.ones((1, QUAD_NODES), dtype=dtype)
    w[0, 0] = 0.5
    w[0, -1] = 0.5
    return x, w


def trapezoid(integrand, v):
    """(b,3) -> (b,), src/fisher/torch_norm_factor.py:21-31."""
    with torch.no_grad():
        x, w = quad_nodes(v.dtype)
        y = integrand(x, v)
        return torch.sum(y * w, dim=1) * 2 / (QUAD_NODES - 1)


# ----------------------------------------------------------------------- a3, a4
def _bessel_product(x, lo, hi, shift):
    """I0e(lo-hi half-difference * (1-x)) * I0e(half-sum * (1+x)) and the
    exponential tail, the part shared by both integrands."""
    d = ((hi - lo) / 2).view(-1, 1) * (1 - x)
    s = ((hi + lo) / 2).view(-1, 1) * (1 + x)
    tail = shift.view(-1, 1) * (x - 1)
    return i0e(d), i0e(s), tail


def integrand_norm(x, s):
    """Integrand of the normaliser, s sorted s1>=s2>=|s3|
    (src/fisher/torch_norm_factor.py:33-47)."""
    b1, b2, tail = _bessel_product(x, s[:, 2], s[:, 1], s[:, 2] + s[:, 0])
    return b1 * b2 * torch.exp(tail)


def integrand_dnorm(x, cv):
    """Integrand of d(normaliser)/d(cv[:,0]); the two other entries are used
    by value as (max, min) (src/fisher/torch_norm_factor.py:50-63)."""
    top = torch.max(cv[:, 1:], dim=1).values
    bot = torch.min(cv[:, 1:], dim=1).values
    b1, b2, tail = _bessel_product(x, bot, top, bot + cv[:, 0])
    return b1 * b2 * x * torch.exp(tail)


# --------------------------------------------------------------------------- a5
class _LogNormaliser(torch.autograd.Function):
    """log c(S) of the matrix-Fisher distribution and its gradient, by 512-node
    trapezoid quadrature (src/fisher/torch_norm_factor.py:66-90)."""

    @staticmethod
    def forward(ctx, S):
        flat = S.reshape(-1, 3)
        half = 0.5 * trapezoid(integrand_norm, flat)
        ctx.save_for_backward(S, half)
        return (torch.log(half) + flat.sum(dim=1)).view(S.shape[:-1])

    @staticmethod
    def backward(ctx, grad):
        S, half = ctx.saved_tensors
        flat = S.reshape(-1, 3)
        out = torch.empty_like(flat)
        for i in range(3):
            rolled = torch.cat((flat[:, i:], flat[:, :i]), dim=1)
            out[:, i] = 0.5 * trapezoid(integrand_dnorm, rolled)
        out = out / half.view(-1, 1)
        out = out * grad.reshape(-1, 1)
        return out.view(*grad.shape, 3)


log_normaliser = _LogNormaliser.apply  # reference name: logC_F (torch_norm_factor.py:92)


def log_normaliser_grad(S):
    """g = d logC / dS evaluated directly (no autograd), plus f = 1/2 * integral."""
    flat = S.reshape(-1, 3)
    half = 0.5 * trapezoid(integrand_norm, flat)
    g = torch.empty_like(flat)
    for i in range(3):
        rolled = torch.cat((flat[:, i:], flat[:, :i]), dim=1)
        g[:, i] = 0.5 * trapezoid(integrand_dnorm, rolled) / half
    return half, g


# ----------------------------------------------------------------------- a6..a9
def _svd_signed(A):
    """torch.svd on the CPU and the *computed* determinant of U V^T as the sign
    (src/fisher/fisher_utils.py:28-31): a float close to +-1, not rounded."""
    U, S, V = torch.svd(A)
    with torch.no_grad():
        sgn = torch.det(torch.matmul(U, V.transpose(1, 2)))
    return U, S, V, sgn


def kl_fisher(A, R, overreg=1.05):
    """Matrix-Fisher NLL, (b,3,3),(b,3,3) -> (b,)  (src/fisher/fisher_utils.py:21-36)."""
    A = A.reshape(-1, 3, 3)
    _, S, _, sgn = _svd_signed(A)
    S_signed = torch.cat((S[:, :2], S[:, 2:] * sgn[:, None]), -1)
    logc = log_normaliser(S_signed)
    inner = -torch.matmul(A.reshape(-1, 1, 9), R.view(-1, 9, 1)).view(-1)
    return inner + overreg * logc


def a_to_r(A):
    """Proper-SVD projection of A onto SO(3) (src/fisher/fisher_utils.py:39-48)."""
    A = A.reshape(-1, 3, 3)
    U, _, V, sgn = _svd_signed(A)
    U = torch.cat((U[:, :, :2], U[:, :, 2:] * sgn[:, None][:, None]), -1)
    return torch.matmul(U, V.transpose(1, 2))


def vmf_loss(net_out, R, overreg=1.05):
    """(NLL, projected rotation)  (src/fisher/fisher_utils.py:14-18)."""
    A = net_out.view(-1, 3, 3)
    return kl_fisher(A, R, overreg=overreg), a_to_r(A)


def fisher_log_pdf(A, R):
    """src/fisher/fisher_utils.py:51-67."""
    A = A.reshape(-1, 3, 3)
    _, S, _, sgn = _svd_signed(A)
    S_signed = torch.cat((S[:, :2], S[:, 2:] * sgn[:, None]), -1)
    return -log_normaliser(S_signed) + torch.matmul(A.reshape(-1, 1, 9), R.view(-1, 9, 1)).view(-1)


# -------------------------------------------------------------------------- a10
def proper_svd(F):
    """U, V in SO(3), signed S  (src/fisher/between_bingham_fisher.py:63-82)."""
    u1, s1, v1 = torch.svd(F)
    du = torch.det(u1).reshape(-1, 1, 1)
    u = torch.cat((u1[:, :, :-1], u1[:, :, -1:] * du), -1)
    duv = torch.det(u1 @ v1).reshape(-1, 1)
    s = torch.cat((s1[:, :-1], s1[:, -1:] * duv), -1)
    dv = torch.det(v1).reshape(-1, 1, 1)
    v = torch.cat((v1[:, :, :-1], v1[:, :, -1:] * dv), -1)
    return u, s, v


def fisher_to_bingham_lambda(s):
    """Fisher-convention Lambda (src/fisher/between_bingham_fisher.py:85-97) moved to
    the Bingham convention: shift so max is 0, sort descending (:138-145)."""
    s1, s2, s3 = s[:, 0], s[:, 1], s[:, 2]
    l1 = s1 - s2 - s3
    l2 = s2 - s1 - s3
    l3 = s3 - s1 - s2
    l4 = -l1 - l2 - l3
    lam = torch.stack((l1, l2, l3, l4), 1)
    lam = lam + (-lam.max(1)[0]).unsqueeze(-1)
    return lam.sort(descending=True)[0]


def _bingham_lambda_to_s(lamB):
    """src/fisher/between_bingham_fisher.py:164-191."""
    l1, l2, l3, l4 = lamB[:, 0], lamB[:, 1], lamB[:, 2], lamB[:, 3]
    raw = torch.stack((1 / 4 * (l1 - l2 - l3 + l4),
                       1 / 4 * (-l1 + l2 - l3 + l4),
                       1 / 4 * (-l1 - l2 + l3 + l4)), -1)
    mag = torch.abs(raw).sort(descending=True)[0]
    sign = torch.sign(raw[:, 0] * raw[:, 1] * raw[:, 2])
    last = mag[:, -1:] * sign[:, None]
    return torch.cat((mag[:, :-1], last), -1)


def _bingham_F(lamB):
    """F = 2 pi^2 exp(logC_F(S) + sum(Lam)/4)
    (src/fisher/bingham_utils.py:47-56, between_bingham_fisher.py:194-200)."""
    c = lamB.sum(1) / 4
    S = _bingham_lambda_to_s(lamB)
    return 2 * np.pi ** 2 * torch.exp(log_normaliser(S) + c)


def bingham_entropy(lamB):
    """log F - sum Lam_i dF/dLam_i / F, dF by autograd through the quadrature's
    custom backward (src/fisher/bingham_utils.py:35-44,59-73)."""
    first = torch.log(_bingham_F(lamB))
    with torch.enable_grad():
        leaf = lamB.detach().requires_grad_(True)
        F = _bingham_F(leaf)
        dF = torch.autograd.grad(F, leaf, torch.ones_like(F))[0]
    second = -(lamB * dF / _bingham_F(lamB)[:, None]).sum(1)
    return first + second


def fisher_entropy(A):
    """(b,9)|(b,3,3) -> (b,)  (src/fisher/fisher_utils.py:70-81).  The quaternion
    frame V the reference also builds (between_bingham_fisher.py:118-133) does not
    enter the entropy and is skipped."""
    A = A.reshape(-1, 3, 3)
    _, s, _ = proper_svd(A)
    lamB = fisher_to_bingham_lambda(s)
    ent = bingham_entropy(lamB)
    return ent - torch.tensor([np.log(2 * np.pi ** 2)], dtype=ent.dtype)


def fisher_entropy_closed_form(A):
    """H = log f(s) + sum_j s_j (1 - g_j(s)), the algebraic
    collapse of the chain above; used with fp64 input as the exact anchor."""
    A = A.reshape(-1, 3, 3)
    _, s, _ = proper_svd(A)
    half, g = log_normaliser_grad(s)
    return torch.log(half) + (s * (1 - g)).sum(1)


# ------------------------------------------------------------------- f1
def fisher_to_bingham_frame(A, dtype_eye=None):
    """Quaternion frame V (b,4,4) and Fisher-convention Lambda (b,4) of a matrix-Fisher
    parameter (src/fisher/between_bingham_fisher.py:107-135): columns are the quaternions
    of U E_k V^T (E_k = 2 e_k e_k^T - I) and of U V^T."""
    u, s, v = proper_svd(A)
    s1, s2, s3 = s[:, 0], s[:, 1], s[:, 2]
    l1 = s1 - s2 - s3
    l2 = s2 - s1 - s3
    l3 = s3 - s1 - s2
    l4 = -l1 - l2 - l3
    lam = torch.stack((l1, l2, l3, l4), 1)
    eye = torch.eye(3, dtype=A.dtype)
    cols = []
    for k in range(4):
        if k < 3:
            e = torch.zeros(3, 1, dtype=A.dtype)
            e[k] += 1
            E = 2 * (e @ e.t()) - eye
        else:
            E = eye
        cols.append(p3d.matrix_to_quaternion(u @ E[None] @ v.transpose(1, 2)))
    return torch.stack(cols, 2), lam


def to_bingham_convention(V, lam):
    """Shift so the largest Lambda is 0, sort descending, permute the frame's columns alike
    (src/fisher/between_bingham_fisher.py:138-152)."""
    lam = lam + (-lam.max(1)[0]).unsqueeze(-1)
    lam, order = lam.sort(descending=True)
    V = torch.gather(V, -1, order[:, None, :].repeat(1, V.shape[1], 1))
    return V, lam


def _bingham_dF(lam3):
    """dF/dLam of the last three Bingham parameters (src/fisher/bingham_utils.py:59-73)."""
    full = torch.cat((torch.zeros_like(lam3[:, :1]), lam3), -1)
    with torch.enable_grad():
        leaf = full.detach().requires_grad_(True)
        F = _bingham_F(leaf)
        dF = torch.autograd.grad(F, leaf, torch.ones_like(F))[0]
    return dF[:, 1:]


def bingham_cross_entropy(VB1, lamB1, VB2, lamB2):
    """h(f1, f2), f1 the target (src/fisher/bingham_utils.py:5-32), bug-for-bug:
    ``LamB1.argmax()`` is a flattened argmax (0, because element [0,0] = 0 is the maximum) used
    as a column index, and ``A[:, i]`` takes ROW i of V1^T V2 where the expectation
    E_1[(v2_i . q)^2] would need column i."""
    mu = VB1[:, :, int(lamB1.argmax())]
    VB1, VB2 = VB1[..., 1:], VB2[..., 1:]
    lamB1, lamB2 = lamB1[..., 1:], lamB2[..., 1:]
    pad = lambda l: torch.cat((torch.zeros_like(l[:, :1]), l), -1)
    first = torch.log(_bingham_F(pad(lamB2)))
    second = 0
    A = VB1.transpose(1, 2) @ VB2
    b = (mu[:, None, :] @ VB2).squeeze(1)
    F1 = _bingham_F(pad(lamB1))
    dF1 = _bingham_dF(lamB1)
    for i in range(3):
        tmp = (A[:, i] ** 2 - b[:, i][:, None] ** 2) * (1 / F1[:, None]) * dF1
        second = second + lamB2[:, i] * (b[:, i] ** 2 + tmp.sum(1))
    return first - second


def fisher_ce(A1, A2):
    """Cross entropy of two matrix-Fisher densities through their Bingham forms, A1 the
    target, A2 the prediction (src/fisher/fisher_utils.py:84-99); differentiable in A2
    through torch.svd, the quaternion frame and the quadrature's custom backward."""
    A1 = A1.reshape(-1, 3, 3)
    A2 = A2.reshape(-1, 3, 3)
    V1, lam1 = fisher_to_bingham_frame(A1)
    V2, lam2 = fisher_to_bingham_frame(A2)
    VB1, lamB1 = to_bingham_convention(V1, lam1)
    VB2, lamB2 = to_bingham_convention(V2, lam2)
    ce = bingham_cross_entropy(VB1, lamB1, VB2, lamB2)
    return ce - torch.tensor([np.log(2 * np.pi ** 2)], dtype=ce.dtype)


def _quat_of_rotation(R):
    return p3d.matrix_to_quaternion(R)


def _polar_rotation(p, q):
    """Symmetric bilinear form of the homogeneous quaternion -> rotation map:
    R~(q, q) = |q|^2 R(q);  x^T K(A) y = <A, R~(x, y)>."""
    pw, px, py, pz = p.unbind(-1)
    qw, qx, qy, qz = q.unbind(-1)
    d = lambda a, b, c, e: a * c + b * e          # helper: a*c + b*e
    r00 = pw * qw + px * qx - py * qy - pz * qz
    r11 = pw * qw - px * qx + py * qy - pz * qz
    r22 = pw * qw - px * qx - py * qy + pz * qz
    xy = px * qy + py * qx
    xz = px * qz + pz * qx
    yz = py * qz + pz * qy
    wx = pw * qx + px * qw
    wy = pw * qy + py * qw
    wz = pw * qz + pz * qw
    return torch.stack((r00, xy - wz, xz + wy, xy + wz, r11, yz - wx, xz - wy, yz + wx, r22), -1).reshape(p.shape[:-1] + (3, 3))


def fisher_ce_closed_form(A1, A2):
    """The value of :func:`fisher_ce` and its gradient w.r.t. A2 without autograd (the form the
    CUDA kernel evaluates; derivation in DESIGN.md).  With (U,s,V) the proper SVDs, g = grad logC,
    the frames VB = [q(U V^T), q(U E_1 V^T), q(U E_2 V^T), q(U E_3 V^T)], W = VB1^T VB2,
    gamma = expected squared projections of f1 (gamma_0 + ... + gamma_3 = 1) and
    LamB2 = -2 (s2+s3, s1+s3, s1+s2) of A2:
        CE = log f(s_2) - sum_{i=1..3} LamB2_i [ gamma_0 W_0i^2 + sum_{j=1..3} gamma_j W_ij^2 ]
    The gradient uses first-order perturbation of the eigenpairs (lam_c, v_c) of the 4x4 matrix
    K(A2) (x^T K(A) x = <A, R(x)> is linear in A):  dCE/dA2 = sum_kc G_kc R~(v_k, v_c)."""
    A1 = A1.reshape(-1, 3, 3)
    A2 = A2.reshape(-1, 3, 3)
    b = A1.shape[0]
    u1, s1, v1 = proper_svd(A1)
    u2, s2, v2 = proper_svd(A2)
    _, g1 = log_normaliser_grad(s1)
    half2, g2 = log_normaliser_grad(s2)

    def frame(u, v):
        cols = [_quat_of_rotation(u @ v.transpose(1, 2))]
        for k in range(3):
            E = -torch.eye(3, dtype=u.dtype)
            E[k, k] = 1
            cols.append(_quat_of_rotation(u @ E[None] @ v.transpose(1, 2)))
        return torch.stack(cols, 2)                          # (b,4,4), column c = v_c

    VB1, VB2 = frame(u1, v1), frame(u2, v2)
    W = VB1.transpose(1, 2) @ VB2                            # W[r][c] = v1_r . v2_c
    gam = torch.stack((1 + g1[:, 0] + g1[:, 1] + g1[:, 2], 1 + g1[:, 0] - g1[:, 1] - g1[:, 2],
                       1 - g1[:, 0] + g1[:, 1] - g1[:, 2], 1 - g1[:, 0] - g1[:, 1] + g1[:, 2]), 1) / 4
    lamB = torch.stack((torch.zeros_like(s2[:, 0]), -2 * (s2[:, 1] + s2[:, 2]), -2 * (s2[:, 0] + s2[:, 2]),
                        -2 * (s2[:, 0] + s2[:, 1])), 1)
    W2 = W * W
    phi = gam[:, :1] * W2[:, 0, :] + torch.einsum('bj,bij->bi', gam[:, 1:], W2[:, :, 1:])   # phi_i, i = 0..3 (0 unused)
    ce = torch.log(half2) - (lamB[:, 1:] * phi[:, 1:]).sum(1)

    # eigenvalues of K(A2) in frame order and d s / d lam
    lam = torch.stack((s2.sum(1), s2[:, 0] - s2[:, 1] - s2[:, 2], -s2[:, 0] + s2[:, 1] - s2[:, 2],
                       -s2[:, 0] - s2[:, 1] + s2[:, 2]), 1)
    ds = torch.tensor([[1, 1, -1, -1], [1, -1, 1, -1], [1, -1, -1, 1]], dtype=A2.dtype) / 4   # ds_m / dlam_c
    L_lam = (g2 - 1) @ ds                                    # log f part
    L_lam[:, 1:] -= phi[:, 1:]
    L_lam[:, 0] += phi[:, 1:].sum(1)
    # D[k][c] = v2_k . d(second)/d v2_c   (c = 1..3)
    D = torch.zeros(b, 4, 4, dtype=A2.dtype)
    for c in range(1, 4):
        for k in range(4):
            D[:, k, c] = 2 * gam[:, 0] * lamB[:, c] * W[:, 0, c] * W[:, 0, k] \
                + 2 * gam[:, c] * (lamB[:, 1:] * W[:, 1:, c] * W[:, 1:, k]).sum(1)
    grad = torch.zeros_like(A2)
    for c in range(4):
        grad = grad + L_lam[:, c, None, None] * _polar_rotation(VB2[:, :, c], VB2[:, :, c])
        for k in range(4):
            if k != c:
                G = -D[:, k, c] / (lam[:, c] - lam[:, k])
                grad = grad + G[:, None, None] * _polar_rotation(VB2[:, :, k], VB2[:, :, c])
    return ce, grad


def fisher_nll_grad_closed_form(A, R, overreg):
    """d nll / dA = -R + overreg * U diag(g1,g2,g3) V^T with the proper (U,s,V)
    (SURVEY.md A.4); per-sample gradient, no batch mean."""
    A = A.reshape(-1, 3, 3)
    u, s, v = proper_svd(A)
    _, g = log_normaliser_grad(s)
    return -R.reshape(-1, 3, 3) + overreg * (u * g[:, None, :]) @ v.transpose(1, 2)


# -------------------------------------------------------------------------- a11
def laplace_power(A, X):
    """-sqrt(max(s1+s2+s3*sign(det A) - tr(A^T X), 1e-8))
    (src/laplace/rotation_laplace.py:140-173)."""
    prod = torch.matmul(torch.transpose(A, -1, -2), X)
    tr = prod[..., 0, 0] + prod[..., 1, 1] + prod[..., 2, 2]
    S = torch.linalg.svdvals(A)
    S = torch.cat((S[..., :-1], S[..., -1:] * torch.sign(torch.det(A))[..., None]), -1)
    return -torch.sqrt(torch.clamp_min(S.sum(-1) - tr, LAPLACE_EPS))


def fisher_power(A, X):
    """tr(A^T X)  (src/laplace/rotation_laplace.py:118-137)."""
    prod = torch.matmul(torch.transpose(A, -1, -2), X)
    return prod[..., 0, 0] + prod[..., 1, 1] + prod[..., 2, 2]


def _grid_log_const(power_fn, A, grids, laplace):
    """log of the grid-sum normaliser, max-shifted
    (src/laplace/rotation_laplace.py:41-72)."""
    n = grids.shape[0]
    p = power_fn(A[:, None], grids[None])
    c = p.max(dim=-1)[0]
    e = torch.exp(p - c[:, None])
    if laplace:
        e = e / (-p)
    return c + torch.log(e.sum(1) * (1 / n))


def grid_log_pdf(fn_type, A, x, grids, broadcast=False):
    """src/laplace/rotation_laplace.py:76-99."""
    laplace = "RLaplace" in fn_type
    power_fn = laplace_power if fn_type == "RLaplace" else fisher_power
    logF = _grid_log_const(power_fn, A, grids, laplace)
    if x.shape[0] == grids.shape[0] or broadcast:
        x, A, logF = x[None], A[:, None], logF[:, None]
    p = power_fn(A, x)
    if laplace:
        return -logF + p - torch.log(-p)
    return -logF + p


def analytical_mode(pred, fn_type="RLaplace"):
    """U diag(1,1,det(U V^T)) V^T  (src/laplace/rotation_laplace.py:102-115)."""
    U, _, Vt = torch.linalg.svd(pred)
    with torch.no_grad():
        sgn = torch.det(torch.matmul(U, Vt))
    d = torch.diag_embed(torch.stack((torch.ones_like(sgn), torch.ones_like(sgn), sgn), -1))
    return U @ d @ Vt, sgn


def laplace_nll(fn_type, pred, gt, grids):
    """(losses (b,), mode (b,3,3))  (src/laplace/rotation_laplace.py:24-34)."""
    pred = pred.reshape(-1, 3, 3)
    losses = -grid_log_pdf(fn_type, pred, gt, grids)
    mode, _ = analytical_mode(pred, fn_type)
    return losses, mode


# -------------------------------------------------------------------------- a12
def pool_threshold(entropies, left_ratio):
    """Ascending sort, k = int(n * left_ratio) (Python float product, truncation),
    threshold = sorted[k]  (src/agent.py:403-407).  ``left_ratio >= 1`` raises
    IndexError exactly like the reference's numpy indexing."""
    e = np.array(entropies, dtype=np.float32, copy=True).reshape(-1)
    e.sort()
    k = int(len(e) * left_ratio)
    return e[k], k


def keep_mask(entropy, conf_thres):
    """Strict ``<`` and kept fraction (src/agent.py:148-150)."""
    mask = entropy < conf_thres
    return mask, mask.sum() / len(mask)


# -------------------------------------------------------------------- a13 .. a16
def euler_from_matrices(R, full_range=False):
    """(b,3,3) -> (b,3) radians (pitch, yaw, roll)  (src/utils.py:232-260).  The
    reference's per-sample Python loop (:240-242) only negates ``sy`` where
    R00 < 0 in full-range mode; vectorised here with identical results."""
    r00, r10 = R[:, 0, 0], R[:, 1, 0]
    sy = torch.sqrt(r00 * r00 + r10 * r10)
    singular = (sy < 1e-6).to(R.dtype)
    if full_range:
        sy = torch.where(r00 < 0, -sy, sy)
    x = torch.atan2(R[:, 2, 1], R[:, 2, 2])
    y = torch.atan2(-R[:, 2, 0], sy)
    z = torch.atan2(r10, r00)
    xs = torch.atan2(-R[:, 1, 2], R[:, 1, 1])
    zs = r10 * 0
    keep = 1 - singular
    return torch.stack((x * keep + xs * singular, y * keep + y * singular,
                        z * keep + zs * singular), dim=1)


def update_ema_variables(net, ema_net, is_ema, alpha, global_step, eman=False):
    """src/agent.py:277-299 restated line by line (torch CPU): the warm-up rule, then the EMAN blend over
    the state_dict or the in-place EMA over the parameters."""
    if is_ema:
        alpha = min(1 - 1 / (global_step + 1), alpha)                     # :279-280
    else:
        alpha = 0                                                         # :283
    with torch.no_grad():
        if eman:                                                          # :286-293
            state_dict_main, state_dict_ema = net.state_dict(), ema_net.state_dict()
            for (k_main, v_main), (k_ema, v_ema) in zip(state_dict_main.items(), state_dict_ema.items()):
                assert k_main == k_ema, "state_dict names are different!"
                assert v_main.shape == v_ema.shape, "state_dict shapes are different!"
                if 'num_batches_tracked' in k_ema:
                    v_ema.copy_(v_main)
                else:
                    v_ema.copy_(v_ema * alpha + (1. - alpha) * v_main)
        else:                                                             # :295-298
            for ema_param, param in zip(ema_net.parameters(), net.parameters()):
                ema_param.data.mul_(alpha).add_(param.detach(), alpha=1 - alpha)
    return alpha


def rotate_aug_adjust(pred_weak, aug_rot_mat, train_labeled):
    """src/agent.py:110-119, the same torch ops in the same order."""
    mat = pred_weak.clone().view(-1, 3, 3)
    if train_labeled == "DAD3DHeads":
        return torch.matmul(aug_rot_mat, mat).view(-1, 9)
    rot_180 = torch.tensor([[1, 0, 0], [0, -1, 0], [0, 0, -1]], dtype=pred_weak.dtype)
    trans = torch.matmul(rot_180, mat.transpose(1, 2))
    trans = torch.matmul(aug_rot_mat, trans)
    return torch.matmul(rot_180, trans).transpose(1, 2).reshape(-1, 9)


def euler_dad_degrees(R):
    """(b,3,3) -> (b,3) float64 degrees (pitch, yaw, roll) of a DAD-trained model: the per-sample
    loop of eval.py:66-74 (scipy ``as_euler("xyz")`` of the transpose, then ``limit_angle``)."""
    from scipy.spatial.transform import Rotation
    out = []
    for rot_mat in R.detach().cpu().numpy():
        angle = Rotation.from_matrix(np.transpose(rot_mat)).as_euler("xyz", degrees=True)
        roll, pitch, yaw = list(map(limit_angle, [angle[2], angle[0] - 180, angle[1]]))
        out.append([pitch, yaw, roll])
    return np.array(out)


def geodesic_deg(pred, gt):
    """rad2deg(so3_relative_angle)  (src/agent.py:449-451, eval.py:88-89)."""
    return torch.rad2deg(p3d.so3_relative_angle(pred, gt))


def err_deg_from_matrices(pred, gt, gt_euler=None):
    """src/agent.py:447-455."""
    if gt_euler is None:
        return geodesic_deg(pred, gt)
    pd = euler_from_matrices(pred, full_range=False) * 180 / np.pi
    return torch.mean(torch.abs(pd - gt_euler), dim=-1)


def frobenius_identity_distance(pred, gt):
    """per-sample ||I - R_pd R_gt^T||_F  (eval.py:93-98; numpy on fp32 data)."""
    D = (pred @ gt.transpose(1, 2)).detach().cpu().numpy()
    return np.array([np.linalg.norm(np.eye(3) - D[i], "fro") for i in range(D.shape[0])])


def euler_mae_summary(err_deg):
    """Column means and the mean of the per-sample 3-angle mean (eval.py:125-133)."""
    p, y, r = err_deg[:, 0], err_deg[:, 1], err_deg[:, 2]
    return float(np.mean(p)), float(np.mean(y)), float(np.mean(r)), float(np.mean((p + y + r) / 3))


def rot_from_euler(x, y, z):
    """R = Rz(z) Ry(y) Rx(x), radians (src/utils.py:204-226)."""
    cx, sx, cy, sy, cz, sz = math.cos(x), math.sin(x), math.cos(y), math.sin(y), math.cos(z), math.sin(z)
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return Rz.dot(Ry.dot(Rx))


def limit_angle(angle, pi=180.0):
    """Wrap degrees into [-180, 180] (src/utils.py:289-300)."""
    if angle < -pi:
        angle = angle + (-2 * (int(angle / pi) // 2)) * pi
    if angle > pi:
        angle = angle - (2 * ((int(angle / pi) + 1) // 2)) * pi
    return angle
