"""The per-sample DEVICE arithmetic (semiuhpe_b200/csrc/so3_math.cuh) compiled for the
host (tests/emul) against the oracle/golden vectors -- catches algorithmic errors in
the SVD, the quadrature and the metrics without a GPU.  MUFU approximations are
replaced by libm on the host, so the tight parity gate is the GPU suite."""
import ctypes
import os

import numpy as np
import pytest
import torch

from helpers import ATOL, RTOL, assert_close, grad_rel_err, no_worse_than_reference
from oracle import so3_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope="module")
def emul(built):
    return ctypes.CDLL(os.path.join(HERE, "emul", "libemul.so"))


def run_fisher(emul, A, R, overreg, cut_bits=26):
    A = np.ascontiguousarray(A.reshape(-1, 9), np.float32)
    R = np.ascontiguousarray(R.reshape(-1, 9), np.float32)
    n = len(A)
    o = {k: np.zeros(s, np.float32) for k, s in dict(nll=n, grad=(n, 9), rot=(n, 9), ent=n, logC=n, S=(n, 3), G=(n, 3)).items()}
    emul.emul_fisher(P(A), P(R), ctypes.c_long(n), ctypes.c_float(overreg), ctypes.c_int(cut_bits), P(o["nll"]), P(o["grad"]), P(o["rot"]),
                     P(o["ent"]), P(o["logC"]), P(o["S"]), P(o["G"]))
    return o


def test_quadrature_nodes_bitwise(emul):
    x = np.zeros(512, np.float32)
    emul.emul_quad_nodes(P(x))
    ref, _ = orc.quad_nodes(torch.float32)
    np.testing.assert_array_equal(x, ref.numpy().ravel())


def test_bessel_polynomials(emul):
    a = np.concatenate([np.linspace(0, 3.75, 2000), np.linspace(3.75, 400, 4000), [1e-8, 3.7499, 3.7501]]).astype(np.float32)
    out = np.zeros_like(a)
    emul.emul_i0e(P(a), ctypes.c_long(len(a)), P(out))
    ref = orc.i0e(torch.from_numpy(a)).numpy()
    np.testing.assert_allclose(out, ref, rtol=1e-6, atol=0)


def test_fisher_against_golden(emul, golden):
    g = golden("fisher")
    o = run_fisher(emul, g["A"], g["R"], float(g["overreg"]))
    names = g["names"]
    regime = np.isin(names, ["generic1", "generic10", "generic30", "realistic", "neardegenerate"])
    assert_close(o["nll"][regime], g["nll"][regime], RTOL, ATOL, "nll")
    assert_close(o["ent"][regime], g["entropy"][regime], RTOL, ATOL, "entropy")
    assert_close(o["logC"][regime], g["logC"][regime], RTOL, ATOL, "logC")
    np.testing.assert_allclose(o["S"][regime], g["S"][regime], rtol=2e-6, atol=2e-6)
    well = np.isin(names, ["generic1", "generic10", "generic30", "realistic"])
    assert grad_rel_err(o["grad"][well], g["grad"][well]).max() < 1e-5
    assert grad_rel_err(o["grad"][names == "neardegenerate"], g["grad"][names == "neardegenerate"]).max() < 1e-4
    # rotation: conditioning ~ eps*s1/(s2+s3); generic rows are well conditioned
    assert np.abs(o["rot"][well] - g["Rest"][well].reshape(-1, 9)).max() < 1e-5
    # large singular values: no worse than the reference against exact arithmetic
    A64 = torch.from_numpy(g["A"]).double()
    ent64 = orc.fisher_entropy_closed_form(A64).numpy()
    edge = ~regime
    # (at s ~ 300 the quadrature itself has broken down and s*(1-g) amplifies rounding: 3e-5 of the reference's fp32 value also passes)
    edge_ok = no_worse_than_reference(o["ent"][edge], g["entropy"][edge], ent64[edge], slack=2.0, floor=2e-6)
    edge_ok |= np.abs(o["ent"][edge] - g["entropy"][edge]) <= 3e-5 * np.abs(g["entropy"][edge])
    assert edge_ok.all()
    zero = list(names).index("zero")
    assert o["nll"][zero] == 0 and o["ent"][zero] == 0
    np.testing.assert_array_equal(o["rot"][zero], np.eye(3, dtype=np.float32).ravel())
    np.testing.assert_allclose(o["grad"][zero], -np.eye(3).ravel(), atol=1e-6)


def test_negligible_node_cut_is_below_rounding(emul, golden):
    """Skipping the provably negligible node prefix (cut_bits=26, the default) must not move any
    output by more than fp32 rounding of the full 512-node evaluation (cut_bits=0)."""
    g = golden("fisher")
    full = run_fisher(emul, g["A"], g["R"], float(g["overreg"]), cut_bits=0)
    cut = run_fisher(emul, g["A"], g["R"], float(g["overreg"]), cut_bits=26)
    assert_close(cut["logC"], full["logC"], 2e-7, 2e-7, "logC")
    assert_close(cut["nll"], full["nll"], 2e-7, 1e-6, "nll")
    assert np.abs(cut["G"] - full["G"]).max() < 6e-7
    # entropy = log f + sum_j s_j (1 - g_j) amplifies the few-ulp regrouping noise of g by |s|
    assert (np.abs(cut["ent"] - full["ent"]) <= 2e-6 + 6e-7 * np.abs(full["S"]).sum(1)).all()
    rng = np.random.default_rng(5)
    A = (rng.standard_normal((4000, 9)) * rng.choice([0.1, 1, 5, 10, 30, 100], (4000, 1))).astype(np.float32)
    R = np.tile(np.eye(3, dtype=np.float32).ravel(), (4000, 1))
    full, cut = run_fisher(emul, A, R, 1.025, 0), run_fisher(emul, A, R, 1.025, 26)
    assert np.abs(cut["G"] - full["G"]).max() < 6e-7
    assert_close(cut["logC"], full["logC"], 2e-7, 2e-7, "logC random")


def cut_info(emul, S, bits=26):
    S = np.ascontiguousarray(S, np.float32)
    n = len(S)
    cut, valid, slots = np.zeros((n, 3), np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
    emul.emul_cut_info(P(S), ctypes.c_long(n), ctypes.c_int(bits), P(cut), P(valid), P(slots))
    return cut, valid, slots


def test_negligible_node_cut_bound_holds_in_float64(emul):
    """The claim behind the cut (cut_threshold / cut_index in so3_math.cuh): for every family the
    trapezoid mass of the skipped prefix [0,cut) is below 2^-bits of the normaliser sum F.  Checked
    with scipy's float64 I0e on the kernel's own fp32 nodes over random spectra from 1e-2 to 1e3
    (both signs of s3) and a structured sweep through degenerate / sign-flipped / switch-point
    spectra; also that the cut does something on the benchmark distribution."""
    from scipy.special import i0e
    i = np.arange(512, dtype=np.float32)
    x = i * np.float32(2.0 / 511.0) - np.float32(1)
    u, v = (np.float32(1) - x).astype(np.float64), (np.float32(1) + x).astype(np.float64)
    w = np.ones(512)
    w[0] = w[-1] = 0.5

    def worst_mass(S, bits):
        S = np.ascontiguousarray(S, np.float32)
        cut, valid, slots = cut_info(emul, S, bits)
        S = S.astype(np.float64)
        fams = [(S[:, 2], S[:, 1], S[:, 0] + S[:, 2]), (S[:, 2], S[:, 0], S[:, 1] + S[:, 2]), (S[:, 1], S[:, 0], S[:, 1] + S[:, 2])]
        ys = lambda lo, hi, c: i0e(np.abs(0.5 * (hi - lo))[:, None] * u) * i0e(np.abs(0.5 * (hi + lo))[:, None] * v) * np.exp(-c[:, None] * u)
        F = (ys(*fams[0]) * w).sum(1)
        worst = 0.0
        for f, fam in enumerate(fams):
            cs = np.cumsum(ys(*fam) * w, 1)
            k = cut[:, f]
            assert (k >= 0).all() and (k <= 511).all()
            mass = np.where(k > 0, cs[np.arange(len(S)), np.maximum(k - 1, 0)], 0.0)
            worst = max(worst, float((mass / F * 2.0 ** bits).max()))
        return worst, cut, valid, slots

    rng = np.random.default_rng(11)
    for scale in [0.01, 0.3, 1, 3, 5, 10, 30, 100, 300, 1000]:
        A = torch.from_numpy(rng.standard_normal((3000, 3, 3)) * scale)
        U, S, Vh = torch.linalg.svd(A)
        S[:, 2] *= torch.det(U @ Vh)
        for bits in (20, 26):
            worst, *_ = worst_mass(S.numpy(), bits)
            assert worst <= 1.0, (scale, bits, worst)
    sweep = [(s1, s1 * r2, s1 * r2 * r3)
             for s1 in [0, 1e-3, 0.5, 1, 2, 3.75, 5, 7.5, 10, 20, 50, 100, 300, 1000, 1e4]
             for r2 in [0, 0.01, 0.3, 0.5, 0.9, 0.999, 1]
             for r3 in [-1, -0.999, -0.9, -0.5, -0.01, 0, 0.01, 0.5, 0.9, 0.999, 1]]
    worst, *_ = worst_mass(np.array(sweep), 26)
    assert worst <= 1.0, worst
    # non-finite spectra: nothing is cut, nothing crashes
    cut, _, _ = cut_info(emul, np.array([[np.nan, 1, 1], [np.inf, 2, 1], [3e38, 3e38, -3e38]], np.float32))
    assert (cut == 0).all()
    # cut off: every node is evaluated
    assert (cut_info(emul, np.array(sweep), 0)[0] == 0).all()
    # the benchmark distribution (A = 10 randn): more than a third of the 1536 nodes are provably negligible
    torch.manual_seed(0)
    U, S, Vh = torch.linalg.svd(10 * torch.randn(20000, 3, 3))
    S[:, 2] *= torch.det(U @ Vh)
    worst, cut, valid, slots = worst_mass(S.numpy(), 26)
    assert worst <= 1.0 and valid.mean() < 1024


def run_laplace(emul, A, R, grid, L):
    A = np.ascontiguousarray(A.reshape(-1, 9), np.float32)
    R = np.ascontiguousarray(R.reshape(-1, 9), np.float32)
    grid = np.ascontiguousarray(grid.reshape(-1, 9), np.float32)
    n = len(A)
    o = {k: np.zeros(s, np.float32) for k, s in dict(nll=n, grad=(n, 9), mode=(n, 9), logF=n).items()}
    emul.emul_laplace(P(A), P(R), ctypes.c_long(n), P(grid), ctypes.c_int(len(grid)), ctypes.c_int(L),
                      P(o["nll"]), P(o["grad"]), P(o["mode"]), P(o["logF"]))
    return o


def test_laplace_setup_mixed_precision(emul):
    """K2L's per-sample set-up (mode R*, T = s1 + s2 + s3 signed) starts its fp64 iteration from the V of the fp32
    one (Newton-Schulz step, two polishing sweeps).  Against seven fp64 sweeps from scratch: T to 2e-12 relative --
    it feeds the fp64 cancellation T - <A,R_gt> -- and R* to fp32 rounding, on random, scaled, ill-conditioned
    (cond 1e3 and 1e6), nearly rank-deficient, nearly degenerate and negative-determinant matrices."""
    rng = np.random.default_rng(5)
    n = 4000
    kinds = []
    for kind in range(8):
        A = rng.standard_normal((n, 3, 3)).astype(np.float32)
        if kind == 1: A *= 50
        if kind == 2: A *= 1e-3
        if kind == 3: A[:, :, 2] *= 1e-3; A *= 10
        if kind == 4: A[:, :, 2] *= 1e-6; A *= 10
        if kind == 5: A[:, :, 1] = A[:, :, 0] * (1 + 1e-4 * rng.standard_normal((n, 3))).astype(np.float32); A *= 5
        if kind == 6: A = (3 * np.eye(3) + 1e-5 * rng.standard_normal((n, 3, 3))).astype(np.float32)
        if kind == 7: A = (np.diag([2, 2, -2]) + 0.1 * rng.standard_normal((n, 3, 3))).astype(np.float32)
        kinds.append(A)
    A = np.ascontiguousarray(np.concatenate(kinds).reshape(-1, 9))
    m = len(A)
    Rs, T = np.empty((m, 9), np.float32), np.empty(m, np.float64)
    Rs64, T64 = np.empty((m, 9), np.float64), np.empty(m, np.float64)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    emul.emul_laplace_setup(P(A), ctypes.c_long(m), P(Rs), P(T), P(Rs64), P(T64))
    assert np.isfinite(T).all() and np.isfinite(Rs).all()
    assert (np.abs(T - T64) <= 2e-12 * np.abs(T64) + 1e-300).all()
    assert np.abs(Rs - Rs64).max() < 6e-8


def test_laplace_against_golden(emul, golden):
    """K2L arithmetic (fp64 per-sample set-up, two-level fp32 grid sums) in both decompositions the
    kernel uses: parity with the reference's fp32 output, closer to exact arithmetic than the
    reference itself, and the two decompositions agree."""
    g = golden("laplace")
    A64 = torch.from_numpy(g["A"]).double().requires_grad_(True)
    nll64, _ = orc.laplace_nll("RLaplace", A64, torch.from_numpy(g["R"]).double(), torch.from_numpy(g["grids"]).double())
    nll64.sum().backward()
    e64, g64 = nll64.detach().numpy(), A64.grad.numpy().reshape(-1, 9)
    outs = {L: run_laplace(emul, g["A"], g["R"], g["grids"], L) for L in (1, 32)}
    for L, o in outs.items():
        assert_close(o["nll"], g["nll"], 5e-5, 5e-5, f"laplace nll L={L}")
        assert grad_rel_err(o["grad"], g["grad"]).max() < 2e-3
        assert no_worse_than_reference(o["nll"], g["nll"], e64, 2.0, 5e-6).all()
        ours_rel, ref_rel = grad_rel_err(o["grad"], g64), grad_rel_err(g["grad"], g64)
        assert ours_rel.max() <= ref_rel.max()
        well = g["A"].reshape(-1, 9).std(1) > 0.5
        assert np.abs(o["mode"] - g["mode"].reshape(-1, 9))[well].max() < 2e-5
    assert grad_rel_err(outs[1]["grad"], outs[32]["grad"]).max() < 2e-4
    assert_close(outs[1]["nll"], outs[32]["nll"], 1e-5, 1e-5, "decompositions")


def test_svd_properties(emul):
    rng = np.random.default_rng(0)
    n = 20000
    A = np.concatenate([rng.standard_normal((n, 3, 3)) * s for s in (1e-20, 1e-3, 1.0, 30.0, 1e20)]).astype(np.float32)
    A9 = np.ascontiguousarray(A.reshape(-1, 9))
    m = len(A9)
    R, S, U, V = (np.zeros((m, k), np.float32) for k in (9, 3, 9, 9))
    ok = np.zeros(m, np.int32)
    emul.emul_proper_svd(P(A9), ctypes.c_long(m), P(R), P(S), P(U), P(V), P(ok))
    assert ok.all()
    U3, V3, R3 = U.reshape(-1, 3, 3).astype(np.float64), V.reshape(-1, 3, 3).astype(np.float64), R.reshape(-1, 3, 3).astype(np.float64)
    scale = np.abs(A).reshape(m, -1).max(1)
    rec = np.einsum("nij,nj,nkj->nik", U3, S.astype(np.float64), V3)
    assert (np.abs(rec - A).reshape(m, -1).max(1) / scale).max() < 3e-6
    assert np.abs(np.einsum("nij,nkj->nik", R3, R3) - np.eye(3)).max() < 3e-6
    assert np.abs(np.linalg.det(R3) - 1).max() < 5e-6          # det(R) = +1 always
    assert np.abs(np.linalg.det(U3) - 1).max() < 5e-6 and np.abs(np.linalg.det(V3) - 1).max() < 5e-6
    # sign convention: sign(S2) == sign(det A) (reference: det(U V^T) of LAPACK's factors)
    detA = np.linalg.det(A.astype(np.float64) / scale[:, None, None])
    clear = np.abs(detA) > 1e-4
    assert np.array_equal(np.sign(S[clear, 2]), np.sign(detA[clear]))
    sv = np.linalg.svd(A.astype(np.float64), compute_uv=False)
    assert (np.abs(np.abs(S) - sv).max(1) / scale).max() < 3e-6
    # non-finite input is flagged
    bad = np.full((2, 9), np.nan, np.float32); bad[1] = np.inf
    ok2 = np.ones(2, np.int32)
    emul.emul_proper_svd(P(bad), ctypes.c_long(2), P(R), P(S), P(U), P(V), P(ok2))
    assert not ok2.any()


def test_keys_follow_numpy_sort_order(emul, golden):
    e = golden("select")["ties"].copy()
    k = np.zeros(len(e), np.uint32)
    back = np.zeros(len(e), np.float32)
    emul.emul_keys(P(e), ctypes.c_long(len(e)), P(k), P(back))
    order = np.argsort(k, kind="stable")
    np.testing.assert_array_equal(np.sort(e), e[order] + 0.0)       # NaN last, -0 == +0
    np.testing.assert_array_equal(back[~np.isnan(e)], e[~np.isnan(e)] + 0.0)
    assert np.isnan(back[np.isnan(e)]).all()


def test_metrics_against_golden(emul, golden):
    g = golden("metrics")
    n = len(g["R_pd"])
    for Rp, full, key in ((g["R_pd"], 0, "euler_pd"), (g["R_full"], 0, "euler_full_false"), (g["R_full"], 1, "euler_full_true")):
        Rp9 = np.ascontiguousarray(Rp.reshape(-1, 9)); Rg9 = np.ascontiguousarray(g["R_gt"].reshape(-1, 9))
        ge = np.ascontiguousarray(g["gt_euler"])
        geo, frob, mae = (np.zeros(n, np.float32) for _ in range(3))
        eul = np.zeros((n, 3), np.float32); ok = np.zeros(n, np.int32)
        emul.emul_metrics(P(Rp9), P(Rg9), P(ge), ctypes.c_long(n), full, P(geo), P(frob), P(eul), P(mae), P(ok))
        np.testing.assert_allclose(eul, g[key], rtol=0, atol=2e-6)
        assert ok.all()
        if key == "euler_pd":
            np.testing.assert_allclose(mae, g["mae"], rtol=1e-5, atol=2e-4)
            np.testing.assert_allclose(geo, g["geodesic_deg"], rtol=1e-4, atol=2e-3)
            np.testing.assert_allclose(frob, g["frob"], rtol=1e-4, atol=1e-6)
        else:
            np.testing.assert_allclose(geo, g["geodesic_deg_full"], rtol=1e-4, atol=2e-3)
            np.testing.assert_allclose(frob, g["frob_full"], rtol=1e-5, atol=1e-6)


def test_run_boundaries_are_exact(emul):
    """The three uniform-type runs the kernel derives per family reproduce the reference's
    per-node choice of Bessel polynomial (|a| <= 3.75 on the fp32 product) for every one of
    the 512 nodes, for singular values from 0 to 1e6."""
    rng = np.random.default_rng(1)
    S = np.abs(rng.standard_normal((200000, 3))).astype(np.float32) * (10.0 ** rng.uniform(-3, 3, (200000, 1))).astype(np.float32)
    S = -np.sort(-S, axis=1)
    S[:, 2] *= rng.choice([-1.0, 1.0], len(S)).astype(np.float32)
    S[:50] = 0
    S[50:100, 1:] = S[50:100, :1]           # equal singular values -> fd = 0
    S[100:150] = np.float32(3.75) * np.array([1.0, 1.0, 1.0], np.float32)
    S = np.ascontiguousarray(S, np.float32)
    assert emul.emul_check_runs(P(S), ctypes.c_long(len(S))) == 0
    # the flat pass-item list the warp replays covers every kept node exactly once, with the right type
    for bits in (0, 26):
        assert emul.emul_check_items(P(S[:60000]), ctypes.c_long(60000), ctypes.c_int(bits)) == 0


def ulp_err(got, exact64):
    ulp = np.spacing(np.abs(exact64).astype(np.float32)).astype(np.float64)
    return np.abs(got.astype(np.float64) - exact64) / ulp


def test_atan2_accuracy_and_special_cases(emul):
    """K4's branch-free atan2 (so3_math.cuh): <= 2 ulp against double atan2 on rotation-like
    arguments of every octant and magnitude, IEEE results on the axes and for signed zeros."""
    rng = np.random.default_rng(0)
    n = 2_000_000
    y = rng.uniform(-1, 1, n).astype(np.float32)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    scale = (10.0 ** rng.uniform(-37, 37, n)).astype(np.float32)      # ratio is what matters
    y[n // 2:] *= scale[n // 2:]; x[n // 2:] *= scale[n // 2:]
    tiny = (10.0 ** rng.uniform(-9, 0, n // 4)).astype(np.float32)    # near-axis angles
    y[:n // 4] *= tiny
    out = np.zeros(n, np.float32)
    emul.emul_atan2(P(y), P(x), ctypes.c_long(n), P(out))
    exact = np.arctan2(y.astype(np.float64), x.astype(np.float64))
    ok = np.isfinite(y) & np.isfinite(x) & (np.abs(exact) > 1e-36)
    assert ulp_err(out[ok], exact[ok]).max() <= 2.0
    ys = np.array([0.0, -0.0, 0.0, -0.0, 1.0, -1.0, 0.0, -0.0, 1.0, 1.0, np.nan, 1.0, 1e-30, 3e-39], np.float32)
    xs = np.array([1.0, 1.0, -1.0, -1.0, 0.0, 0.0, 0.0, -0.0, -0.0, 1.0, 1.0, np.nan, -1e-30, 1e-39], np.float32)
    got = np.zeros(len(ys), np.float32)
    emul.emul_atan2(P(ys), P(xs), ctypes.c_long(len(ys)), P(got))
    want = np.arctan2(ys, xs)
    np.testing.assert_array_equal(np.isnan(got), np.isnan(want))
    fin = ~np.isnan(want)
    np.testing.assert_allclose(got[fin], want[fin], rtol=3e-7, atol=0)
    np.testing.assert_array_equal(np.signbit(got[fin]), np.signbit(want[fin]))


def test_degrees_conversion_is_the_reference_rounding(emul):
    """euler*180/np.pi in fp32 (two roundings, src/agent.py:452): the Markstein form is bit-identical."""
    rng = np.random.default_rng(1)
    rad = np.concatenate([rng.uniform(-np.pi, np.pi, 2_000_000), [0.0, np.pi, -np.pi, 1e-20, 1e-7]]).astype(np.float32)
    out = np.zeros_like(rad)
    emul.emul_rad_to_deg(P(rad), ctypes.c_long(len(rad)), P(out))
    ref = (torch.from_numpy(rad) * 180 / np.pi).numpy()
    np.testing.assert_array_equal(out, ref)


def test_fisher_ce_against_golden(emul, golden):
    """The kernels' closed-form fisher_CE (SVD, quadrature, frames, eigenpair perturbation) on the host
    against the reference's value and autograd gradient."""
    g = golden("fisher_ce")
    n = len(g["A1"])
    a1 = np.ascontiguousarray(g["A1"].reshape(n, 9))
    a2 = np.ascontiguousarray(g["A2"].reshape(n, 9))
    ce, grad = np.zeros(n, np.float32), np.zeros((n, 9), np.float32)
    emul.emul_fisher_ce(P(a1), P(a2), ctypes.c_long(n), ctypes.c_int(26), P(ce), P(grad))
    # value: 1e-5, or no further from fp64 than the reference's own fp32 result (large |CE| rows)
    ok, _ = np.abs(ce - g["ce"]) <= ATOL + RTOL * np.abs(g["ce"]), None
    assert (ok | no_worse_than_reference(ce, g["ce"], g["ce64"])).all()
    err64 = grad_rel_err(grad, g["grad64"])
    ref_err64 = grad_rel_err(g["grad"], g["grad64"])
    stable = g["names"] != "student_neardegenerate"
    assert (err64[stable] <= np.maximum(2 * ref_err64[stable], 2e-5)).all()
    # near-degenerate students: 1/(s_i - s_j) amplifies fp32 rounding sample by sample; compare the class
    assert err64[~stable].max() <= max(2 * ref_err64[~stable].max(), 1e-4)


def angle_diff_deg(a, b):
    d = np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))
    return np.minimum(d, np.abs(d - 360.0))


def test_dad_euler_against_golden(emul, golden):
    """DAD-trained Euler convention (eval.py:66-74: scipy as_euler("xyz") of R^T + limit_angle),
    fixture generated by executing those lines; fp32 rotations against scipy's float64."""
    g = golden("dad_euler")
    R = np.ascontiguousarray(g["R"].reshape(-1, 9))
    out = np.zeros((len(R), 3), np.float32)
    emul.emul_euler_dad(P(R), ctypes.c_long(len(R)), P(out))
    assert angle_diff_deg(out, g["euler_deg"]).max() < 1e-3
    near = g["names"] == "near_frontal"
    assert angle_diff_deg(out[near], g["euler_deg"][near]).max() < 1e-4
    assert np.abs(out).max() <= 180.0
