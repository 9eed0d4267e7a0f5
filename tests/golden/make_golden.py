"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE.

Run in the build container only (needs ``/root/reference``)::

    python tests/golden/make_golden.py

The reference (hnuzhy/SemiUHPE) has no tests or golden vectors of its own
(SURVEY.md section 4), so the pins are outputs of its unmodified code on seeded
inputs, imported through ``oracle/ref_shim.py``.  Everything written here is
fp32 exactly as the reference produced it; gradients are of ``nll.sum()`` so
each row is the per-sample gradient.

Files
-----
fisher.npz   vmf_loss / KL_Fisher / batch_torch_A_to_R / fisher_log_pdf /
             fisher_entropy / logC_F (+ its backward) on generic, realistic and
             edge-case A (src/fisher/*.py)
laplace.npz  NLL_loss("RLaplace") fwd+bwd, analytical_mode, log_pdf("RFisher")
             on the reference's own 4608-point grid (src/laplace/rotation_laplace.py,
             src/laplace/eq_grids2.npy; the grid is stored as a fixture INPUT)
select.npz   pool threshold by the literal lines src/agent.py:405-407, masks by :148
fisher_ce.npz  fisher_CE(A1 target, A2 prediction) and its autograd gradient w.r.t. A2
             (src/fisher/fisher_utils.py:84-99; quaternion frames through the restated
             pytorch3d matrix_to_quaternion -- the value and the gradient do not depend on its
             branch or sign choices), plus fp64 anchors from the oracle's restatement
dad_euler.npz  the DAD-trained Euler convention of eval.py:66-74 (scipy as_euler + limit_angle)
metrics.npz  compute_euler_angles_from_rotation_matrices (src/utils.py:232),
             compute_err_deg_from_matrices (src/agent.py:447-455, extracted by AST),
             so3_relative_angle via the restated pytorch3d (PARITY UNPINNED for that
             one function), Frobenius loop of eval.py:93-98, limit_angle
"""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shim  # noqa: E402
from oracle import pytorch3d_restated as p3d  # noqa: E402


def random_rotations(n, gen):
    q, _ = torch.linalg.qr(torch.randn(n, 3, 3, generator=gen))
    d = torch.det(q)
    q[:, :, 2] *= d[:, None]
    return q.contiguous()


def fisher_inputs():
    gen = torch.Generator().manual_seed(20240117)
    blocks, names = [], []
    for scale in (1.0, 10.0, 30.0):
        blocks.append(scale * torch.randn(128, 3, 3, generator=gen))
        names += [f"generic{int(scale)}"] * 128
    kappa = 5 + 45 * torch.rand(128, 1, 1, generator=gen)
    blocks.append(kappa * random_rotations(128, gen) + 0.5 * torch.randn(128, 3, 3, generator=gen))
    names += ["realistic"] * 128
    # near-degenerate singular values under random frames (gradient tolerance 1e-4 class)
    Ud, Vd = random_rotations(48, gen), random_rotations(48, gen)
    sv = torch.tensor([[10.0, 10.0 - 1e-3, 3.0], [10.0, 5.0, 5.0 - 1e-3], [8.0, 8.0 - 1e-4, 8.0 - 2e-4],
                       [20.0, 3.0, 1e-3], [20.0, 3.0, -1e-3], [12.0, 6.0, -5.9]]).repeat(8, 1)
    blocks.append(Ud @ torch.diag_embed(sv) @ Vd.transpose(1, 2))
    names += ["neardegenerate"] * 48
    eye = torch.eye(3)
    edge = [("zero", torch.zeros(3, 3)), ("eye1e-3", 1e-3 * eye),
            ("diag10_5_0", torch.diag(torch.tensor([10.0, 5.0, 0.0]))),
            ("diag10_5_-2", torch.diag(torch.tensor([10.0, 5.0, -2.0]))),
            ("diag10_10_3", torch.diag(torch.tensor([10.0, 10.0, 3.0]))),
            ("diag300_200_100", torch.diag(torch.tensor([300.0, 200.0, 100.0])))]
    for k in (1, 2, 5, 10, 20, 30, 50, 100, 200, 500):
        edge.append((f"eye{k}", float(k) * eye))
    blocks.append(torch.stack([m for _, m in edge]))
    names += [n for n, _ in edge]
    A = torch.cat(blocks).contiguous()
    R = random_rotations(A.shape[0], gen)
    n_edge = len(edge)
    R[-n_edge:] = eye  # edge rows use R_gt = I like SURVEY appendix C
    return A, R, np.array(names)


def make_fisher(ref):
    fu, tnf = ref.fisher_utils, ref.torch_norm_factor
    A, R, names = fisher_inputs()
    overreg = 1.025
    leaf = A.clone().requires_grad_(True)
    nll, Rest = fu.vmf_loss(leaf.view(-1, 9), R, overreg=overreg)
    nll.sum().backward()
    ent = fu.fisher_entropy(A.clone())
    logpdf = fu.fisher_log_pdf(A.clone(), R)
    u, s, v = ref.bbf.proper_svd(A.clone())
    sl = s.clone().requires_grad_(True)
    logc = tnf.logC_F(sl)
    logc.sum().backward()
    np.savez_compressed(
        os.path.join(HERE, "fisher.npz"), A=A.numpy(), R=R.numpy(), names=names,
        overreg=np.float32(overreg), nll=nll.detach().numpy(), grad=leaf.grad.numpy(),
        Rest=Rest.detach().numpy(), entropy=ent.detach().numpy(), logpdf=logpdf.detach().numpy(),
        S=s.numpy(), logC=logc.detach().numpy(), dlogC=sl.grad.numpy())
    print("fisher.npz", A.shape[0], "samples")


def make_fisher_ce(ref):
    """Teacher / student parameter pairs: student near the teacher (the training regime),
    unrelated frames, and students with close singular values (gradient tolerance class 1e-4)."""
    from oracle import so3_oracle as orc
    fu = ref.fisher_utils
    gen = torch.Generator().manual_seed(20240612)
    t, p, names = [], [], []
    for scale in (1.0, 10.0, 30.0):
        a1 = scale * torch.randn(96, 3, 3, generator=gen)
        t.append(a1); p.append(a1 + 0.25 * scale * torch.randn(96, 3, 3, generator=gen))
        names += [f"near{int(scale)}"] * 96
    kappa = 5 + 45 * torch.rand(96, 1, 1, generator=gen)
    R0 = random_rotations(96, gen)
    t.append(kappa * R0 + 0.5 * torch.randn(96, 3, 3, generator=gen))
    p.append((kappa * (0.6 + 0.8 * torch.rand(96, 1, 1, generator=gen))) * R0 + 1.5 * torch.randn(96, 3, 3, generator=gen))
    names += ["realistic"] * 96
    t.append(10 * torch.randn(96, 3, 3, generator=gen)); p.append(10 * torch.randn(96, 3, 3, generator=gen))
    names += ["unrelated"] * 96
    Ud, Vd = random_rotations(48, gen), random_rotations(48, gen)
    sv = torch.tensor([[10.0, 10.0 - 1e-2, 3.0], [10.0, 5.0, 5.0 - 1e-2], [8.0, 8.0 - 1e-2, 8.0 - 2e-2],
                       [20.0, 3.0, 1e-2], [20.0, 3.0, -1e-2], [12.0, 6.0, -5.99]]).repeat(8, 1)
    p.append(Ud @ torch.diag_embed(sv) @ Vd.transpose(1, 2)); t.append(8 * torch.randn(48, 3, 3, generator=gen))
    names += ["student_neardegenerate"] * 48
    A1, A2 = torch.cat(t).contiguous(), torch.cat(p).contiguous()
    leaf = A2.clone().requires_grad_(True)
    ce = fu.fisher_CE(A1.clone(), leaf)
    ce.sum().backward()
    leaf64 = A2.double().requires_grad_(True)
    ce64 = orc.fisher_ce(A1.double(), leaf64)
    ce64.sum().backward()
    np.savez_compressed(os.path.join(HERE, "fisher_ce.npz"), A1=A1.numpy(), A2=A2.numpy(), names=np.array(names),
                        ce=ce.detach().numpy(), grad=leaf.grad.numpy(),
                        ce64=ce64.detach().numpy(), grad64=leaf64.grad.numpy())
    print("fisher_ce.npz", A1.shape[0], "pairs")


def make_laplace(ref):
    rl = ref.rotation_laplace
    grids = torch.from_numpy(np.load(ref.grid_path("eq_grids2.npy")))
    gen = torch.Generator().manual_seed(777)
    A = torch.cat([s * torch.randn(32, 3, 3, generator=gen) for s in (1.0, 5.0, 20.0)])
    kappa = 5 + 45 * torch.rand(32, 1, 1, generator=gen)
    A = torch.cat([A, kappa * random_rotations(32, gen) + 0.5 * torch.randn(32, 3, 3, generator=gen)]).contiguous()
    R = random_rotations(A.shape[0], gen)
    leaf = A.clone().requires_grad_(True)
    nll, mode = rl.NLL_loss("RLaplace", leaf, R, grids)
    nll.sum().backward()
    _, s3 = rl.analytical_mode(A.clone(), "RLaplace")
    lf = rl.log_pdf("RFisher", A.clone(), R, grids)
    np.savez_compressed(
        os.path.join(HERE, "laplace.npz"), A=A.numpy(), R=R.numpy(), grids=grids.numpy(),
        nll=nll.detach().numpy(), grad=leaf.grad.numpy(), mode=mode.detach().numpy(),
        s3sign=s3.numpy(), rfisher_logpdf=lf.detach().numpy())
    print("laplace.npz", A.shape[0], "samples x", grids.shape[0], "grid points")


def _agent_source():
    return open(os.path.join(ref_shim.REFERENCE_ROOT, "src", "agent.py")).read()


def reference_pool_threshold(entropies, left_ratio):
    """Execute the literal three lines src/agent.py:405-407."""
    lines = _agent_source().splitlines()[404:407]
    assert lines[0].strip() == "entropy_all.sort()", lines
    assert lines[2].strip() == "entropy_thre = entropy_all[index]", lines
    ns = {"entropy_all": np.array(entropies, dtype=np.float32, copy=True), "int": int, "len": len,
          "self": types.SimpleNamespace(config=types.SimpleNamespace(left_ratio=left_ratio))}
    exec("\n".join(l.strip() for l in lines), ns)
    return ns["entropy_thre"], ns["index"]


def reference_err_deg(ref):
    """``SSLAgent.compute_err_deg_from_matrices`` cut out of src/agent.py by AST
    (the module itself needs configargparse/timm/pytorchcv to import)."""
    tree = ast.parse(_agent_source())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "SSLAgent")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "compute_err_deg_from_matrices")
    fn.decorator_list = []
    ns = {"torch": torch, "np": np, "trans": p3d,
          "compute_euler_angles_from_rotation_matrices": ref.euler_from_matrices}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "agent.py", "exec"), ns)
    return ns["compute_err_deg_from_matrices"]


def make_select(ref):
    gen = torch.Generator().manual_seed(99)
    A = 10 * torch.randn(4096, 3, 3, generator=gen)
    ent = ref.fisher_utils.fisher_entropy(A).detach().numpy().astype(np.float32)
    out = {"entropy": ent, "A": A.numpy()}
    ratios = np.array([0.95, 0.75, 0.5, 0.999, 0.0, 0.05])
    thr, idx = zip(*(reference_pool_threshold(ent, float(r)) for r in ratios))
    out.update(ratios=ratios, thresholds=np.array(thr, np.float32), indices=np.array(idx, np.int64))
    # teacher batch of config 2: 128 samples, left_ratio 0.95 -> k=121
    t128, k128 = reference_pool_threshold(ent[:128], 0.95)
    out.update(thr128=np.float32(t128), k128=np.int64(k128),
               mask128=(torch.from_numpy(ent[:128]) < float(t128)).numpy())
    # ties, infinities and NaNs (np.sort places NaN last)
    rng = np.random.default_rng(5)
    ties = np.round(rng.normal(-5, 1, 5000), 1).astype(np.float32)
    ties[::97] = np.nan
    ties[5::501] = np.inf
    ties[7::499] = -np.inf
    ties[11::250] = 0.0
    ties[12::250] = -0.0
    tt, ti = zip(*(reference_pool_threshold(ties, float(r)) for r in ratios))
    out.update(ties=ties, ties_thresholds=np.array(tt, np.float32), ties_indices=np.array(ti, np.int64))
    np.savez_compressed(os.path.join(HERE, "select.npz"), **out)
    print("select.npz", ent.shape[0], "entropies")


def make_metrics(ref):
    gen = torch.Generator().manual_seed(4242)
    n = 2048
    # AFLW2000-style labels: Euler degrees (pitch, yaw, roll) in (-89.99, 89.99)
    gt_euler = (torch.rand(n, 3, generator=gen) * 2 - 1) * 89.99
    rad = torch.deg2rad(gt_euler).double().numpy()
    R_gt = torch.from_numpy(np.stack([ref.get_6DRepNet_Rot(*row) for row in rad]).astype(np.float32))
    omega = torch.randn(n, 3, generator=gen) * np.deg2rad(5.0)
    K = torch.zeros(n, 3, 3)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 2] = -omega[:, 2], omega[:, 1], -omega[:, 0]
    K = K - K.transpose(1, 2)
    R_pd = (R_gt @ torch.matrix_exp(K)).contiguous()
    R_full = random_rotations(n, gen)  # full-range poses incl. R00 < 0
    # gimbal-lock rows (sy < 1e-6) and exact identity pairs
    lock = torch.tensor([[0.0, 0.0, 1.0], [0.0, 1.0, 0.0], [-1.0, 0.0, 0.0]])
    R_full[0] = lock
    R_full[1] = lock.t() @ torch.diag(torch.tensor([1.0, -1.0, -1.0]))
    R_pd[0] = R_gt[0]
    err_deg = reference_err_deg(ref)
    out = dict(
        R_pd=R_pd.numpy(), R_gt=R_gt.numpy(), R_full=R_full.numpy(), gt_euler=gt_euler.numpy(),
        euler_pd=ref.euler_from_matrices(R_pd, full_range=False, use_gpu=False).numpy(),
        euler_full_false=ref.euler_from_matrices(R_full, full_range=False, use_gpu=False).numpy(),
        euler_full_true=ref.euler_from_matrices(R_full, full_range=True, use_gpu=False).numpy(),
        mae=err_deg(R_pd, R_gt, gt_euler).numpy(),
        geodesic_deg=err_deg(R_pd, R_gt).numpy(),
        geodesic_deg_full=err_deg(R_full, R_gt).numpy(),
    )
    # eval.py:93-98 verbatim semantics
    D = (R_pd @ R_gt.transpose(1, 2)).numpy()
    out["frob"] = np.array([np.linalg.norm(np.eye(3) - D[i], "fro") for i in range(n)])
    D = (R_full @ R_gt.transpose(1, 2)).numpy()
    out["frob_full"] = np.array([np.linalg.norm(np.eye(3) - D[i], "fro") for i in range(n)])
    angles = np.array([-725.5, -540.0, -360.0, -181.0, -180.0, -90.0, 0.0, 90.0, 180.0, 181.0, 359.0, 360.0, 540.0, 725.5])
    out["limit_in"] = angles
    out["limit_out"] = np.array([ref.limit_angle(a) for a in angles])
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)
    print("metrics.npz", n, "pairs")


def make_dad_euler(ref):
    """eval.py:66-74 executed literally (scipy + the reference's limit_angle) on full-range rotations,
    near-frontal heads of a DAD-trained model (pitch offset by 180) and the gimbal-lock poses."""
    from scipy.spatial.transform import Rotation
    gen = torch.Generator().manual_seed(4242)
    R = random_rotations(2048, gen)
    ang = (torch.rand(1024, 3, generator=gen) * 2 - 1) * torch.tensor([60.0, 80.0, 50.0])
    rx180 = torch.diag(torch.tensor([1.0, -1.0, -1.0]))
    near = torch.stack([torch.from_numpy(ref.get_6DRepNet_Rot(*np.deg2rad(a.numpy()))).float() for a in ang])
    near = (near @ rx180).transpose(1, 2).contiguous()
    lock = torch.stack([torch.from_numpy(Rotation.from_euler("xyz", [t, s * 90.0, 20.0], degrees=True).as_matrix().T).float()
                        for t in (-170.0, -30.0, 0.0, 45.0, 120.0) for s in (-1.0, 1.0)])
    Rall = torch.cat((R, near, lock)).contiguous()
    import warnings
    rows = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                    # scipy warns at gimbal lock
        for rot_mat in Rall.numpy():
            rot_mat_2 = np.transpose(rot_mat)
            angle = Rotation.from_matrix(rot_mat_2).as_euler("xyz", degrees=True)
            roll, pitch, yaw = list(map(ref.limit_angle, [angle[2], angle[0] - 180, angle[1]]))
            rows.append([pitch, yaw, roll])
    names = np.array(["random"] * 2048 + ["near_frontal"] * 1024 + ["gimbal_lock"] * len(lock))
    np.savez_compressed(os.path.join(HERE, "dad_euler.npz"), R=Rall.numpy(), euler_deg=np.array(rows), names=names)
    print("dad_euler.npz", len(Rall), "rotations")


if __name__ == "__main__":
    torch.set_num_threads(1)  # fixed reduction order inside the reference's torch ops
    ref = ref_shim.load()
    makers = dict(fisher=make_fisher, fisher_ce=make_fisher_ce, dad_euler=make_dad_euler, laplace=make_laplace, select=make_select, metrics=make_metrics)
    for name in (sys.argv[1:] or list(makers)):          # `make_golden.py fisher_ce` regenerates one file
        makers[name](ref)
