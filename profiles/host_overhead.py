"""Where does the host time of a small-batch step go?  python profiles/host_overhead.py"""
import cProfile, pstats, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semiuhpe_b200
from semiuhpe_b200 import _capi, _ops
from semiuhpe_b200.agent import _quat_to_matrix, ssl_loss, unsupervised_terms
from semiuhpe_b200.fisher.fisher_utils import vmf_loss

semiuhpe_b200.set_error_checking(False)
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(5)
rot = lambda m: _quat_to_matrix(torch.nn.functional.normalize(torch.randn(m, 4, device=dev, generator=gen), dim=1)).contiguous()
A32, R32 = 10 * torch.randn(32, 9, device=dev, generator=gen), rot(32)
A128 = 10 * torch.randn(128, 9, device=dev, generator=gen)
leaf32 = A32.clone().requires_grad_(True)
leaf128 = (A128 + 0.5).requires_grad_(True)
lib, P, S = _capi.lib(), _capi.ptr, _capi.stream
nll, grad = torch.empty(32, device=dev), torch.empty(32, 9, device=dev)


def wall(fn, reps=2000):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    host = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) / reps
    return host * 1e6, total * 1e6


def c1():
    leaf32.grad = None
    loss, _ = vmf_loss(leaf32, R32, overreg=1.025)
    loss.mean().backward()


def c1_one():
    leaf32.grad = None
    ssl_loss(leaf32, R32, overreg=1.025)[0].backward()


def c2_one():
    leaf32.grad = None
    leaf128.grad = None
    ssl_loss(leaf32, R32, A128, leaf128, -3.6)[0].backward()


cases = {
    "raw ctypes K2 launch (b=32)": lambda: lib.suhpe_fisher_fused_f32(P(A32), P(R32), 32, 1.025, 26, P(nll), P(grad), None, None, None, None, None, None, None, S()),
    "_ops.fisher_fused nll+grad+rot": lambda: _ops.fisher_fused(A32, R32, 1.025, nll=True, grad=True, rot=True),
    "_ops.scale_rows": lambda: _ops.scale_rows(grad, nll),
    "torch: x.mean()": lambda: nll.mean(),
    "torch: empty(32,9)": lambda: torch.empty(32, 9, device=dev),
    "vmf_loss forward only (no_grad)": lambda: vmf_loss(A32, R32, overreg=1.025),
    "c1 vmf_loss fwd+bwd": c1,
    "c1 one call fwd+bwd": c1_one,
    "c2 ssl one call fwd+bwd": c2_one,
}
for name, fn in cases.items():
    h, t = wall(fn)
    print(f"{name:40s} host {h:8.1f} us   host+drain {t:8.1f} us")

for name, fn in (("c1", c1), ("c2_one", c2_one)):
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(500):
        fn()
    pr.disable()
    torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
    print(f"---- cProfile {name} (500 iterations)")
    print("\n".join(l for l in s.getvalue().splitlines() if l.strip())[:3500])
