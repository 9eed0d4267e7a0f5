"""Small-batch latency of K2 through the C ABI (BASELINE configs 1-2: 32 labeled / 128 unlabeled):
back-to-back launches (device time per launch) and the Python-level calls the agent makes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semiuhpe_b200
from semiuhpe_b200 import _capi
from semiuhpe_b200.agent import _quat_to_matrix, dynamic_entropy_filter
from semiuhpe_b200.fisher.fisher_utils import vmf_loss, fisher_entropy

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
lib, P, S = _capi.lib(), _capi.ptr, _capi.stream
for n in (32, 128, 1024, 8192):
    A = 10 * torch.randn(n, 9, device=dev, generator=gen)
    R = _quat_to_matrix(torch.nn.functional.normalize(torch.randn(n, 4, device=dev, generator=gen), dim=1)).reshape(n, 9).contiguous()
    nll, grad, ent = torch.empty(n, device=dev), torch.empty(n, 9, device=dev), torch.empty(n, device=dev)
    run = lambda: _capi.check(lib.suhpe_fisher_fused_f32(P(A), P(R), n, 1.025, 26, P(nll), P(grad), None, P(ent), None, None, None, None, None, S()), "f")
    for _ in range(10): run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(200): run()
    b.record(); torch.cuda.synchronize()
    print(f"K2 n={n:5d}: {a.elapsed_time(b) / 200 * 1e3:7.1f} us per launch (back to back)")
from semiuhpe_b200 import _ops
grid = _quat_to_matrix(torch.nn.functional.normalize(torch.randn(4608, 4, device=dev, generator=gen), dim=1)).contiguous()
for n in (32, 128, 160, 1024, 8192, 32768):
    A = 5 * torch.randn(n, 9, device=dev, generator=gen)
    R = _quat_to_matrix(torch.nn.functional.normalize(torch.randn(n, 4, device=dev, generator=gen), dim=1)).contiguous()
    run = lambda: _ops.laplace_nll(A, R, grid, grad=True, mode=True)
    for _ in range(10): run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100): run()
    b.record(); torch.cuda.synchronize()
    print(f"K2L n={n:5d} x 4608: {a.elapsed_time(b) / 100 * 1e3:7.1f} us per launch (back to back)")
A32 = 10 * torch.randn(32, 9, device=dev, generator=gen)
R32 = _quat_to_matrix(torch.nn.functional.normalize(torch.randn(32, 4, device=dev, generator=gen), dim=1)).contiguous()
A128 = 10 * torch.randn(128, 9, device=dev, generator=gen)
def wall(fn, reps=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / reps * 1e6
def c1():
    leaf = A32.clone().requires_grad_(True)
    loss, _ = vmf_loss(leaf, R32, overreg=1.025)
    loss.mean().backward()
for chk in (True, False):
    semiuhpe_b200.set_error_checking(chk)
    print(f"error_checking={chk}: vmf_loss fwd+bwd b=32 {wall(c1):6.1f} us   fisher_entropy b=128 {wall(lambda: fisher_entropy(A128)):6.1f} us   "
          f"dynamic_entropy_filter b=128 {wall(lambda: dynamic_entropy_filter(A128, 0.95, return_threshold=False)):6.1f} us")
