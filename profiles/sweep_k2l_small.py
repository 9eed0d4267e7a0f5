"""K2L fwd+bwd launch time for training-sized batches (4608-point grid): python profiles/sweep_k2l_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semiuhpe_b200
from semiuhpe_b200 import _capi
from semiuhpe_b200.agent import _quat_to_matrix

semiuhpe_b200.set_error_checking(False)
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(3)
rot = lambda m: _quat_to_matrix(torch.nn.functional.normalize(torch.randn(m, 4, device=dev, generator=gen), dim=1)).contiguous()
grid = rot(4608)
lib, P, S = _capi.lib(), _capi.ptr, _capi.stream
for n in (1, 32, 128, 160, 256, 512, 1024, 2048, 4096):
    A, R = 5 * torch.randn(n, 9, device=dev, generator=gen), rot(n)
    nll, grad, mode = torch.empty(n, device=dev), torch.empty(n, 9, device=dev), torch.empty(n, 9, device=dev)
    run = lambda: lib.suhpe_laplace_nll_f32(P(A), P(R), n, P(grid), 4608, P(nll), P(grad), P(mode), None, None, S())
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"K2L n = {n:5d} x 4608: {e0.elapsed_time(e1) / 50 * 1e3:8.1f} us per launch", flush=True)
