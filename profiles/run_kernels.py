"""Tiny driver used under ncu: launches each hot kernel a few times on synthetic data.
usage: python profiles/run_kernels.py [fisher|laplace|metrics|select|all] [log2_n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import semiuhpe_b200
from semiuhpe_b200 import _ops
from semiuhpe_b200.agent import _quat_to_matrix, entropy_threshold, entropy_mask

which = sys.argv[1] if len(sys.argv) > 1 else "all"
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 21)
semiuhpe_b200.set_error_checking(False)
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
rot = lambda m: _quat_to_matrix(torch.nn.functional.normalize(torch.randn(m, 4, device=dev, generator=gen), dim=1)).contiguous()
A, R = 10 * torch.randn(n, 9, device=dev, generator=gen), rot(n)
ws = _ops.SelectWorkspace(dev)
for _ in range(1 if which == "others" else 3):
    if which in ("fisher", "all"):
        out = _ops.fisher_fused(A, R, 1.025, nll=True, grad=True, entropy=True, hist=ws.hist[0])
    if which in ("select", "others", "all"):
        e = torch.randn(n, device=dev, generator=gen)
        entropy_mask(e, entropy_threshold(e, 0.95))
    if which in ("laplace", "others", "all"):
        m = min(n, 1 << 18)
        grid = rot(4608)
        _ops.laplace_nll(A[:m] * 0.5, R[:m], grid, grad=True, mode=True)
    if which in ("ce", "others", "all"):
        m = min(n, 1 << 20)
        _ops.fisher_ce(A[:m], A[:m] + 2 * torch.randn(m, 9, device=dev, generator=gen), grad=True)
    if which in ("metrics", "others", "all"):
        ge = torch.rand(n, 3, device=dev, generator=gen) * 90
        _ops.so3_metrics(rot(n), R, ge, geo=True, frob=True, abs_err=True, sums=True)
torch.cuda.synchronize()
print("done", which, n)
