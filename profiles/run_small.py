"""Driver for ncu: K2 at small batch sizes (device duration per launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semiuhpe_b200 import _ops
from semiuhpe_b200.agent import _quat_to_matrix
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
for n in (32, 128, 1024, 8192, 65536):
    A = 10 * torch.randn(n, 9, device=dev, generator=gen)
    R = _quat_to_matrix(torch.nn.functional.normalize(torch.randn(n, 4, device=dev, generator=gen), dim=1)).contiguous()
    for _ in range(3):
        _ops.fisher_fused(A, R, 1.025, nll=True, grad=True, entropy=True)
torch.cuda.synchronize()
