"""K2L fwd+bwd time against the batch size (4608-point grid): python profiles/sweep_k2l.py  -- used to check launch_laplace's choice
between the warp kernel and the stream kernel's plain and cluster forms (r02ae / r02af: builds with -DSUHPE_K2L_PACK_SAMPLES=0 / 1 / 2 while the
round-1 point-packed stream kernel was still in the tree; now -DSUHPE_K2L_FORCE_STREAM=1 forces the stream kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semiuhpe_b200
from semiuhpe_b200 import _ops
from semiuhpe_b200.agent import _quat_to_matrix

semiuhpe_b200.set_error_checking(False)
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(3)
rot = lambda m: _quat_to_matrix(torch.nn.functional.normalize(torch.randn(m, 4, device=dev, generator=gen), dim=1)).contiguous()
sms = torch.cuda.get_device_properties(dev).multi_processor_count
grid = rot(4608)
nmax = sms * 512 * 9
A, R = 5 * torch.randn(nmax, 9, device=dev, generator=gen), rot(nmax)
sizes = [1024, 2048, 4096, 8192, 16384, 32768] + [sms * 256 * k for k in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16, 18)]
for n in sizes:
    k = n / (sms * 256)
    for _ in range(3):
        _ops.laplace_nll(A[:n], R[:n], grid, grad=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _ops.laplace_nll(A[:n], R[:n], grid, grad=True)
    e1.record()
    torch.cuda.synchronize()
    print(f"n = {k:6.2f} x {sms} x 256 = {n:7d}: {e0.elapsed_time(e1) / 10:7.3f} ms", flush=True)
