"""Does CPU-side torch work (the cpu_baseline leg) slow the eager small-batch steps that follow in the same process?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semiuhpe_b200
from semiuhpe_b200.agent import _quat_to_matrix
from semiuhpe_b200.fisher.fisher_utils import vmf_loss
semiuhpe_b200.set_error_checking(False)
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(5)
A32 = 10 * torch.randn(32, 9, device=dev, generator=gen)
R32 = _quat_to_matrix(torch.nn.functional.normalize(torch.randn(32, 4, device=dev, generator=gen), dim=1)).contiguous()
leaf = A32.clone().requires_grad_(True)

def c1():
    leaf.grad = None
    loss, _ = vmf_loss(leaf, R32, overreg=1.025)
    loss.mean().backward()

def timed(reps=300):
    for _ in range(10): c1()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): c1()
    b.record(); torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / reps

print(f"fresh process                      c1 = {timed():7.1f} us   threads={torch.get_num_threads()}")
x = torch.randn(4096, 4096)
for n in (os.cpu_count(), 8, 2, 1):
    torch.set_num_threads(n)
    t0 = time.perf_counter(); (x @ x).sum().item(); dt = time.perf_counter() - t0
    print(f"after CPU matmul with {n:2d} threads ({dt:.2f} s)  c1 = {timed():7.1f} us")
from oracle import so3_oracle as orc
torch.set_num_threads(8)
Ac = 10 * torch.randn(2048, 9)
orc.fisher_entropy(Ac)
print(f"after oracle fisher_entropy (8 thr)     c1 = {timed():7.1f} us")
time.sleep(1.0)
print(f"... 1 s later                          c1 = {timed():7.1f} us")
torch.set_num_threads(1)
print(f"after set_num_threads(1)               c1 = {timed():7.1f} us")
