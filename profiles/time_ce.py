"""fisher_CE timing (value + gradient) on the C ABI, CUDA events.  usage: python profiles/time_ce.py [log2_n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from semiuhpe_b200 import _capi

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(5)
A1 = 10 * torch.randn(n, 9, device=dev, generator=gen)
A2 = A1 + 2 * torch.randn(n, 9, device=dev, generator=gen)
ce, grad = torch.empty(n, device=dev), torch.empty(n, 9, device=dev)
work = torch.empty(_capi.FISHER_CE_WORKSPACE_FLOATS * n, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
P, lib = _capi.ptr, _capi.lib()
call = lambda: _capi.check(lib.suhpe_fisher_ce_f32(P(A1), P(A2), n, 26, None, P(ce), P(grad), P(work), P(status), _capi.stream()), "ce")
for _ in range(2):
    call()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    call()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(f"fisher_CE n=2^{n.bit_length() - 1}: {ms:.3f} ms  {n / ms / 1e3:.1f} M pairs/s  "
      f"{n * 2 * 69120 / ms / 1e9:.2f} TFLOP/s at 2 x 69,120 FLOP/pair  status={int(status.item())} ce.mean={ce.mean().item():.5f}")
