"""K2L timing (C3: 2^20 samples x 4608 grid points) on the C ABI, CUDA events.
usage: python profiles/time_k2l.py [log2_n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from semiuhpe_b200 import _capi
from semiuhpe_b200.agent import _quat_to_matrix

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 20)
N = 4608
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(5)
rot = lambda m: _quat_to_matrix(torch.nn.functional.normalize(torch.randn(m, 4, device=dev, generator=gen), dim=1)).contiguous()
grid, A, R = rot(N), 5 * torch.randn(n, 9, device=dev, generator=gen), rot(n)
new = lambda *s: torch.empty(s, device=dev)
nll, grad, mode = new(n), new(n, 9), new(n, 9)
status = torch.zeros(1, dtype=torch.int32, device=dev)
P, lib = _capi.ptr, _capi.lib()
call = lambda: _capi.check(lib.suhpe_laplace_nll_f32(P(A), P(R), n, P(grid), N, P(nll), P(grad), P(mode), None, P(status), _capi.stream()), "laplace")
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
print(f"K2L n=2^{n.bit_length() - 1} N={N} unroll={os.environ.get('SUHPE_LAP_UNROLL', '1')}: {ms:.3f} ms  {n / ms / 1e3:.1f} M rot/s  "
      f"{n * N * 50 / ms / 1e9:.2f} TFLOP/s at 50 FLOP/pair  nll.sum={nll.double().sum().item():.6f}")
# forward-only (no gradient requested: validation under no_grad)
call_f = lambda: _capi.check(lib.suhpe_laplace_nll_f32(P(A), P(R), n, P(grid), N, P(nll), None, P(mode), None, P(status), _capi.stream()), "laplace")
ref = nll.clone()
for _ in range(2):
    call_f()
torch.cuda.synchronize()
a.record()
for _ in range(5):
    call_f()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(f"K2L forward only: {ms:.3f} ms  {n / ms / 1e3:.1f} M rot/s  max |nll - nll(full launch)| = {(ref - nll).abs().max().item():.2e}")
