"""Device-time the fused Fisher kernel alone (CUDA events, 2^23 pairs): python profiles/time_fisher.py [log2n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semiuhpe_b200
from semiuhpe_b200 import _capi, _ops
from semiuhpe_b200.agent import _quat_to_matrix

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 23)
semiuhpe_b200.set_error_checking(False)
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
A = 10 * torch.randn(n, 9, device=dev, generator=gen)
R = _quat_to_matrix(torch.nn.functional.normalize(torch.randn(n, 4, device=dev, generator=gen), dim=1)).reshape(n, 9).contiguous()
nll, grad, ent = torch.empty(n, device=dev), torch.empty(n, 9, device=dev), torch.empty(n, device=dev)
hist = torch.zeros(2048, dtype=torch.int64, device=dev)
lib, P, S = _capi.lib(), _capi.ptr, _capi.stream
BITS = [26]
def run():
    _capi.check(lib.suhpe_fisher_fused_f32(P(A), P(R), n, 1.025, BITS[0], P(nll), P(grad), None, P(ent), None, None, None, P(hist), None, S()), "f")
for bits in [int(b) for b in os.environ.get("BITS", "0,26").split(",")]:
    BITS[0] = bits
    for _ in range(3): run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): run()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"cut_bits={bits:2d}  {ms:8.3f} ms  {n/ms/1e3:9.1f} Mrot/s  {n*69120/ms/1e9:6.2f} TFLOP/s(alg)  nll.mean={nll.mean().item():.6f} ent.mean={ent.mean().item():.6f}")
# forward-only NLL (no gradient, no entropy): the normaliser family alone
BITS[0] = 26
def run_fwd():
    _capi.check(lib.suhpe_fisher_fused_f32(P(A), P(R), n, 1.025, 26, P(nll), None, None, None, None, None, None, None, None, S()), "f")
for _ in range(3): run_fwd()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run_fwd()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"forward-only NLL  {ms:8.3f} ms  {n/ms/1e3:9.1f} Mrot/s  nll.mean={nll.mean().item():.6f}")
