"""K4 timing on the C ABI with preallocated outputs (CUDA events, current stream).
usage: python profiles/time_k4.py [n_pairs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from semiuhpe_b200 import _capi
from semiuhpe_b200.agent import _quat_to_matrix

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(0)
rot = lambda m: _quat_to_matrix(torch.nn.functional.normalize(torch.randn(m, 4, device=dev, generator=gen), dim=1)).contiguous()
Rp, Rg = rot(n), rot(n)
ge = (torch.rand(n, 3, device=dev, generator=gen) * 2 - 1) * 89
new = lambda *s: torch.empty(s, device=dev)
geo, frob, eul, err, mae = new(n), new(n), new(n, 3), new(n, 3), new(n)
sums = torch.zeros(8, dtype=torch.float64, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
P, lib = _capi.ptr, _capi.lib()


def run(args, nbytes, name):
    call = lambda: _capi.check(lib.suhpe_so3_metrics_f32(P(Rp), *args, P(status), _capi.stream()), name)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        call()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"{name:34s} {ms:7.4f} ms  {n * nbytes / ms / 1e6:8.1f} GB/s at {nbytes} B/pair", flush=True)


run((P(Rg), P(ge), n, 0, P(geo), P(frob), None, P(err), None, P(sums)), 104, "geo+frob+abs_err+sums (C4)")
run((P(Rg), None, n, 0, P(geo), None, None, None, None, None), 76, "geodesic only")
run((P(Rg), P(ge), n, 0, None, None, None, None, P(mae), None), 88, "euler MAE only")
run((None, None, n, 1, None, None, P(eul), None, None, None), 48, "euler angles only (full range)")
run((P(Rg), P(ge), n, 0, P(geo), P(frob), P(eul), P(err), P(mae), P(sums)), 120, "everything")
