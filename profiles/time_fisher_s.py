"""Per-run-type cost of the quadrature: time K2 on GIVEN singular values that make every family one
long run of a single type (cut disabled).  python profiles/time_fisher_s.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semiuhpe_b200
from semiuhpe_b200 import _capi
n = 1 << 22
dev = torch.device("cuda:0")
lib, P, S = _capi.lib(), _capi.ptr, _capi.stream
out = torch.empty(n, device=dev); G = torch.empty(n, 3, device=dev)
cases = {"all-LL (1000,600,300)": (1000., 600., 300.), "all-SS (1.5,1,0.5)": (1.5, 1.0, 0.5), "generic (25,13,6)": (25., 13., 6.),
         "generic (25,13,-6)": (25., 13., -6.), "kappa 40 (40,40,40)": (40., 40., 40.)}
for name, sv in cases.items():
    Sv = torch.tensor(sv, device=dev).repeat(n, 1).contiguous()
    f = lambda: _capi.check(lib.suhpe_fisher_from_s_f32(P(Sv), n, 0, P(out), P(G), None, None, S()), "s")
    for _ in range(3): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    cyc = ms * 1e-3 * 1.965e9 * 592 / n
    print(f"{name:28s} {ms:8.3f} ms  {cyc:8.1f} SMSP-cycles/sample  ({cyc/24:.1f} per 64 nodes)  logC={out[0].item():.5f}")
