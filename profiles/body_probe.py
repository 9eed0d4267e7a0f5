"""K2 pass-body probe: cycles per 128-node pass per SM sub-partition for the run bodies in isolation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semiuhpe_b200 import _capi
dev = torch.device("cuda:0"); lib = _capi.lib()
sink = torch.zeros(4, device=dev)
iters, blocks = 20000, 148
names = {1: "LL full", 1 + 4: "LL no-LDS", 1 + 8: "LL no-MUFU", 1 + 16: "LL no-mask", 1 + 28: "LL FMA only", 2: "SS full", 2 + 28: "SS FMA only"}
for v, name in names.items():
    f = lambda: _capi.check(lib.suhpe_fp32_probe(_capi.ptr(sink), 100 + v, iters, blocks, _capi.stream()), "probe")
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); f(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    cyc = ms * 1e-3 * 1.965e9 / (iters * 4)     # 16 warps per SM = 4 per sub-partition
    print(f"{name:14s} {ms:8.3f} ms   {cyc:6.1f} cycles per 128-node pass per SMSP  (FMA-pipe floor 92 for LL, 88 for SS)")
