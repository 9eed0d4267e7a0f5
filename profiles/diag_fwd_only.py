"""Diagnostic: is the forward-only launch (NFAM=1) bit-identical to the full launch (NFAM=3)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semiuhpe_b200 import _ops
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(2024)
scales = torch.tensor([0.3, 1.0, 3.0, 5.0, 10.0, 20.0]).repeat_interleave(1000)
A = (torch.randn(len(scales), 9, generator=gen) * scales[:, None]).to(dev)
for n in (6000, 200000):
    a = A if n == 6000 else (10 * torch.randn(n, 9, generator=gen)).to(dev)
    for bits in (26, 0):
        full = _ops.fisher_fused(a, None, 1.0, nll=True, entropy=True, S=True, logC=True, cut_bits=bits)
        fwd = _ops.fisher_fused(a, None, 1.0, nll=True, S=True, logC=True, cut_bits=bits)
        bad = (full["nll"] != fwd["nll"])
        print(f"n={n} bits={bits}: nll mismatches {int(bad.sum())}  S mismatches {int((full['S'] != fwd['S']).any(1).sum())}  "
              f"logC mismatches {int((full['logC'] != fwd['logC']).sum())}  max|d nll|={(full['nll'] - fwd['nll']).abs().max().item():.3e}")
