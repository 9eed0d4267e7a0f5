"""Small driver for compute-sanitizer (memcheck / racecheck): every round-2 kernel path once, small sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semiuhpe_b200
from semiuhpe_b200 import _ops
from semiuhpe_b200.agent import ssl_loss, validation_terms, _quat_to_matrix
from semiuhpe_b200.fisher.fisher_utils import fisher_CE, batch_torch_A_to_R, vmf_loss
semiuhpe_b200.set_error_checking(False)
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
rot = lambda m: _quat_to_matrix(torch.nn.functional.normalize(torch.randn(m, 4, device=dev, generator=g), dim=1)).contiguous()
for n in (1, 33, 1000, 5003):
    A = 10 * torch.randn(n, 9, device=dev, generator=g)
    hist = torch.zeros(2048, dtype=torch.int64, device=dev)
    _ops.fisher_fused(A, rot(n), 1.025, nll=True, grad=True, rot=True, entropy=True, hist=hist)
    _ops.fisher_fused(A, None, 1.0, nll=True)
    keep = torch.rand(n, device=dev, generator=g) < 0.7
    s = (A + 1).requires_grad_(True)
    fisher_CE(A, s, keep=keep).sum().backward()
    leaf = A.clone().requires_grad_(True)
    batch_torch_A_to_R(leaf).sum().backward()
    l2 = A.clone().requires_grad_(True)
    vmf_loss(l2, rot(n), keep=keep)[0].sum().backward()
    validation_terms(A, batch_torch_A_to_R(A), rot(n), -4.0)
for bl, bu in ((32, 128), (5, 0), (64, 257)):
    a = (10 * torch.randn(bl, 9, device=dev, generator=g)).requires_grad_(True)
    if bu:
        w = 10 * torch.randn(bu, 9, device=dev, generator=g)
        st = (w + 0.5).requires_grad_(True)
        for unsup in ("ce", "nll"):
            ssl_loss(a, rot(bl), w, st, -4.0, type_unsuper=unsup, aug_rot_mat=rot(bu))[0].backward()
    else:
        ssl_loss(a, rot(bl))[0].backward()
# K2L: CTA-per-sample kernel, warp kernel, point-packed stream kernel, sample-packed stream kernel (single chunk with trailing points; three chunks),
# and the sample-packed kernel as clusters of 8 / 4 / 2 CTAs that slice the grid and merge through distributed shared memory
sms = torch.cuda.get_device_properties(dev).multi_processor_count
for n, N in ((33, 37), (33, 4608), (700, 37), (sms * 256 + 5, 37), (sms * 1024 - 3, 39), (sms * 1024 - 3, 7451), (8191, 4608), (8191, 13826), (33001, 130), (90001, 4607)):
    A = 5 * torch.randn(n, 9, device=dev, generator=g)
    A[:4] = 0
    grid = rot(N)
    _ops.laplace_nll(A, rot(n), grid, grad=True, mode=True)
    _ops.laplace_nll(A, rot(n), grid, grad=False)
torch.cuda.synchronize()
print("sanitize driver done")
