"""Host-link probe: pinned H2D alone, D2H alone, both at once (two streams) -- context for bench.py's e2e leg."""
import torch
dev = torch.device("cuda:0")
n = 1 << 23
h_in = torch.empty(n, 18).pin_memory(); h_in.normal_()
h_out = torch.empty(n, 11).pin_memory()
d_in = torch.empty(n, 18, device=dev); d_out = torch.randn(n, 11, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=5):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for _ in range(2): run(True, True, 1)
t = run(True, False); print(f"H2D alone   {h_in.numel()*4/t/1e6:7.1f} GB/s  ({t:.2f} ms for {h_in.numel()*4/1e6:.0f} MB)")
t = run(False, True); print(f"D2H alone   {h_out.numel()*4/t/1e6:7.1f} GB/s  ({t:.2f} ms for {h_out.numel()*4/1e6:.0f} MB)")
t = run(True, True);  print(f"both        H2D {h_in.numel()*4/t/1e6:7.1f} GB/s + D2H {h_out.numel()*4/t/1e6:7.1f} GB/s  ({t:.2f} ms)")
