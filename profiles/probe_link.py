"""Host-link probe: what bounds bench.py's end-to-end leg?

    python profiles/probe_link.py                      one GPU: pinned H2D alone, D2H alone, both at once
    python profiles/probe_link.py --ranks 1,2,4,8      N processes (one per GPU) doing ONLY the pinned H2D + D2H
                                                       copies of one end-to-end step (604 MB in, 377 MB out per
                                                       2^23 rotations), all at once: the aggregate GB/s is the ceiling
                                                       the N-GPU end-to-end number can reach on this host
    ... --pin                                          each rank first restricts itself to its own slice of the host
                                                       cores (os.sched_setaffinity) and allocates its pinned buffers
                                                       afterwards (first-touch NUMA placement follows the cores)

No kernel of the library runs here: this measures the box (PCIe tree, host memory, IOMMU), not the product.
"""
import argparse
import json
import os
import sys
import time


def single():
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


def rank_cores(rank, world):
    """This rank's slice of the cores the process may run on (contiguous, disjoint, at least one core)."""
    cores = sorted(os.sched_getaffinity(0))
    per = max(len(cores) // world, 1)
    mine = cores[(rank * per) % len(cores):(rank * per) % len(cores) + per]
    return mine or cores


def _worker(rank, world, pin, n, reps, barrier, out):
    if pin:
        os.sched_setaffinity(0, rank_cores(rank, world))
    import torch
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    # the bytes of one end-to-end step: A, R_gt in (72 B / rotation); nll, grad, entropy, mask out (45 B / rotation)
    h_in = torch.empty(n, 18).pin_memory(); h_in.normal_()
    h_out = torch.empty(n * 45, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, 18, device=dev)
    d_out = torch.zeros(n * 45, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    chunk = 1 << 19                                     # the pipeline's chunking: many medium copies, not one huge one

    def step():
        for c0 in range(0, n, chunk):
            c1 = min(c0 + chunk, n)
            with torch.cuda.stream(s1):
                d_in[c0:c1].copy_(h_in[c0:c1], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[c0 * 45:c1 * 45].copy_(d_out[c0 * 45:c1 * 45], non_blocking=True)

    res = {}
    for mode in ("both", "h2d", "d2h"):
        def run_mode():
            if mode == "both":
                step()
            elif mode == "h2d":
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            else:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        run_mode(); torch.cuda.synchronize()
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(reps):
            run_mode()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        barrier.wait()
        res[mode] = dt
    out[rank] = dict(res, cores=sorted(os.sched_getaffinity(0)))


def multi(ranks, pin, n, reps):
    import torch
    import torch.multiprocessing as mp
    have = torch.cuda.device_count()
    rows = []
    for world in ranks:
        if world > have:
            print(f"ranks={world}: only {have} GPU(s) visible, skipped")
            continue
        ctx = mp.get_context("spawn")
        barrier, out = ctx.Barrier(world), ctx.Manager().dict()
        procs = [ctx.Process(target=_worker, args=(r, world, pin, n, reps, barrier, out)) for r in range(world)]
        for p in procs: p.start()
        for p in procs: p.join()
        if len(out) != world:
            print(f"ranks={world}: a worker failed")
            continue
        row = {"ranks": world, "pinned_to_cores": bool(pin), "rotations_per_rank": n}
        for mode, bytes_in, bytes_out in (("both", 72, 45), ("h2d", 72, 0), ("d2h", 0, 45)):
            worst = max(out[r][mode] for r in range(world))
            row[mode] = {"ms": 1e3 * worst, "aggregate_gbs": world * n * (bytes_in + bytes_out) / worst / 1e9,
                         "per_gpu_gbs": n * (bytes_in + bytes_out) / worst / 1e9}
        row["e2e_ceiling_rot_per_s"] = world * n / max(out[r]["both"] for r in range(world))
        row["cores_rank0"] = out[0]["cores"]
        rows.append(row)
        print(json.dumps(row))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ranks", default="")
    ap.add_argument("--pin", action="store_true")
    ap.add_argument("--n", type=int, default=1 << 23)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    if not args.ranks:
        single()
    else:
        multi([int(r) for r in args.ranks.split(",")], args.pin, args.n, args.reps)


if __name__ == "__main__":
    main()
