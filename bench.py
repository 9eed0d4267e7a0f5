#!/usr/bin/env python
"""bench.py -- rotations/s of the fused Fisher NLL + gradient + entropy + percentile filter.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json metric, config 5 sharded): every GPU owns 2^23 synthetic (A, R_gt)
pairs (2^26 at 8 GPUs).  One step = one pass of the hot path over that pool:
  K2  fused matrix-Fisher kernel -> NLL, d NLL/dA, entropy, first radix histogram
  K3  global percentile threshold (k = int(n_total * 0.95)): 2 more histogram passes,
      3 scans; on N > 1 GPUs the 2048-bin histograms are all-gathered over NCCL
  K3  keep-mask emit (entropy < threshold)
Inputs are resident in HBM for `value`; `e2e` runs the same step through the C ABI's
host-buffer entry (pinned host inputs, H2D and D2H copies inside the timed region).
`--impl reference` times the reference's CPU algorithm (oracle port: torch CPU, all host
threads) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rotations/sec fused Fisher NLL+grad+entropy+filter"
UNIT = "rotations/s"
N_PER_GPU = 1 << 23
LEFT_RATIO = 0.95
OVERREG = 1.025
FLOP_PER_ROTATION = 69120          # 3 families x 512 nodes x 45 FLOP
HBM_BYTES_PER_ROTATION = 36 + 36 + 4 + 36 + 4   # A, R in; nll, grad, entropy out (K2)
CPU_SAMPLE = 1 << 16


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = str(gpu_index)
        self.proc, self.rows = None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", self.gpu, f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------ CPU baseline
def cpu_step(orc, torch, np, A, R):
    """The reference's CPU algorithm for one step on a sample: vmf_loss fwd+bwd (agent.py:79,209),
    fisher_entropy (agent.py:139), sort/index threshold + strict-< mask (agent.py:403-407,148)."""
    leaf = A.clone().requires_grad_(True)
    loss, _ = orc.vmf_loss(leaf, R, overreg=OVERREG)
    loss.mean().backward()
    with torch.no_grad():
        ent = orc.fisher_entropy(A)
    thr, _ = orc.pool_threshold(ent.numpy(), LEFT_RATIO)
    mask, _ = orc.keep_mask(ent, float(thr))
    return float(loss.detach().mean()), int(mask.sum())


def synth_cpu(torch, n, seed):
    gen = torch.Generator().manual_seed(seed)
    A = 10 * torch.randn(n, 9, generator=gen)
    q = torch.randn(n, 4, generator=gen)
    q = q / q.norm(dim=1, keepdim=True)
    from oracle import pytorch3d_restated as p3d
    return A, p3d.quaternion_to_matrix(q).contiguous()


def pick_threads(orc, torch, np):
    """The reference arm gets the thread count it runs fastest with on this host (on
    oversubscribed VMs torch's OpenMP pool can be slower at cpu_count() than at 2)."""
    cores = os.cpu_count() or 1
    A, R = synth_cpu(torch, 512, 99)
    best_t, best_n = None, cores
    n = cores
    while n >= 1:
        torch.set_num_threads(n)
        cpu_step(orc, torch, np, A[:64], R[:64])
        t0 = time.perf_counter()
        cpu_step(orc, torch, np, A, R)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best_t, best_n = dt, n
        n //= 2
    torch.set_num_threads(best_n)
    return best_n


def time_cpu(sample, repeats):
    import numpy as np
    import torch
    from oracle import so3_oracle as orc
    cores = pick_threads(orc, torch, np)
    A, R = synth_cpu(torch, sample, 1234)
    cpu_step(orc, torch, np, A[:256], R[:256])            # warm-up
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        cpu_step(orc, torch, np, A, R)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return sample / best, cores, best


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import torch
    from oracle import so3_oracle as orc
    cores = pick_threads(orc, torch, np)
    sample = max(args.cpu_sample // 4, 1024)          # per step; K steps stay within a few minutes
    A, R = synth_cpu(torch, sample, 1234)
    for _ in range(max(args.warmup, 1)):
        cpu_step(orc, torch, np, A[:2048], R[:2048])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(orc, torch, np, A, R)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, args.n_per_gpu),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "host_cpus": os.cpu_count(), "kind": "port",
                         "sample": f"{sample} of the pool's rotations per step: oracle/so3_oracle.py (torch CPU restatement "
                                   "of vmf_loss fwd+bwd, fisher_entropy, numpy sort threshold, mask), all host threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(n_gpus, n_per_gpu):
    return {"workload": f"BASELINE config 5 sharded: {n_per_gpu} synthetic (A=10*randn, R_gt random rotation) pairs per GPU "
                        f"({n_per_gpu * n_gpus} total), Fisher NLL+grad+entropy (overreg={OVERREG}) + global percentile "
                        f"filter left_ratio={LEFT_RATIO}",
            "rotations_per_gpu": n_per_gpu, "rotations_total": n_per_gpu * n_gpus, "left_ratio": LEFT_RATIO,
            "parallelism": f"batch-sharded x{n_gpus}, all-gather of 2048-bin radix histograms",
            "l2": "inputs (604 MB per GPU) exceed the 126 MB L2; no flush needed"}


# ------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as graft
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    cores = None
    if world > 1 and not args.no_pin:
        # one disjoint slice of the host cores per rank, set BEFORE any pinned buffer is allocated (first-touch
        # placement follows the cores): the ranks' copy threads and Python threads stop migrating over each other
        avail = sorted(os.sched_getaffinity(0))
        per = max(len(avail) // world, 1)
        cores = avail[(local * per) % len(avail):(local * per) % len(avail) + per] or avail
        os.sched_setaffinity(0, cores)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        graft.build()
    if world > 1:
        dist.barrier()
    from semiuhpe_b200 import _capi, _ops
    import semiuhpe_b200
    from semiuhpe_b200.agent import pool_index, _quat_to_matrix
    semiuhpe_b200.set_error_checking(False)          # sync-free steps; finiteness is asserted after the run
    lib = _capi.lib()

    n = args.n_per_gpu
    n_total = n * world
    k = pool_index(n_total, LEFT_RATIO)
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    A = 10 * torch.randn(n, 9, device=dev, generator=gen)
    q = torch.randn(n, 4, device=dev, generator=gen)
    R = _quat_to_matrix(q / q.norm(dim=1, keepdim=True)).reshape(n, 9).contiguous()
    del q
    new = lambda *s, dt=torch.float32: torch.empty(s, dtype=dt, device=dev)
    nll, grad, ent, mask = new(n), new(n, 9), new(n), new(n, dt=torch.bool)
    ws = _ops.SelectWorkspace(dev)
    kept = torch.zeros(1, dtype=torch.int64, device=dev)
    gathered = torch.empty((world, _capi.HIST_BINS), dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    P, S = _capi.ptr, _capi.stream
    launches = [0]

    def fused(bits=_capi.CUT_BITS_DEFAULT):
        # K2 counts the first radix digit of every entropy itself (hist): one launch
        _capi.check(lib.suhpe_fisher_fused_f32(P(A), P(R), n, OVERREG, bits, P(nll), P(grad), None, P(ent), None, None, None,
                                               P(ws.hist[0]), P(status), S()), "fused")
        launches[0] += 1

    def step():
        ws.hist.zero_()
        kept.zero_()
        fused()
        _capi.check(lib.suhpe_select_init(P(ws.state), k, S()), "init"); launches[0] += 1
        for p in (1, 2, 3):
            if p == 1:
                local_hist = ws.hist[0]
            else:
                local_hist = ws.hist[1]
                _capi.check(lib.suhpe_select_hist_f32(P(ent), n, p, P(ws.state), P(local_hist), S()), "hist")
                launches[0] += 1
            if world > 1:
                dist.all_gather_into_tensor(gathered, local_hist.reshape(1, -1))
                src, parts = gathered, world
            else:
                src, parts = local_hist, 1
            _capi.check(lib.suhpe_select_scan(P(src), parts, p, P(ws.state), S()), "scan"); launches[0] += 1
        _capi.check(lib.suhpe_entropy_mask_f32(P(ent), n, ws.threshold_ptr(), 0.0, P(mask), P(kept), S()), "mask")
        launches[0] += 1

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    step_launches = launches[0]

    # dominant kernel alone (same stream, CUDA events): roofline numerator
    for _ in range(3):
        fused()
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(args.steps, 5)
    k0.record()
    for _ in range(reps):
        fused()
    k1.record()
    torch.cuda.synchronize()
    fused_ms = k0.elapsed_time(k1) / reps
    # the same kernel with the negligible-node cut disabled (all 3 x 512 nodes of every rotation evaluated)
    prev_bits = _capi.CUT_BITS_DEFAULT
    for _ in range(2):
        fused(0)
    torch.cuda.synchronize()
    k0.record()
    for _ in range(5):
        fused(0)
    k1.record()
    torch.cuda.synchronize()
    fused_all_ms = k0.elapsed_time(k1) / 5
    clocks = sampler.stop() if sampler else None
    ws.hist.zero_()
    kept.zero_()
    step()                                               # leave the outputs of one default-setting step in place
    torch.cuda.synchronize()

    thr = ws.read()[0]
    kept_n = int(kept.item())
    assert int(status.item()) == 0 and bool(torch.isfinite(nll).all()) and bool(torch.isfinite(ent).all())
    value = n_total * args.steps / (ms_total * 1e-3)
    # parity of the (all-gathered) radix select against a sort of the concatenated pool (src/agent.py:403-407)
    check_dev = threshold_check(torch, dist, dev, world, rank, ent, thr, kept_n, k)

    # end-to-end leg on every rank at once (they share the host's PCIe/memory system), max over ranks
    e2e = None if args.no_e2e else run_e2e(torch, dist, dev, n, args, world, rank, k, cores)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- rank 0 extras -------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    probe = fp32_probe(torch, lib, dev, _capi)
    achieved_tflops = n * FLOP_PER_ROTATION / (fused_ms * 1e-3) / 1e12
    peak_tflops = max(probe.values())
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_gbs = n * HBM_BYTES_PER_ROTATION / (fused_ms * 1e-3) / 1e9
    roofline = {
        "bound": "fp32", "kernel": "fisher_fused_kernel", "achieved": achieved_tflops, "peak": peak_tflops,
        "unit": "TFLOP/s", "frac": achieved_tflops / peak_tflops,
        "frac_all_nodes_evaluated": n * FLOP_PER_ROTATION / (fused_all_ms * 1e-3) / 1e12 / peak_tflops,
        "peak_source": "FP32 FMA micro-benchmark run on this GPU in this process (suhpe_fp32_probe; MEASURED_PEAKS.json "
                       "has no FP32-pipe figure); nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
        "algorithmic_flop_per_rotation": FLOP_PER_ROTATION, "rotations_per_launch": n, "kernel_ms": fused_ms,
        "kernel_share_of_step": fused_ms / (ms_total / args.steps),
        "traffic": (k2_traffic(n) or {}).get("bytes"), "traffic_detail": k2_traffic(n),
        "negligible_node_cut": {
            "bits": prev_bits,
            "note": "nodes whose whole prefix is provably below 2^-bits of the normaliser sum are skipped (DESIGN.md K2; "
                    "float64 proof test in tests/test_emul_math.py); `achieved` counts the algorithmic 69,120 FLOP per "
                    "rotation either way",
            "all_nodes_evaluated": {"kernel_ms": fused_all_ms,
                                    "achieved": n * FLOP_PER_ROTATION / (fused_all_ms * 1e-3) / 1e12,
                                    "frac": n * FLOP_PER_ROTATION / (fused_all_ms * 1e-3) / 1e12 / peak_tflops}},
        "hbm_view": {"achieved_gbs": hbm_gbs, "peak_gbs": hbm_peak, "frac": hbm_gbs / hbm_peak,
                     "bytes_per_rotation": HBM_BYTES_PER_ROTATION,
                     "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback"},
        "probe": probe,
    }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(world, n),
        "clocks": clocks, "e2e": e2e, "gpu_launches": step_launches, "roofline": roofline,
        "threshold": thr, "kept": kept_n,
        "threshold_check": {"device": check_dev, "e2e": (e2e or {}).pop("threshold_check", None)},
    }
    # GPU side numbers first, while the device is warm: the CPU legs below leave it idle for tens of seconds and a
    # 20 us launch timed on a GPU that has dropped to its idle clocks reads 3-4x slow
    if world == 1 and not args.skip_extra:
        line["extra"] = side_configs(torch, dev, _ops, peak_tflops, hbm_peak, with_cpu=not args.no_cpu, warm=fused)
    if world == 1 and not args.no_cpu:
        v, cores, secs = time_cpu(args.cpu_sample, 1)
        line["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{args.cpu_sample} rotations of the same workload, 1 pass ({secs:.1f} s): oracle/so3_oracle.py "
                      "(torch-CPU restatement of the reference: vmf_loss fwd+bwd, fisher_entropy, numpy sort + mask), "
                      "all host threads"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    emit(line)


def k2_traffic(n):
    """DRAM bytes of one K2 launch from the committed ncu capture (never measured in the timed run);
    None when the capture was taken at another launch size."""
    for name in ("r02_k2_traffic.json", "r01_k2_traffic.json"):          # newest capture first
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        t = json.load(open(path))
        if t.get("rotations_per_launch") != n:
            continue
        return {"bytes": t["dram_bytes_read"] + t["dram_bytes_write"], "algorithmic_bytes": t["algorithmic_bytes"],
                "source": f"profiles/{name} (ncu dram__bytes_read.sum + dram__bytes_write.sum, one launch)"}
    return None


def fp32_probe(torch, lib, dev, _capi):
    """FP32-pipe peak on this GPU: 8 independent FMA chains per thread, 148*8 CTAs x 256 threads."""
    sink = torch.zeros(4, device=dev)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    blocks, iters = sms * 8, 4096
    out = {}
    for variant, name, per in ((0, "ffma", 1), (1, "ffma2", 2), (2, "ffma_mufu", 1), (3, "ffma2_ffma_mixed", 3),
                               (4, "ffma2_plus_lop3", 2), (5, "ffma2_plus_iadd", 2),
                               (6, "ffma2_imm_addend", 2), (7, "ffma2_bcast_scalar", 2),
                               (8, "ffma2_three_distinct_regs", 2), (9, "ffma2_two_distinct_regs_imm", 2),
                               (10, "ffma2_plus_fsel", 2), (11, "ffma2_plus_fmnmx", 2),
                               (12, "ffma2_plus_lds128_per8", 2), (13, "ffma2_plus_mufu_per8", 2)):
        for _ in range(2):
            _capi.check(lib.suhpe_fp32_probe(_capi.ptr(sink), variant, iters, blocks, _capi.stream()), "probe")
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            _capi.check(lib.suhpe_fp32_probe(_capi.ptr(sink), variant, iters, blocks, _capi.stream()), "probe")
        b.record()
        torch.cuda.synchronize()
        fmas = 3.0 * blocks * 256 * iters * 64 * per
        out[name + "_tflops"] = 2 * fmas / (a.elapsed_time(b) * 1e-3) / 1e12
    return out


def threshold_check(torch, dist, dev, world, rank, ent, thr, kept_local, k):
    """Driver-visible parity of the threshold: every rank must hold the bit-identical value, and it must equal
    ``sort(concat(all shards))[k]`` with ``sum(kept) == count(e < thr)`` (src/agent.py:403-407,148).  The
    entropies of every rank are all-gathered (268 MB at 8 x 2^23) and sorted with torch.sort on the device."""
    import struct
    bits = struct.unpack("<i", struct.pack("<f", thr))[0]
    mine = torch.tensor([bits, kept_local], dtype=torch.int64, device=dev)
    if world > 1:
        every = torch.empty((world, 2), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(every, mine.reshape(1, 2))
        pool = torch.empty((world, ent.numel()), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(pool, ent.reshape(1, -1).contiguous())
    else:
        every, pool = mine.reshape(1, 2), ent.reshape(1, -1)
    same = bool((every[:, 0] == every[0, 0]).all())
    kept_sum = int(every[:, 1].sum())
    out = {"ranks": world, "n_total": int(pool.numel()), "k": int(k), "identical_on_all_ranks": same,
           "kept_sum": kept_sum}
    if rank == 0:
        flat = pool.reshape(-1)
        ref = torch.sort(flat).values[k]
        below = int((flat < ref).sum())
        out.update(sorted_k=float(ref), equals_sorted_k=bool(ref.view(torch.int32) == bits),
                   kept_equals_count_below=bool(kept_sum == below), count_below=below)
        assert same and out["equals_sorted_k"] and out["kept_equals_count_below"], out
    del pool
    return out


def run_e2e(torch, dist, dev, n, args, world, rank, k, cores=None):
    """Same step through the C ABI's host-buffer entry: pinned host A/R -> H2D -> K2/K3 -> D2H."""
    from semiuhpe_b200.host_pipeline import FisherFilterPipeline
    gen = torch.Generator().manual_seed(77 + rank)
    A_h = (10 * torch.randn(n, 9, generator=gen)).pin_memory()
    q = torch.randn(n, 4, generator=gen)
    q = q / q.norm(dim=1, keepdim=True)
    from semiuhpe_b200.agent import _quat_to_matrix
    R_h = _quat_to_matrix(q).reshape(n, 9).contiguous().pin_memory()
    pipe = FisherFilterPipeline(max_n=n, chunk=args.e2e_chunk, device=dev.index)
    group = True if world > 1 else None                    # N > 1: global threshold over all ranks' shards
    res = pipe.run(A_h, R_h, OVERREG, LEFT_RATIO, group=group)   # warm-up (allocates pinned outputs)
    pipe.run(A_h, R_h, OVERREG, LEFT_RATIO, group=group)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    steps = max(2, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(steps):
        res = pipe.run(A_h, R_h, OVERREG, LEFT_RATIO, group=group)   # blocking call: results are on the host on return
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    # context for the number: the host link alone (one pinned -> device copy of the step's A and R_gt,
    # nothing else running) -- the end-to-end leg cannot be faster than its H2D bytes over this rate
    link = None
    if world == 1:
        d_A, d_R = torch.empty_like(A_h, device=dev), torch.empty_like(R_h, device=dev)
        for _ in range(2):
            d_A.copy_(A_h, non_blocking=True); d_R.copy_(R_h, non_blocking=True)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            d_A.copy_(A_h, non_blocking=True); d_R.copy_(R_h, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_gbs = 3 * (A_h.numel() + R_h.numel()) * 4 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        link = {"h2d_gbs": h2d_gbs, "h2d_bound_rotations_per_s": h2d_gbs * 1e9 / 72.0,
                "note": "pinned-host -> device copy of one step's inputs (72 B per rotation) timed alone"}
        del d_A, d_R
    out = {"value": n * world * steps / dt, "unit": UNIT, "h2d_bytes_per_step": res["h2d_bytes"],
           "d2h_bytes_per_step": res["d2h_bytes"], "steps": steps, "ms_per_step": 1e3 * dt / steps,
           "api": ("semiuhpe_b200.host_pipeline.FisherFilterPipeline.run -> "
                   + ("suhpe_fisher_pool_host + all-gathered radix select + mask" if world > 1 else "suhpe_fisher_filter_host")
                   + f" ({args.e2e_chunk}-pair chunks, H2D / kernel / D2H queues over 4 buffers)"),
           "threshold": res["threshold"], "kept": res["kept"],
           "host_cores_per_rank": None if cores is None else len(cores),
           "note": "all ranks concurrently (they share the host's PCIe/memory system), max over ranks; global threshold"
                   if world > 1 else "single GPU"}
    if link:
        out["host_link"] = link
    out["threshold_check"] = threshold_check(torch, dist, dev, world, rank, res["entropy"].to(dev), res["threshold"],
                                             res["kept"], k)
    pipe.close()
    return out


def side_configs(torch, dev, _ops, fp32_peak_tflops, hbm_peak_gbs, with_cpu=True, warm=None):
    """BASELINE configs 1-4 (parity-test cases, reported for context): device-timed numbers, their roofline
    fractions where a roofline applies (C3: FP32 pipe, C4: HBM), and the reference's CPU algorithm (oracle port)
    timed on this box's host cores in the same run, bounded samples."""
    import semiuhpe_b200
    from semiuhpe_b200.agent import dynamic_entropy_filter, _quat_to_matrix, ssl_loss, unsupervised_terms
    from semiuhpe_b200.fisher.fisher_utils import vmf_loss
    out = {}

    def timed(fn, reps):
        if warm is not None:                      # ~100 ms of the big kernel: SM clocks at their loaded value
            for _ in range(10):
                warm()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def graphed(fn):
        """fn captured once into a CUDA graph (warm-up on the capture stream first: handles, workspaces and
        kernel attributes are created outside the capture); returns the replay callable."""
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fn()
        return g.replay

    gen = torch.Generator(device=dev).manual_seed(5)
    rot = lambda m: _quat_to_matrix(torch.nn.functional.normalize(torch.randn(m, 4, device=dev, generator=gen), dim=1)).contiguous()
    A32, R32 = 10 * torch.randn(32, 9, device=dev, generator=gen), rot(32)
    A128 = 10 * torch.randn(128, 9, device=dev, generator=gen)
    S128 = A128 + 0.5
    leaf32 = A32.clone().requires_grad_(True)
    leaf128 = S128.clone().requires_grad_(True)

    def c1():
        leaf32.grad = None
        loss, _ = vmf_loss(leaf32, R32, overreg=OVERREG)
        loss.mean().backward()

    def c1_one_call():
        leaf32.grad = None
        ssl_loss(leaf32, R32, overreg=OVERREG)[0].backward()

    def c2():
        c1()
        dynamic_entropy_filter(A128, LEFT_RATIO, return_threshold=False)

    def c2_ssl_mirrors():
        # the whole SSL step of BASELINE config 2 through the per-function mirrors (type_unsuper 'ce'):
        # supervised NLL fwd+bwd on 32 + entropy/mask/fisher_CE fwd+bwd on 128, no host sync
        c1()
        leaf128.grad = None
        unsupervised_terms(A128, leaf128, -3.6, type_unsuper="ce")["unsuper_loss"].backward()

    def c2_ssl_one_call():
        # the same loss head as ONE C call (suhpe_ssl_step_f32) + its backward
        leaf32.grad = None
        leaf128.grad = None
        ssl_loss(leaf32, R32, A128, leaf128, -3.6, SSL_lambda=1.0, type_unsuper="ce", overreg=OVERREG)[0].backward()

    out["c1_fisher_b32_fwd_bwd_us"] = 1e3 * timed(c1, 500)
    out["c1_fisher_b32_fwd_bwd_one_call_us"] = 1e3 * timed(c1_one_call, 500)
    out["c2_teacher_step_32_128_us"] = 1e3 * timed(c2, 500)
    out["c2_ssl_step_ce_32_128_mirrors_us"] = 1e3 * timed(c2_ssl_mirrors, 300)
    out["c2_ssl_step_ce_32_128_us"] = 1e3 * timed(c2_ssl_one_call, 500)
    try:
        out["c1_fisher_b32_fwd_bwd_graph_us"] = 1e3 * timed(graphed(c1), 1000)
        out["c2_ssl_step_ce_32_128_graph_us"] = 1e3 * timed(graphed(c2_ssl_one_call), 1000)
    except Exception as exc:                               # a capture failure must not take the bench line down
        out["graph_capture_error"] = f"{type(exc).__name__}: {exc}"[:300]
        torch.cuda.synchronize()
    out["small_batch_note"] = ("wall time per step in a tight loop (CUDA events over 300-1000 steps on a warm GPU: the larger of host launch cost "
                               "and device time); *_graph_us = the same Python step captured once in a CUDA graph and replayed")
    nce = 1 << 22
    Ace1 = 10 * torch.randn(nce, 9, device=dev, generator=gen)
    Ace2 = Ace1 + 2 * torch.randn(nce, 9, device=dev, generator=gen)
    ms = timed(lambda: _ops.fisher_ce(Ace1, Ace2, grad=True), 3)
    out["fisher_ce_2p22_fwd_bwd_ms"] = ms
    out["fisher_ce_pairs_per_s"] = nce / (ms * 1e-3)
    del Ace1, Ace2
    n3, N = 1 << 20, 4608
    grid = rot(N)
    A3, R3 = 5 * torch.randn(n3, 9, device=dev, generator=gen), rot(n3)
    ms = timed(lambda: _ops.laplace_nll(A3, R3, grid, grad=True, mode=True), 3)
    out["c3_laplace_2p20_N4608_ms"] = ms
    out["c3_laplace_rot_per_s"] = n3 / (ms * 1e-3)
    out["c3_laplace_tflops_at_230400_flop"] = n3 * 230400 / (ms * 1e-3) / 1e12
    out["c3_roofline"] = {"bound": "fp32", "achieved": out["c3_laplace_tflops_at_230400_flop"], "peak": fp32_peak_tflops,
                          "unit": "TFLOP/s", "frac": out["c3_laplace_tflops_at_230400_flop"] / fp32_peak_tflops,
                          "algorithmic_flop_per_rotation": 230400}
    # the same call at other batch sizes: training-sized (CTA-per-sample kernel), mid-sized (thread-block clusters that
    # slice the grid and merge through distributed shared memory)
    for nb, reps in ((160, 200), (8192, 50), (32768, 20)):
        out[f"c3_laplace_n{nb}_N4608_us"] = 1e3 * timed(lambda: _ops.laplace_nll(A3[:nb], R3[:nb], grid, grad=True, mode=True), reps)
    n4 = 10_000_000
    Rp, Rg = rot(n4), rot(n4)
    ge = (torch.rand(n4, 3, device=dev, generator=gen) * 2 - 1) * 89
    ms = timed(lambda: _ops.so3_metrics(Rp, Rg, ge, geo=True, frob=True, abs_err=True, sums=True), 5)
    out["c4_metrics_10M_ms"] = ms
    out["c4_metrics_gbs_at_104B"] = n4 * 104 / (ms * 1e-3) / 1e9
    out["c4_roofline"] = {"bound": "hbm", "achieved": out["c4_metrics_gbs_at_104B"], "peak": hbm_peak_gbs, "unit": "GB/s",
                          "frac": out["c4_metrics_gbs_at_104B"] / hbm_peak_gbs, "algorithmic_bytes_per_pair": 104}
    if with_cpu:
        cpu = side_cpu_baselines(torch, A32.cpu(), R32.cpu(), A128.cpu(), S128.cpu(), A3[:256].cpu(), R3[:256].cpu(), grid.cpu(),
                                 Rp[:1_000_000].cpu(), Rg[:1_000_000].cpu(), ge[:20_000].cpu())
        out["cpu_baseline"] = cpu
        out["c1_speedup_vs_cpu"] = cpu["c1_fisher_b32_fwd_bwd_us"] / out["c1_fisher_b32_fwd_bwd_us"]
        out["c2_ssl_speedup_vs_cpu"] = cpu["c2_ssl_step_ce_32_128_us"] / out["c2_ssl_step_ce_32_128_us"]
        out["c3_speedup_vs_cpu"] = out["c3_laplace_rot_per_s"] / cpu["c3_laplace_rot_per_s"]
        out["c4_speedup_vs_cpu"] = (n4 / (out["c4_metrics_10M_ms"] * 1e-3)) / cpu["c4_metrics_pairs_per_s"]
    return out


def side_cpu_baselines(torch, A32, R32, A128, S128, A3, R3, grid, Rp, Rg, ge):
    """The reference's CPU algorithm (oracle/so3_oracle.py, torch CPU, pinned by golden vectors generated from
    the live reference) for BASELINE configs 1-4 on this host: C1/C2 at full size, C3 on a 256-rotation chunk
    (the reference materialises (b,N,3,3): 166 KB per rotation plus autograd copies), C4 geodesic on 10^6 pairs
    and the Euler MAE on 2*10^4 through the oracle's VECTORISED restatement (the reference itself loops over the samples
    in Python, src/utils.py:240-242: 13 us per sample at survey time, ~100x slower) -- rates, stated as such."""
    from oracle import so3_oracle as orc
    import numpy as np
    cores = torch.get_num_threads()

    def best(fn, reps):
        fn()
        t = None
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            t = dt if t is None else min(t, dt)
        return t

    def c1():
        leaf = A32.clone().requires_grad_(True)
        loss, _ = orc.vmf_loss(leaf, R32, overreg=OVERREG)
        loss.mean().backward()

    def c2_teacher():
        c1()
        with torch.no_grad():
            ent = orc.fisher_entropy(A128)
        thr, _ = orc.pool_threshold(ent.numpy(), LEFT_RATIO)
        orc.keep_mask(ent, float(thr))

    def c2_ssl():
        c1()
        with torch.no_grad():
            ent = orc.fisher_entropy(A128)
        mask, _ = orc.keep_mask(ent, -3.6)
        strong = S128.clone().requires_grad_(True)
        if int(mask.sum()) > 0:
            ce = orc.fisher_ce(A128[mask], strong[mask])
            (ce.mean() * mask.float().mean()).backward()

    def c3():
        leaf = A3.clone().requires_grad_(True)
        losses, _ = orc.laplace_nll("RLaplace", leaf, R3, grid.reshape(-1, 3, 3))
        losses.mean().backward()

    out = {"kind": "port", "cores": cores, "host_cpus": os.cpu_count()}
    out["c1_fisher_b32_fwd_bwd_us"] = 1e6 * best(c1, 5)
    out["c2_teacher_step_32_128_us"] = 1e6 * best(c2_teacher, 3)
    out["c2_ssl_step_ce_32_128_us"] = 1e6 * best(c2_ssl, 3)
    t = best(c3, 2)
    out["c3_laplace_rot_per_s"] = A3.shape[0] / t
    out["c3_sample"] = f"{A3.shape[0]} rotations x {grid.shape[0]} grid points fwd+bwd, {t:.2f} s, rate extrapolates linearly"
    t = best(lambda: orc.geodesic_deg(Rp.reshape(-1, 3, 3), Rg.reshape(-1, 3, 3)), 2)
    geo_rate = Rp.shape[0] / t
    m = ge.shape[0]
    t2 = best(lambda: orc.err_deg_from_matrices(Rp[:m].reshape(-1, 3, 3), Rg[:m].reshape(-1, 3, 3), ge), 1)
    euler_rate = m / t2
    out["c4_geodesic_pairs_per_s"] = geo_rate
    out["c4_euler_mae_pairs_per_s"] = euler_rate
    out["c4_metrics_pairs_per_s"] = 1.0 / (1.0 / geo_rate + 1.0 / euler_rate)     # both metrics per pair, as K4 computes
    out["c4_sample"] = f"geodesic on {Rp.shape[0]} pairs ({t:.2f} s), Euler MAE on {m} pairs ({t2:.2f} s)"
    return out


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly one JSON line.  Libraries write there too (NCCL prints its version banner
    on stdout whenever NCCL_DEBUG is WARN or VERSION), so the real stdout is kept aside for the result
    and file descriptor 1 is pointed at stderr for everything else in the process."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-per-gpu", type=int, default=N_PER_GPU)
    ap.add_argument("--cpu-sample", type=int, default=CPU_SAMPLE)
    ap.add_argument("--e2e-chunk", type=int, default=1 << 19, help="pairs per chunk of the host-buffer pipeline")
    ap.add_argument("--skip-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs only)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs only)")
    ap.add_argument("--no-pin", action="store_true", help="N > 1: do not restrict each rank to its own slice of the host cores")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
