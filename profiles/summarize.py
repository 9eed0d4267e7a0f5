"""Turn the ncu artefacts a gpurun call left under gpurun_out/<tag>/ into the small text
summaries committed under profiles/ (the .ncu-rep files themselves stay in gpurun_out/).

    python profiles/summarize.py <tag> [<out-prefix>]

Reads   gpurun_out/<tag>/launches.csv          (ncu --metrics gpu__time_duration.sum launch list)
        gpurun_out/<tag>/*_full.ncu-rep        (ncu --set full captures)
        gpurun_out/<tag>/bench.json, bench_ref.json
Writes  profiles/<prefix>_launches.md, profiles/<prefix>_<rep>_ncu.md, profiles/<prefix>_bench.json
"""
import collections
import csv
import glob
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEEP = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
]


def launches(tag, prefix):
    path = os.path.join(ROOT, "gpurun_out", tag, "launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, bi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Block Size"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0][-70:]
        a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    total = sum(a[1] for a in agg.values())
    out = [f"# ncu launch list `{tag}` (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: shares only)",
           "", "command: `python bench.py --gpus 1 --steps 2 --warmup 3 --skip-extra --no-cpu --no-e2e` "
           "(includes input synthesis by torch and the FP32 probe of the roofline denominator)", "",
           "| kernel | launches | total us | share | grid | block |", "|---|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| `{k}` | {a[0]} | {a[1] / 1e3:.1f} | {100 * a[1] / total:.2f}% | {a[2]} | {a[3]} |")
    ours = {k: a for k, a in agg.items() if "suhpe" in k}
    step = sum(a[1] for k, a in ours.items())
    out += ["", "Share of the hot-path step (our kernels only):", ""]
    for k, a in sorted(ours.items(), key=lambda x: -x[1][1]):
        out.append(f"- `{k}`: {100 * a[1] / step:.2f}% ({a[1] / a[0] / 1e3:.1f} us per launch)")
    open(os.path.join(ROOT, "profiles", f"{prefix}_launches.md"), "w").write("\n".join(out) + "\n")


def reps(tag, prefix):
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", tag, "*.ncu-rep"))):
        name = os.path.basename(rep)[:-8]
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        out = [f"# ncu --set full `{name}` ({tag}); per launch, --clock-control none", ""]
        seen = set()
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            kn = d.get("Kernel Name", "?")
            key = (kn, d.get("Grid Size"))
            if key in seen:
                continue
            seen.add(key)
            out += [f"## `{kn}` grid {d.get('Grid Size')} block {d.get('Block Size')}", "", "| metric | value | unit |", "|---|---|---|"]
            for m in KEEP:
                if m in d:
                    out.append(f"| {m} | {d[m]} | {units[hdr.index(m)]} |")
            out.append("")
        open(os.path.join(ROOT, "profiles", f"{prefix}_{name}_ncu.md"), "w").write("\n".join(out) + "\n")


def bench(tag, prefix):
    for f in ("bench.json", "bench_ref.json"):
        p = os.path.join(ROOT, "gpurun_out", tag, f)
        if os.path.exists(p):
            lines = [l for l in open(p).read().splitlines() if l.startswith("{")]
            if lines:
                json.dump(json.loads(lines[-1]), open(os.path.join(ROOT, "profiles", f"{prefix}_{f}"), "w"), indent=1)


if __name__ == "__main__":
    tag = sys.argv[1]
    prefix = sys.argv[2] if len(sys.argv) > 2 else tag
    launches(tag, prefix)
    reps(tag, prefix)
    bench(tag, prefix)
