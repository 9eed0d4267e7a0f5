"""Split a kernel's stall samples by how often its SASS lines execute (ncu --set full --import-source on capture):

    ncu -i gpurun_out/<tag>/fisher_full.ncu-rep --page source --csv --print-source sass > /tmp/src.csv
    python profiles/time_by_frequency.py /tmp/src.csv <units>        # units = rotations in the launch

Classes for K2: lines that run once per pass / run, once per rotation, once per 32-rotation tile, rarely; plus the
share of instructions issued with fewer than 26 active threads (divergent regions) and every S2R in per-rotation code."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_s = sum(int(r[ix["# Samples"]]) for r in data)
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)


def klass(e):
    per = e / units
    if per >= 1.5: return "per pass / run (>= 1.5 per unit)"
    if per >= 0.5: return "per unit (0.5 - 1.5)"
    if per >= 0.02: return "per tile (0.02 - 0.5)"
    return "rare"


agg, cnt, inst = collections.Counter(), collections.Counter(), collections.Counter()
low_i = low_s = 0
for r in data:
    e = int(r[ix["Instructions Executed"]])
    k = klass(e)
    agg[k] += int(r[ix["# Samples"]]); cnt[k] += 1; inst[k] += e
    if e and float(r[ix["Avg. Threads Executed"]]) < 26:
        low_i += e; low_s += int(r[ix["# Samples"]])
print(f"{tot_i / units:.0f} warp instructions per unit, {tot_s} stall samples")
for k in ("per pass / run (>= 1.5 per unit)", "per unit (0.5 - 1.5)", "per tile (0.02 - 0.5)", "rare"):
    print(f"{k:34s} {cnt[k]:5d} SASS lines  {inst[k] / units:7.1f} instr/unit  {100 * agg[k] / tot_s:5.1f} % of samples")
print(f"issued with < 26 active threads: {100 * low_i / tot_i:.1f} % of instructions, {100 * low_s / tot_s:.1f} % of samples")
for i, r in enumerate(data):
    e = int(r[ix["Instructions Executed"]])
    op = r[ix["Source"]].strip().split()
    op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "")
    if op in ("S2R", "S2UR") and e / units >= 0.5:
        print(f"S2R in per-unit code: line {i}, {e / units:.2f} per unit, {r[ix['# Samples']]} samples: {r[ix['Source']].strip()}")
