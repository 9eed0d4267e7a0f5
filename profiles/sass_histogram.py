"""Writes profiles/r02_sass_histogram.md from the built library (no GPU): python profiles/sass_histogram.py"""
import collections
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_loops import LIB, interleave, kernel_sass, opcode

KERNELS = [("fisher_fused_kernelILi3", "fisher_fused_kernel<3> (K2, full head)", 4),
           ("laplace_stream2_kernelILb1", "laplace_stream2_kernel<true> (K2L stream kernel, forward+backward)", 2)]


def section(pat, title, max_loops):
    ins = kernel_sass(pat)
    out = [f"## `{title}` — {len(ins)} SASS instructions (sm_100a, `cuobjdump -sass` of the committed build)", ""]
    total = collections.Counter(opcode(t) for _, t in ins).most_common(24)
    out += ["| opcode | count | | opcode | count |", "|---|---|---|---|---|"]
    for i in range(0, len(total), 2):
        a = total[i]
        b = total[i + 1] if i + 1 < len(total) else ("", "")
        out.append(f"| {a[0]} | {a[1]} | | {b[0]} | {b[1]} |")
    out += ["", "Hot loops (innermost backward branches whose body holds packed FMA-pipe ops; `interleave` = mean distance, in packed ops, between a packed op "
            "and the producer of its operand — 1 = a chain running alone):", "",
            "| loop | instr | FFMA2 | FMUL2 | FADD2 | MUFU | LDS | FSEL/FMNMX/FSETP | other | interleave |", "|---|---|---|---|---|---|---|---|---|---|"]
    loops = []
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if not m or int(m.group(1), 16) >= a:
            continue
        tgt = int(m.group(1), 16)
        body = [x for x in ins if tgt <= x[0] <= a]
        c = collections.Counter(opcode(t2) for _, t2 in body)
        if c["FFMA2"] + c["FMUL2"] + c["FADD2"] >= 8:
            loops.append((len(body), tgt, a, body, c))
    loops.sort()
    for n, tgt, a, body, c in loops[:max_loops]:
        mufu = sum(v for k, v in c.items() if k.startswith("MUFU"))
        lds = sum(v for k, v in c.items() if k.startswith("LDS"))
        sel = sum(v for k, v in c.items() if k.split(".")[0] in ("FSEL", "FMNMX", "FMNMX3", "FSETP"))
        other = n - c["FFMA2"] - c["FMUL2"] - c["FADD2"] - mufu - lds - sel
        out.append(f"| {tgt:#06x}–{a:#06x} | {n} | {c['FFMA2']} | {c['FMUL2']} | {c['FADD2']} | {mufu} | {lds} | {sel} | {other} | {interleave(body)[0]:.2f} |")
    return out + [""]


lines = ["# Opcode histograms of the hot kernels (round 2, final tree)", "",
         f"Produced by `python profiles/sass_histogram.py` on `{os.path.relpath(LIB, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))}`; no GPU involved.",
         "K2's four 64-pair pass loops (LS, LL, SS, SL) each carry 44–46 packed ops, 4 MUFU.EX2 and 2–4 LDS.128 per 128 nodes (the round-1 build had the same counts but",
         "`interleave` 1.0–1.5: a Horner chain running alone).  K2L's loop body includes its rarely taken rescale block and clamp path; the common path of a trip",
         "(4 grid points x 2 samples) is 100 packed ops + 16 MUFU + 9 LDS.128 + 19 others.", ""]
for k in KERNELS:
    lines += section(*k)
open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "r02_sass_histogram.md"), "w").write("\n".join(lines))
print("\n".join(lines[-14:]))
