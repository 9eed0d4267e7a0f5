"""Static look at the SASS of a kernel in the built library (no GPU needed):

    python profiles/sass_loops.py fisher_fused_kernelILi3      # loops + opcode histogram
    python profiles/sass_loops.py fisher_fused_kernelILi3 0x7570 0x7960   # print that address range

For every backward branch it prints the loop's address range, its instruction count and opcode mix,
and a "chain interleave" figure for the packed FMA ops: the average distance (in FMA-pipe
instructions) between an FFMA2 and the FFMA2 that produced its accumulator operand -- 1 means a
chain runs alone (every op waits out the full dependent-issue latency), >= 3 covers it.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "semiuhpe_b200", "_lib", "libsemiuhpe_b200.so")


def kernel_sass(pattern, lib=LIB):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    ins, on = [], False
    for line in out.splitlines():
        if "Function :" in line:
            on = pattern in line
            name = line.split(":", 1)[1].strip()
            if on:
                ins = []
                found = name
            continue
        if on:
            m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        if on and ".section" in line:
            break
    return ins


def opcode(text):
    parts = text.split()
    return parts[1] if parts[0].startswith("@") else parts[0]


def interleave(body):
    """mean distance (in FMA-pipe ops) from a packed op to the producer of its first source that is a packed result"""
    fma = [(i, t) for i, (_, t) in enumerate(body) if opcode(t) in ("FFMA2", "FMUL2", "FADD2")]
    last_writer, dists = {}, []
    for pos, (_, t) in enumerate(fma):
        regs = re.findall(r"\bR(\d+)", t)
        if not regs:
            continue
        dst, srcs = regs[0], regs[1:]
        d = [pos - last_writer[s] for s in srcs if s in last_writer]
        if d:
            dists.append(min(d))
        last_writer[dst] = pos
    return sum(dists) / len(dists) if dists else 0.0, collections.Counter(min(x, 6) for x in dists)


def main():
    pat = sys.argv[1]
    ins = kernel_sass(pat)
    if not ins:
        raise SystemExit(f"no kernel matching {pat!r} in {LIB}")
    if len(sys.argv) >= 4:
        lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
        for a, t in ins:
            if lo <= a <= hi:
                print(f"{a:05x}  {t[:110]}")
        return
    print(f"{len(ins)} instructions")
    total = collections.Counter(opcode(t) for _, t in ins)
    print("whole kernel:", dict(total.most_common(14)))
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= a:
            continue
        body = [x for x in ins if tgt <= x[0] <= a]
        c = collections.Counter(opcode(t2) for _, t2 in body)
        packed = c["FFMA2"] + c["FMUL2"] + c["FADD2"]
        if packed < 8:
            continue
        mean, hist = interleave(body)
        print(f"loop {tgt:#06x}-{a:#06x}: {len(body):4d} instr, packed {packed}, MUFU {sum(v for k, v in c.items() if k.startswith('MUFU'))}, "
              f"LDS {sum(v for k, v in c.items() if k.startswith('LDS'))}, interleave {mean:.2f} {dict(sorted(hist.items()))}")


if __name__ == "__main__":
    main()
