#!/usr/bin/env python3
"""tools/sass_regions.py LIB KERNEL_SUBSTR [min_instrs] -- static opcode mix of every device function inside one kernel.

cuobjdump lists a kernel together with the noinline device functions it calls; the functions are separated at their RET
instructions.  Printed per region: start offset, instruction count, code bytes and the opcode histogram -- enough to
see before a GPU run whether a multiplier body spills predicates (LOP3 / P2R), moves registers (IMAD.MOV / MOV / SEL) or
touches local memory (LDL / STL), and how large the hot code is against the instruction caches.
"""
import re
import subprocess
import sys
from collections import Counter


def regions(lib, kernel):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(txt) if "Function :" in l and kernel in l)
    ins = []
    for l in txt[start + 1:]:
        if "Function :" in l:
            break
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    out, cur = [], []
    for a, t in ins:
        cur.append((a, t))
        if re.match(r"(@!?U?P\d+\s+)?(RET|EXIT)", t) and not t.startswith("@"):
            out.append(cur)
            cur = []
    if cur:
        out.append(cur)
    return out


def opcode(t):
    parts = t.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    f = op.split(".")
    return f[0] + (".WIDE" if "WIDE" in f else "") + (".MOV" if "MOV" in f[1:] else "") + (".128" if "128" in f else "")


if __name__ == "__main__":
    lib, kernel = sys.argv[1], sys.argv[2]
    min_n = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    regs = regions(lib, kernel)
    total = sum(len(r) for r in regs)
    print("%s: %d instructions, %.1f KB" % (kernel, total, total * 16 / 1024))
    for r in regs:
        if len(r) < min_n:
            continue
        c = Counter(opcode(t) for _, t in r)
        print("  0x%05x n=%5d %5.1f KB | %s" % (r[0][0], len(r), len(r) * 16 / 1024,
                                                " ".join("%s=%d" % kv for kv in c.most_common(10))))
