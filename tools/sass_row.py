#!/usr/bin/env python3
"""tools/sass_row.py LIB KERNEL_SUBSTR [rows] -- one multiplier row as ptxas emitted it: the first run of carry-chained wide
MACs (IMAD.WIDE.U32[.X]) inside the kernel's code, with the scheduling control fields decoded from the upper 64-bit word
of each 128-bit instruction (Volta+ encoding: stall count = bits 105-108, yield = 109, write / read barrier = 110-112 /
113-115, wait mask = 116-121, operand reuse = 122-125).  Evidence for DESIGN.md's "one IMAD.WIDE per 4 clocks" statement."""
import re
import subprocess
import sys

lib, kernel = sys.argv[1], sys.argv[2]
want = int(sys.argv[3]) if len(sys.argv) > 3 else 26
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(txt) if "Function :" in l and kernel in l)
ins = []
i = start + 1
while i < len(txt) and "Function :" not in txt[i]:
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", txt[i])
    if m and i + 1 < len(txt):
        m2 = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", txt[i + 1])
        if m2:
            ins.append((int(m.group(1), 16), m.group(2).strip(), int(m2.group(1), 16)))
            i += 1
    i += 1
best = None
run = 0
for k, (a, t, hi) in enumerate(ins):
    run = run + 1 if "IMAD.WIDE.U32" in t else 0
    if run >= 12:
        best = k - run + 1
        break
if best is None:
    sys.exit("no run of wide MACs found")
print(f"{lib}: {kernel}: first run of carry-chained wide MACs (offset 0x{ins[best][0]:x})")
print("  offset  stall yield wbar rbar wait reuse  instruction")
for a, t, hi in ins[max(best - 2, 0):best + want]:
    stall, yld, wbar, rbar, wait, reuse = (hi >> 41) & 0xF, (hi >> 45) & 1, (hi >> 46) & 7, (hi >> 49) & 7, (hi >> 52) & 0x3F, (hi >> 58) & 0xF
    print(f"  0x{a:05x}  {stall:5d} {yld:5d} {wbar if wbar != 7 else '-':>4} {rbar if rbar != 7 else '-':>4} {wait:04x} {reuse:5x}  {t}")
wide = [(hi >> 41) & 0xF for a, t, hi in ins if "IMAD.WIDE.U32" in t]
hist = {s: wide.count(s) for s in sorted(set(wide))}
print(f"stall-count histogram over all {len(wide)} IMAD.WIDE.U32 of the kernel's code: {hist}")
