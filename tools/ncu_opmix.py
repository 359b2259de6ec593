#!/usr/bin/env python3
"""tools/ncu_opmix.py <report.ncu-rep> <kernel-regex> -- executed-instruction mix per SASS opcode and the
stall samples per opcode, from the ncu source page (needs -lineinfo / --import-source)."""
import csv, io, subprocess, sys, collections, re
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
mix, samp = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iE: continue
    op = r[iS].strip().split()
    if not op: continue
    name = op[1] if op[0].startswith("@") else op[0]
    name = re.sub(r"\.(U32|X|E|64|128|LUT|AND|OR|NE|GE|EQ|LT|GT|NOINC|NODEC|REL|U|CONSTANT|SYS).*", lambda m: m.group(0) if False else "." + m.group(1), name)
    if not r[iE].isdigit(): continue
    n = int(r[iE]); tot += n
    mix[name] += n; samp[name] += int(r[iN] or 0)
ts = sum(samp.values())
print(f"kernel {kern}: {tot} warp instructions executed")
for k, v in mix.most_common(24):
    print(f"  {k:22s} {v:14d} {100.0*v/tot:6.2f}% inst   {100.0*samp[k]/max(ts,1):6.2f}% stall samples")
