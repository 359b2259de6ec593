#!/usr/bin/env python3
"""tools/ncu_regions.py <report.ncu-rep> <kernel-regex> -- splits a kernel's SASS (ncu source page) into the
device functions it calls (regions end at RET/EXIT) and prints, per region: executed warp instructions, share of
stall samples, the dominant stall reasons and a signature (IMAD.WIDE / IADD3 / LDL / STL counts) to recognise it."""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# several captured launches are concatenated: keep the first kernel block only
hdr = rows[1]
body = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    body.append(r)
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
regions, cur = [], {"n": 0, "exec": 0, "samp": 0, "ops": collections.Counter(), "st": collections.Counter(), "first": None}
for r in body:
    if len(r) <= iE or not r[iE].isdigit():
        continue
    toks = r[iS].strip().split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    if cur["first"] is None:
        cur["first"] = r[0]
    cur["n"] += 1; cur["exec"] += int(r[iE]); cur["samp"] += int(r[iN] or 0)
    cur["ops"][op.split(".")[0] + ("." + op.split(".")[1] if op.startswith(("IMAD.WIDE", "LDL", "STL", "LDG", "STG")) and "." in op else "")] += int(r[iE])
    for i, h in stall_cols:
        cur["st"][h] += int(r[i] or 0)
    if op.startswith(("RET", "EXIT")):
        regions.append(cur)
        cur = {"n": 0, "exec": 0, "samp": 0, "ops": collections.Counter(), "st": collections.Counter(), "first": None}
if cur["n"]:
    regions.append(cur)
te, ts = sum(r["exec"] for r in regions), sum(r["samp"] for r in regions)
print(f"{kern}: {te} warp instr, {ts} samples, {len(regions)} regions")
tot_st = collections.Counter()
for r in regions:
    tot_st.update(r["st"])
print("stall totals:", ", ".join(f"{k[6:]}={100*v/max(ts,1):.1f}%" for k, v in tot_st.most_common(8)))
for r in sorted(regions, key=lambda r: -r["samp"])[:18]:
    sig = " ".join(f"{k}={100*v/max(r['exec'],1):.0f}%" for k, v in r["ops"].most_common(5))
    st = " ".join(f"{k[6:]}={100*v/max(r['samp'],1):.0f}%" for k, v in r["st"].most_common(3))
    print(f"  {r['first']} n={r['n']:5d} exec={100*r['exec']/te:5.1f}% samples={100*r['samp']/max(ts,1):5.1f}% | {sig} | {st}")
