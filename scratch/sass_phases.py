#!/usr/bin/env python3
"""Per-phase (between BAR.SYNCs) instruction / stall-sample breakdown of one kernel of an ncu report.
usage: sass_phases.py report.ncu-rep kernel_name"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[0] == "Address":
        break   # first launch only
    data.append(r)
iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot_i = sum(int(r[iI]) for r in data)
tot_s = sum(int(r[iN]) for r in data)
print("sass lines", len(data), "warp instructions", tot_i, "samples", tot_s)
seg = 0
ops = collections.defaultdict(collections.Counter)
sm, si = collections.Counter(), collections.Counter()
st = collections.defaultdict(collections.Counter)
for r in data:
    t = r[iS].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[seg][op] += int(r[iI]); si[seg] += int(r[iI]); sm[seg] += int(r[iN])
    for j in stall:
        st[seg][hdr[j]] += int(r[j])
    if "BAR.SYNC" in r[iS]:
        seg += 1
for s in sorted(si):
    print("phase", s, "instr %.1f%%" % (100 * si[s] / max(tot_i, 1)), "samples %.1f%%" % (100 * sm[s] / max(tot_s, 1)),
          ops[s].most_common(10))
    print("    stalls", st[s].most_common(5))
