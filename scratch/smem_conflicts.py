#!/usr/bin/env python3
"""Shared-memory wavefronts per SASS line of the kernels of an ncu report: actual, ideal, excess (bank conflicts).
usage: smem_conflicts.py report.ncu-rep kernel_base_name [top]"""
import csv, subprocess, sys
rep, base = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 10
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", base, "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
names = [r[1] for r in rows if r and r[0] == "Kernel Name"]
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
seen = set()
for name, hi in zip(names, his):
    if name in seen: continue
    seen.add(name)
    h = rows[hi]
    iS, iW, iI, iE, iX = h.index("Source"), h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Ideal"), h.index("L1 Wavefronts Shared Excessive"), h.index("Instructions Executed")
    data = []
    for r in rows[hi + 1:]:
        if len(r) != len(h) or r[0] == "Address": break
        data.append(r)
    tot = sum(int(r[iW] or 0) for r in data); exc = sum(int(r[iE] or 0) for r in data)
    print("==", name[:60], "wavefronts %d excessive %d (%.0f%%)" % (tot, exc, 100.0 * exc / max(tot, 1)))
    for k, r in sorted(enumerate(data), key=lambda kr: -int(kr[1][iE] or 0))[:top]:
        print("   line %5d exec %9s  wavefronts %9s ideal %9s excess %9s  %s" % (k, r[iX], r[iW], r[iI], r[iE], r[iS][:60]))
