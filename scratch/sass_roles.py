#!/usr/bin/env python3
"""Opcode histogram of one kernel of an ncu report, split by how many warps execute each instruction (the roles of a
warp-specialised kernel run different code: instructions with equal execution counts belong to one role's loop).
usage: sass_roles.py report.ncu-rep kernel_regex"""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
base = kern.split("<")[0]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", base, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
ki = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kern in r[1])
hi = next(i for i in range(ki, len(rows)) if rows[i] and rows[i][0] == "Address")
hdr = rows[hi]
data = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r[0] == "Address":
        break
    data.append(r)
iS, iN, iI = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iI]) for r in data); tots = sum(int(r[iN]) for r in data)
# cluster by execution count (rounded to 2 significant digits)
cl = collections.defaultdict(list)
for k, r in enumerate(data):
    n = int(r[iI])
    if n == 0: continue
    key = float("%.2g" % n)
    cl[key].append((k, r))
print("total warp instr", tot, "samples", tots)
for key in sorted(cl, key=lambda k: -sum(int(r[iI]) for _, r in cl[k]))[:12]:
    rs = cl[key]
    ni = sum(int(r[iI]) for _, r in rs); ns = sum(int(r[iN]) for _, r in rs)
    ops = collections.Counter()
    st = collections.Counter()
    for _, r in rs:
        t = r[iS].split(); op = (t[1] if t[0].startswith("@") else t[0])
        op = ".".join(op.split(".")[:2]) if op.startswith(("I2F", "F2I", "LDS", "STS", "IMAD", "LDG", "STG")) else op.split(".")[0]
        ops[op] += 1
        for j in stall: st[hdr[j]] += int(r[j])
    print("exec/instr %.3g: %d sass lines (%d..%d), %.1f%% of instr, %.1f%% of samples" % (key, len(rs), rs[0][0], rs[-1][0], 100 * ni / tot, 100 * ns / tots))
    print("    ", ops.most_common(14))
    print("    ", st.most_common(6))
if len(sys.argv) > 3:
    lo, hi2 = int(sys.argv[3]), int(sys.argv[4])
    print("---- hottest lines in [%d, %d]" % (lo, hi2))
    sel = [(k, r) for k, r in enumerate(data) if lo <= k <= hi2]
    for k, r in sorted(sel, key=lambda kr: -int(kr[1][iN]))[:int(sys.argv[5]) if len(sys.argv) > 5 else 40]:
        top = sorted(((int(r[j]), hdr[j]) for j in stall), reverse=True)[:2]
        print(k, r[iN], r[iS][:70], top)
if len(sys.argv) > 6:
    for a, b in [tuple(map(int, x.split("-"))) for x in sys.argv[6].split(",")]:
        print("---- lines %d-%d" % (a, b))
        for k in range(a, b + 1):
            print(k, data[k][iN], data[k][iS][:90])
