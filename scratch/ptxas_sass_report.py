#!/usr/bin/env python3
"""Writes profiles/r2_ptxas.txt (ptxas -v per kernel) and profiles/r2_sass.txt (mnemonic census of the hot kernels and short
excerpts around the instructions DESIGN.md names) for the current sources.  CPU only: nvcc cross-compiles sm_100a."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uwimageproc_b200 import build as B

def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return r.stdout.splitlines()

def ptxas():
    rows = []
    for src in B.SOURCES:
        r = subprocess.run([B._nvcc()] + B.NVCC_FLAGS + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, src), "-o", "/tmp/_rep.o"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        cur = None
        for l in r.stderr.splitlines():
            m = re.search(r"Compiling entry function '(\S+)'", l)
            if m: cur = {"name": m.group(1), "src": src}; rows.append(cur); continue
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", l)
            if m and cur: cur.update(stack=int(m.group(1)), st=int(m.group(2)), ld=int(m.group(3))); continue
            m = re.search(r"Used (\d+) registers", l)
            if m and cur:
                cur["regs"] = int(m.group(1))
                sm = re.search(r"(\d+) bytes smem", l)
                cur["smem"] = int(sm.group(1)) if sm else 0
    for row, d in zip(rows, demangle([r["name"] for r in rows])):
        row["dn"] = re.sub(r"\(.*", "", d)
    with open(os.path.join(ROOT, "profiles", "r2_ptxas.txt"), "w") as f:
        f.write("# ptxas -v (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3) per kernel of libuwip.so, round 2 final build\n")
        f.write("# (scratch/ptxas_sass_report.py).  registers = per thread at launch; the gp_kernel<...> marches re-split their 128 launch\n")
        f.write("# registers per role with setmaxnreg: GF1a (narrow strips, dehaze_gf1a.cu) ACC 176 / AUX 88 / SOLVE 120; wide strips\n")
        f.write("# (dehaze.cu) ACC+AUX / SOLVE: GF2a 144 / 112, plane readers 136 / 120.  static smem only (dynamic smem is set at launch).\n\n")
        f.write("| kernel | source | registers | spill stores B | spill loads B | stack B | static smem B |\n|---|---|---|---|---|---|---|\n")
        for r in rows:
            f.write("| %s | %s | %d | %d | %d | %d | %d |\n" % (r["dn"], r["src"], r.get("regs", -1), r.get("st", 0), r.get("ld", 0), r.get("stack", 0), r.get("smem", 0)))
    print("ptxas:", len(rows), "kernels;", sum(1 for r in rows if r.get("st", 0) or r.get("ld", 0)), "with spills:",
          [(r["dn"], r["st"], r["ld"]) for r in rows if r.get("st", 0) or r.get("ld", 0)])

def sass():
    out = ["# SASS of libuwip.so (cuobjdump -sass, sm_100a), round 2 final build (scratch/ptxas_sass_report.py): mnemonic census of the",
           "# hot kernels and excerpts around the instructions DESIGN.md names\n"]
    for obj in ("dehaze_gf1a.o", "dehaze.o", "clahe.o"):
        r = subprocess.run(["cuobjdump", "-sass", os.path.join(B.OBJ, obj)], capture_output=True, text=True)
        funcs = {}
        cur = None
        for l in r.stdout.splitlines():
            m = re.search(r"Function : (\S+)", l)
            if m: cur = m.group(1); funcs[cur] = []; continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
            if m and cur: funcs[cur].append(m.group(1).strip())
        names = list(funcs)
        for mn, dn in zip(names, demangle(names)):
            dn = re.sub(r"\(.*", "", dn)
            if not any(k in dn for k in ("gp_kernel", "window15", "clahe_apply_kernel<1, true>")): continue
            ins = funcs[mn]
            c = collections.Counter()
            for i in ins:
                t = i.split()
                op = t[1] if t[0].startswith("@") else t[0]
                c[op] += 1
            keep = [k for k in c if re.match(r"(USETMAXREG|SYNCS|UBLKCP|LDGSTS|DFMA|DADD|DMUL|MUFU\.RCP64H|IMAD$|LDS|STS|STG|LDG|F2I|I2F|SHFL|LDL|STL|BAR|VIMNMX|PRMT|REDUX|FADD|FMUL|FFMA|NANOSLEEP)", k)]
            out.append("## %s (%s): %d instructions" % (dn, obj, len(ins)))
            out.append("   " + ", ".join("%s x%d" % (k, c[k]) for k in sorted(keep, key=lambda k: -c[k])))
            def excerpt(title, pat, before=6, after=10):
                for k, i in enumerate(ins):
                    if re.search(pat, i):
                        out.append("### " + title)
                        out.extend("    " + x for x in ins[max(0, k - before):k + after])
                        return
            if "PipGF1a" in dn:
                excerpt("register re-split per role (setmaxnreg)", r"USETMAXREG", 2, 4)
                excerpt("ACC accumulate: guide moments IMAD in place, signal moments DFMA on exact integer-valued doubles", r"DFMA", 10, 14)
                excerpt("SOLVE window sums, fp64 moments: the same five loads per moment for both halves of a quad, no branch", r"LDS\.128 .*\+0x[0-9a-f]+\]", 8, 16)
            if "PipGF1b" in dn:
                excerpt("TMA producer: bulk copy of a coefficient row into the ring, completion on an mbarrier", r"UBLKCP", 6, 6)
            if "PipGF2b" in dn:
                excerpt("SOLVE inputs staged two rows ahead (cp.async into the thread's own ring slots)", r"LDGSTS", 4, 8)
            if "window15" in dn:
                excerpt("window runs: three-input packed min/max", r"VIMNMX3\.U16x2", 2, 12)
                excerpt("tile minimum of the integer differences", r"REDUX", 3, 5)
            if "clahe_apply" in dn:
                excerpt("float -> byte by magic add, branch-free sector weights", r"FADD\.RZ", 8, 10)
            out.append("")
    open(os.path.join(ROOT, "profiles", "r2_sass.txt"), "w").write("\n".join(out) + "\n")
    print("sass: written", len(out), "lines")

if __name__ == "__main__":
    B.build()
    ptxas()
    sass()
