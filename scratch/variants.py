"""Build libuwip variants that differ in -D switches of csrc/gfpipe.cuh (only dehaze.cu is recompiled) and, on the GPU
box, time the chain kernels of each:  python scratch/variants.py build  |  python scratch/variants.py run"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uwimageproc_b200 import build as B

VARIANTS = {
    "dpad0": ["-DGP_DPAD=0"],
    "dpad2": ["-DGP_DPAD=2"],
    "dpad0b": ["-DGP_DPAD=0"],
    "dpad2b": ["-DGP_DPAD=2"],
}
VARIANT_SOURCES = ["dehaze.cu", "dehaze_gf1a.cu"]   # the translation units the -D switches apply to
OUT = os.path.join(ROOT, "scratch", "variants")

def build():
    os.makedirs(OUT, exist_ok=True)
    B.build()
    objs = [os.path.join(B.OBJ, s.replace(".cu", ".o")) for s in B.SOURCES if s not in VARIANT_SOURCES]
    for name, flags in VARIANTS.items():
        vobjs, sp = [], []
        for src in VARIANT_SOURCES:
            obj = os.path.join(OUT, "var_%s_%s.o" % (name, src.replace(".cu", "")))
            r = subprocess.run([B._nvcc()] + B.NVCC_FLAGS + flags + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, src), "-o", obj], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-3000:]
            sp += [l.strip() for l in r.stderr.splitlines() if "spill" in l and "0 bytes spill stores" not in l]
            vobjs.append(obj)
        lib = os.path.join(OUT, "libuwip_%s.so" % name)
        r = subprocess.run([B._nvcc(), "-shared", "-o", lib] + objs + vobjs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-lnvjpeg_static", "-lculibos", "-Xcompiler", "-fPIC"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        print(name, flags, "spills:", sp)

def run():
    for name in VARIANTS:
        env = dict(os.environ, UWIP_LIB=os.path.join(OUT, "libuwip_%s.so" % name))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--frames", "148", "--steps", "2", "--warmup", "3", "--no-e2e", "--no-cpu-baseline"],
                           capture_output=True, text=True, env=env)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(name, "fps %.1f" % d["value"], {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items() if k.startswith("dz_")}, flush=True)
        except Exception as e:
            print(name, "FAILED", r.stderr[-500:], flush=True)

def trace():
    """One chain pass with the trace build; prints per-role clock stamps (cycles, relative) of 64 rows of one CTA."""
    os.environ["UWIP_LIB"] = os.path.join(OUT, "libuwip_trace.so")
    import ctypes
    import numpy as np
    import torch
    import uwimageproc_b200 as u
    ctx = u.Context(0)
    n, W, H = 86, 3840, 2160
    d_in = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda")
    d_out = torch.empty_like(d_in)
    ctx.synth_dev(d_in, 0x5EED0004, 0, n, W, H)
    for _ in range(2):
        ctx.chain_dev(d_in, d_out, n, W, H)
    ctx.synchronize()
    buf = np.zeros((4, 64, 12), np.int64)
    rc = ctx.lib.uwip_exp_trace(ctypes.c_void_p(buf.ctypes.data))
    assert rc == 0, rc
    names = ["gf1a", "gf1b", "gf2a", "gf2b"]
    ev = ["acc_done", "acc_gotEMPTY", "acc_published", "aux_gotFULL", "aux_scanned", "sol_gotREADY", "sol_loaded", "sol_done"]
    for k in range(4):
        t0 = buf[k, 0, 0]
        print("==", names[k], " (cycles relative to acc_done of row 600)")
        print("row " + " ".join("%13s" % e for e in ev))
        for r in list(range(0, 12)) + [63]:
            print("%3d " % r + " ".join("%13d" % (buf[k, r, e] - t0) for e in range(8)))
        d = buf[k]
        per_row = (buf[k, 63, :8] - buf[k, 0, :8]) / 63.0
        print("cycles/row per event:", " ".join("%.0f" % v for v in per_row))
        d = buf[k]
        print("ACC inside: top->inputs ready %.0f  inputs->entering row done %.0f  leaving row %.0f   (a-kernels: cp.async wait / enter / leave; b: TMA wait / both rows / -)" % (
            (d[:, 9] - d[:, 8]).mean(), (d[:, 10] - d[:, 9]).mean() if k in (0, 2) else (d[:, 0] - d[:, 9]).mean(), (d[:, 0] - d[:, 10]).mean() if k in (0, 2) else 0.0))
        print("means: acc wait EMPTY %.0f  acc publish %.0f  aux scan %.0f  FULL->auxwake %.0f  READY->solwake %.0f  sol load %.0f  sol math %.0f" % (
            (d[:, 1] - d[:, 0]).mean(), (d[:, 2] - d[:, 1]).mean(), (d[:, 4] - d[:, 3]).mean(), (d[:, 3] - d[:, 2]).mean(),
            (d[:, 5] - d[:, 4]).mean(), (d[:, 6] - d[:, 5]).mean(), (d[:, 7] - d[:, 6]).mean()))


if __name__ == "__main__":
    {"build": build, "run": run, "trace": trace}[sys.argv[1]]()
