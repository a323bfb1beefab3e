"""Build libuwip variants that differ in -D switches of csrc/gfpipe.cuh (only dehaze.cu is recompiled) and, on the GPU
box, time the chain kernels of each:  python scratch/variants.py build  |  python scratch/variants.py run"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from uwimageproc_b200 import build as B

VARIANTS = {
    "pipe": ["-DGP_GF1A_ACC_REGS=152"],
    "nopipe": ["-DGP_GF1A_ACC_REGS=152", "-DGP_PIPE_A=0", "-DGP_PIPE_B=0"],
    "pipe_a104": ["-DGP_GF1A_ACC_REGS=152", "-DGP_GF2A_ACC_REGS=104"],
}
OUT = os.path.join(ROOT, "scratch", "variants")

def build():
    os.makedirs(OUT, exist_ok=True)
    B.build()
    objs = [os.path.join(B.OBJ, s.replace(".cu", ".o")) for s in B.SOURCES if s != "dehaze.cu"]
    for name, flags in VARIANTS.items():
        obj = os.path.join(OUT, "dehaze_%s.o" % name)
        r = subprocess.run([B._nvcc()] + B.NVCC_FLAGS + flags + ["-Xptxas", "-v", "-c", os.path.join(B.CSRC, "dehaze.cu"), "-o", obj], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        sp = [l for l in r.stderr.splitlines() if "spill" in l and "0 bytes spill stores" not in l]
        lib = os.path.join(OUT, "libuwip_%s.so" % name)
        r = subprocess.run([B._nvcc(), "-shared", "-o", lib] + objs + [obj, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-Xcompiler", "-fPIC"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        print(name, flags, "spills:", len(sp))

def run():
    for name in VARIANTS:
        env = dict(os.environ, UWIP_LIB=os.path.join(OUT, "libuwip_%s.so" % name))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--frames", "86", "--steps", "2", "--warmup", "3", "--no-e2e", "--no-cpu-baseline"],
                           capture_output=True, text=True, env=env)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(name, "fps %.1f" % d["value"], {k: round(v["ms_per_step"], 2) for k, v in d["kernels"].items() if k.startswith("dz_gf")}, flush=True)
        except Exception as e:
            print(name, "FAILED", r.stderr[-500:], flush=True)

if __name__ == "__main__":
    {"build": build, "run": run}[sys.argv[1]]()
