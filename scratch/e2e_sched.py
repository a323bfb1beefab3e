#!/usr/bin/env python3
"""Times uwip_chain_bgr8 (host buffers, H2D + D2H inside) under explicit sub-batch schedules (UWIP_E2E_SIZES).
usage: e2e_sched.py [frames]   - prints frames/s per schedule; run on the GPU box."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import uwimageproc_b200 as u

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W, H = 3840, 2160
SCHEDULES = {
    "default": None,
    "flat32": [16] + [32] * 7 + [16],
    "ramp55": [4, 8, 16, 24, 37, 55, 55, 32, 16, 6, 3],
    "default2": None,
}
ctx = u.Context(0)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
params = ctx.chain_params()
with torch.cuda.stream(stream):
    d_in = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda")
    ctx.synth_dev(d_in, 0x5EED0004, 0, n, W, H)
h_in = torch.empty((n, H, W, 3), dtype=torch.uint8).pin_memory()
h_out = torch.empty((n, H, W, 3), dtype=torch.uint8).pin_memory()
h_in.copy_(d_in)
torch.cuda.synchronize()
del d_in
ref = None
for name, sizes in SCHEDULES.items():
    if sizes is None:
        os.environ.pop("UWIP_E2E_SIZES", None)
    else:
        if sum(sizes) != n:
            continue
        os.environ["UWIP_E2E_SIZES"] = ",".join(map(str, sizes))
    ctx.chain_host_ptr(h_in.data_ptr(), h_out.data_ptr(), n, W, H, params)
    torch.cuda.synchronize()
    best = 1e9
    tot = 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        ctx.chain_host_ptr(h_in.data_ptr(), h_out.data_ptr(), n, W, H, params)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = min(best, dt); tot += dt
    crc = int(h_out.view(-1)[::4099].to(torch.int64).sum())
    if ref is None:
        ref = crc
    print("%-8s mean %.1f fps  best %.1f fps  same=%s  %s" % (name, 3 * n / tot, n / best, crc == ref, sizes), flush=True)
