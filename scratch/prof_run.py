import sys
sys.path.insert(0, '.')
import torch
import uwimageproc_b200 as u
W, H, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
ctx = u.Context(0)
d_in = torch.empty((n, H, W, 3), dtype=torch.uint8, device='cuda')
d_out = torch.empty_like(d_in)
torch.cuda.synchronize()
ctx.synth_dev(d_in, 0x5EED0004, 0, n, W, H)
for _ in range(reps):
    ctx.chain_dev(d_in, d_out, n, W, H)
ctx.synchronize()
print('ok', ctx.launch_count())
