import sys
sys.path.insert(0, '.')
import numpy as np
import uwimageproc_b200 as u
from oracle import uwip_oracle as O
ctx = u.Context(0)
for (w, h, n) in [(320, 180, 3), (213, 97, 2), (1000, 300, 1)]:
    frames = np.stack([O.synth_frame(0x5EED0004, i, w, h) for i in range(n)])
    out = ctx.chain(frames)
    print(w, h, n, int(out.sum()))
fr = O.synth_frame(0x5EED0003, 2, 212, 118)
for r in (7, 8):
    rb, rg = ctx.refined_transmission(fr, ctx.dehaze_params(radius=r))
    print('r', r, float(rb.sum()))
print('done', ctx.launch_count())
