import sys
sys.path.insert(0, '.')
import numpy as np, torch
import uwimageproc_b200 as u
ctx = u.Context(0)
def mm(t, n):
    o = t.cpu().numpy()
    return [(int(o[i].min()), int(o[i].max())) for i in range(n)]
for (W, H, n) in [(1920, 1080, 3), (3840, 2160, 1), (3840, 2160, 2), (3840, 2160, 3), (2560, 1440, 3), (3840, 1080, 3), (1920, 2160, 3)]:
    d_in = torch.empty((n, H, W, 3), dtype=torch.uint8, device='cuda')
    ctx.synth_dev(d_in, 0x5EED0004, 0, n, W, H)
    d_out = torch.zeros_like(d_in)
    ctx.chain_dev(d_in, d_out, n, W, H); ctx.synchronize()
    print(W, H, n, 'chain_dev', mm(d_out, n))
    a = torch.zeros_like(d_in); b = torch.zeros_like(d_in); c = torch.zeros_like(d_in)
    ctx.histretch_dev(d_in, a, n, W, H, "V", 1, 99); ctx.aclahe_dev(a, b, n, W, H, 2.0, (8, 8)); ctx.bgdehaze_dev(b, c, n, W, H); ctx.synchronize()
    print('   staged', mm(c, n), 'equal', bool((c == d_out).all()))
    p = ctx.chain_params(channels="xV")
    d2 = torch.zeros_like(d_in)
    ctx.chain_dev(d_in, d2, n, W, H, p); ctx.synchronize()
    print('   unfused head', mm(d2, n), 'equal staged', bool((c == d2).all()))
