import sys, time
sys.path.insert(0, '.')
import numpy as np
import uwimageproc_b200 as u
from oracle import uwip_oracle as O
ctx = u.Context(0)
W, H = int(sys.argv[1]), int(sys.argv[2])
fr = O.synth_frame(0x5EED0004, 0, W, H)
a = ctx.histretch(fr, "V", 1, 99); b = ctx.aclahe(a, 2.0, (8, 8))
print('aclahe min/max', b.reshape(-1,3).min(0), b.reshape(-1,3).max(0))
out8, outf = ctx.bgdehaze(b, return_float=True)
print('gpu out8 min/max', out8.min(), out8.max(), 'nan', np.isnan(outf).sum(), 'fmin/fmax', np.nanmin(outf), np.nanmax(outf))
B, idx = ctx.background_light(b); print('B', B, idx)
rb, rg = ctx.refined_transmission(b); print('t_ref nan', np.isnan(rb).sum(), np.isnan(rg).sum(), rb.min(), rg.min(), rb.max(), rg.max())
rest = ctx.rc_correction(b); print('restored nan', np.isnan(rest).sum(), rest.reshape(-1,3).min(0), rest.reshape(-1,3).max(0))
if len(sys.argv) > 3:
    st = {}
    t = time.time(); out, o8 = O.bgdehaze_frame(b, 15, st); print('oracle %.0fs' % (time.time()-t))
    print('B oracle', st['B'])
    print('t_ref err', np.abs(rb - st['t_blue']).max(), np.abs(rg - st['t_green']).max())
    print('restored err', np.abs(rest - st['restored']).max())
    print('out err', np.nanmax(np.abs(outf - out)), 'out8 maxdiff', np.abs(out8.astype(int) - o8.astype(int)).max())
