"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = {}
fname = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] == 'Function Name' or hdr is None: continue
    if r[2] != '-':  # sass rows carry an address
        continue
    li = hdr.index('Instructions Executed'); si = hdr.index('# Samples')
    try: inst = int(r[li]); smp = int(r[si])
    except ValueError: continue
    key = (fname, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1]])
    a[0] += smp; a[1] += inst
tots = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print('total samples', tots, 'warp inst', toti)
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * a[0] / tots, 100.0 * a[1] / toti, k[0], k[1], a[2].strip()[:100]))

if len(sys.argv) > 3:
    # ranges: name:lo-hi,... on dehaze.cu
    for spec in sys.argv[3].split(','):
        name, rng = spec.split(':'); lo, hi = map(int, rng.split('-'))
        s = sum(a[0] for k, a in agg.items() if k[0] == 'dehaze.cu' and lo <= k[1] <= hi)
        i = sum(a[1] for k, a in agg.items() if k[0] == 'dehaze.cu' and lo <= k[1] <= hi)
        print("%-10s %5.1f%% smp %5.1f%% inst" % (name, 100.0 * s / tots, 100.0 * i / toti))
