"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares."""
import csv, sys, re, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).strip()
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v_us = v / 1000.0 if u in ("ns", "nsecond") else (v * 1000.0 if u in ("ms", "msecond") else v)
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1; t[1] += v_us
s = sum(t[1] for t in tot.values())
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.1f | %.1f%% |" % (k, n, us, 100 * us / s))
