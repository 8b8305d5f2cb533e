"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel.
usage: python scripts/launch_summary.py launches.csv > summary.txt"""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    a = agg.setdefault(r[ki], [0.0, 0])
    a[0] += ms; a[1] += 1
tot = sum(a[0] for a in agg.values())
print("ncu launch list of `python bench.py --steps 1 --warmup 1` (both steps; cold-cache serialised times; compare SHARES)")
print("total %.1f ms over %d launches" % (tot, sum(a[1] for a in agg.values())))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%6.2f ms  %5.1f%%  x%-3d %s" % (a[0], 100 * a[0] / tot, a[1], k[:64]))
