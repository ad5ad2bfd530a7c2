"""Aggregate an ncu --metrics gpu__time_duration.sum --csv launch list by kernel name.
usage: python tools/agg_launches.py launches.csv"""
import csv, collections, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    k = re.sub(r'\(.*', '', r[ki])[:90]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-90s n=%6d  %10.3f ms  %5.1f%%" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot))
print("total %.3f ms" % (tot / 1e6))
