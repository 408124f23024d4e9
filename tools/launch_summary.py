"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel. usage: launch_summary.py file.csv"""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if len(r) > 5 and r[0] == 'ID':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        v *= {'us': 1e-3, 'ns': 1e-6, 'ms': 1.0}.get(d['Metric Unit'], 1.0)
        agg[d['Kernel Name'].split('(')[0]][0] += 1
        agg[d['Kernel Name'].split('(')[0]][1] += v
tot = sum(v[1] for v in agg.values())
print("total %.3f ms over %d launches" % (tot, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-44s n=%5d  %9.3f ms  %5.1f%%  avg %8.1f us" % (k[:44], v[0], v[1], 100 * v[1] / tot, 1e3 * v[1] / v[0]))
