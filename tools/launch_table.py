"""Per-launch table of an ncu launch list with time and DRAM bytes (`--metrics gpu__time_duration.sum,dram__bytes_read.sum,
dram__bytes_write.sum --csv`), in launch order, plus per-kernel totals. usage: launch_table.py file.csv [title]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
title = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
hdr = None
L = collections.OrderedDict()
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = d["Metric Unit"]
        v *= {"us": 1.0, "ns": 1e-3, "ms": 1e3, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
        e = L.setdefault(d["ID"], {"name": d["Kernel Name"].split("(")[0].replace("void ", "")})
        e[d["Metric Name"]] = v
print(title)
print("source:", sys.argv[1], "(ncu launch list: launches are serialised and start cold; shares, not absolutes)")
print("%4s %-40s %10s %10s %10s" % ("#", "kernel", "time us", "dram rd MB", "dram wr MB"))
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for k, e in L.items():
    t, rd, wr = e.get("gpu__time_duration.sum", 0.0), e.get("dram__bytes_read.sum", 0.0), e.get("dram__bytes_write.sum", 0.0)
    print("%4s %-40s %10.1f %10.1f %10.1f" % (k, e["name"][:40], t, rd, wr))
    a = agg[e["name"]]
    a[0] += 1
    a[1] += t
    a[2] += rd + wr
tot = sum(a[1] for a in agg.values())
print("\ntotal %.3f ms over %d launches" % (tot / 1e3, sum(a[0] for a in agg.values())))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-40s n=%4d %9.1f us %5.1f%%  dram %8.1f MB" % (k[:40], a[0], a[1], 100 * a[1] / tot, a[2]))
