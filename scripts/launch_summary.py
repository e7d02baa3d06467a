"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total, share, mean."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = None
agg = collections.OrderedDict()
seen = 0
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
        seen += 1
        if seen <= skip:
            continue
        k = d["Kernel Name"].split("(")[0][-60:]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:60s} {v[0]:5d} {v[1] / 1e6:10.3f} ms {100 * v[1] / tot:5.1f}%  {v[1] / v[0] / 1e3:9.1f} us/launch")
