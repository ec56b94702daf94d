"""Per-kernel time shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split('(')[0].replace('void ', '')
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += float(r[vi].replace(',', '')) / 1e6
tot = sum(v[1] for v in agg.values())
for k, (n, ms) in agg.items():
    print(f"{k:60s} n={n:4d} total={ms:9.3f} ms  avg={ms/n:8.4f} ms  share={100*ms/tot:5.1f}%")
