"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    n, v = r[ki][:70], float(r[vi].replace(",", ""))
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-72s n=%3d total %.3f ms avg %.1f us" % (n, c, t / 1e6, t / c / 1e3))
