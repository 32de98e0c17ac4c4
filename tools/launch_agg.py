"""aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel over the LAST third of the launches (= the last of
three identical steps):  python tools/launch_agg.py gpurun_out/x_launches.csv [parts]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 3
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = rows[hi + 1:]
last = data[len(data) - len(data) // parts:]
agg = collections.OrderedDict()
for r in last:
    a = agg.setdefault(r[kn].split("(")[0], [0, 0.0])
    a[0] += 1
    a[1] += float(r[mv].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print("%d launches in the file; last part: %d launches, %.1f us" % (len(data), len(last), tot / 1e3))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-40s %4d %10.1f us %5.1f%%" % (k, a[0], a[1] / 1e3, 100 * a[1] / tot))
