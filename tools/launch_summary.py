"""condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel launches/step, us/step and share
   python tools/launch_summary.py launches.csv <steps> ["header line" ...]"""
import csv
import re
import sys
from collections import OrderedDict

path, steps = sys.argv[1], int(sys.argv[2])
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
agg = OrderedDict()
for r in rows[1:]:
    if r[ci["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ci["Kernel Name"]])[:50]
    ns = float(r[ci["Metric Value"]].replace(",", ""))
    if r[ci["Metric Unit"]] in ("us", "usecond"):
        ns *= 1e3
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ns
total = sum(a[1] for a in agg.values())
for h in sys.argv[3:]:
    print("# " + h)
for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-50s launches/step %5.2f  us/step %8.2f  share %5.1f%%" % (name, n / steps, ns / 1e3 / steps, 100.0 * ns / total))
print("total us/step %.1f   launches/step %.1f" % (total / 1e3 / steps, sum(a[0] for a in agg.values()) / steps))
