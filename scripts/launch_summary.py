"""Per-kernel averages of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 5 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r[4].split('(')[0][:60], []).append(float(r[-1].replace(',', '')))
    print(path)
    for k, v in agg.items():
        print("  %-60s n=%3d avg=%8.1f us" % (k, len(v), sum(v) / len(v) / 1000))
