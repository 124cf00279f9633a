"""Per-kernel averages (duration, DRAM bytes) of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list."""
import collections
import csv
import sys

for path in sys.argv[1:]:
    hdr, agg = None, collections.OrderedDict()
    for r in csv.reader(open(path)):
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        k = d["Kernel Name"].split("(")[0][:50]
        agg.setdefault(k, {}).setdefault(d["Metric Name"], []).append(float(d["Metric Value"].replace(",", "")))
    print(path)
    tot = 0.0
    for k, m in agg.items():
        t = m.get("gpu__time_duration.sum", [0.0])
        us = sum(t) / len(t) / 1000
        rd = sum(m.get("dram__bytes_read.sum", [0.0])) / len(t) / 1024
        wr = sum(m.get("dram__bytes_write.sum", [0.0])) / len(t) / 1024
        print("  %-52s n=%3d avg=%7.2f us  dram rd %8.1f KB wr %8.1f KB" % (k, len(t), us, rd, wr))
        if "mmg::k_" in k and "init_rng" not in k:
            tot += us
    print("  sum of the iteration's kernels: %.1f us" % tot)
