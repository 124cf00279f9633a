"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
`bench.py` into per-iteration DRAM traffic and per-kernel averages (profiles/*_dram_traffic.json).
usage: python scripts/dram_traffic.py <launches.csv> <config description> [<protocol note>] > profiles/rNN_dram_traffic.json"""
import collections
import csv
import json
import sys

path, desc = sys.argv[1], sys.argv[2]
protocol = sys.argv[3] if len(sys.argv) > 3 else None
rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit() and "mmg::" in r[4]]
per = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0].replace("void ", "").replace("mmg::", "").split("<")[0]
    per.setdefault(name, collections.defaultdict(list))[r[12]].append(float(r[14].replace(",", "")))
iters = max(len(v["gpu__time_duration.sum"]) for k, v in per.items() if k != "k_init_rng")
out = {"source": path, "config": desc, "iterations_profiled": iters, "kernels_us": {}, "kernels_dram_read_kb": {},
       "kernels_dram_write_kb": {}}
tot_r = tot_w = tot_t = 0.0
for k, v in per.items():
    if k == "k_init_rng":
        continue
    t = sum(v["gpu__time_duration.sum"]) / iters / 1e3
    rd = sum(v["dram__bytes_read.sum"]) / iters
    wr = sum(v["dram__bytes_write.sum"]) / iters
    out["kernels_us"][k] = round(t, 1)
    out["kernels_dram_read_kb"][k] = round(rd / 1e3, 1)
    out["kernels_dram_write_kb"][k] = round(wr / 1e3, 1)
    tot_r += rd; tot_w += wr; tot_t += t
out["per_iteration"] = {"dram_read_bytes": tot_r, "dram_write_bytes": tot_w, "total_bytes": tot_r + tot_w,
                        "kernel_time_us_ncu": round(tot_t, 1)}
if protocol:
    out["protocol"] = protocol
print(json.dumps(out, indent=1))
