"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
python tools/launch_summary.py profiles/r02_launches_bench.csv > profiles/r02_launches_bench_summary.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10]
h = rows[0]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("nd4b::", "")
    ns = float(r[vi].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1)
    tot[name] += ns
    cnt[name] += 1
s = sum(tot.values())
print("kernel,launches,total_ns,share")
for k, v in tot.most_common():
    print("%s,%d,%d,%.4f" % (k, cnt[k], v, v / s))
