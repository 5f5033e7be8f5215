"""Splits the SASS page of an .ncu-rep into regions by executed-instruction count (loops show up as multiples of the
grid's warp count) and prints samples per region and per opcode: python tools/ncu_regions.py rep.ncu-rep [warps]"""
import collections
import csv
import io
import re
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
tot = sum(int(r[iS]) for r in data)
warps = int(sys.argv[2]) if len(sys.argv) > 2 else max(int(r[iI]) for r in data[:20])
print("total samples", tot, "warps", warps)
regions = []
cur = None
for r in data:
    n = int(r[iI])
    mult = round(n / warps, 2)
    if cur is None or cur[0] != mult:
        cur = [mult, 0, 0, collections.Counter()]
        regions.append(cur)
    cur[1] += int(r[iS])
    cur[2] += 1
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iSrc])
    cur[3][m.group(2) if m else "?"] += int(r[iS])
for mult, smp, n, ops in regions:
    if smp * 200 < tot:
        continue
    print("x%-5s %5d instr %6.2f%% samples  %s" % (mult, n, 100.0 * smp / tot,
          " ".join("%s:%.1f" % (o, 100.0 * c / tot) for o, c in ops.most_common(6))))
