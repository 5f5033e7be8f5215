"""Shared-memory wavefronts per opcode and per (wavefronts / execution) class from the source page of an .ncu-rep, plus the
share of stall samples by stall reason: python tools/ncu_shared_wf.py report.ncu-rep [units]"""
import collections
import csv
import io
import subprocess
import sys


def main(path, units=65536):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    for n, k in enumerate(starts):
        end = starts[n + 1] - 1 if n + 1 < len(starts) else len(rows)
        print("==", rows[k - 1][1][:110] if k and len(rows[k - 1]) > 1 else "")
        one(rows[k:end], units)


def one(rows, units):
    k = 0
    h = rows[k]
    si, ei = h.index("Source"), h.index("Instructions Executed")
    wi, xi = h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Excessive")
    reasons = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    tab, stall = collections.defaultdict(lambda: [0, 0, 0, 0]), collections.Counter()
    for r in rows[k + 1:]:
        if len(r) <= wi or not r[ei].isdigit():
            continue
        for c in reasons:
            v = r[h.index(c)]
            stall[c] += int(v) if v.isdigit() else 0
        wf = int(r[wi]) if r[wi].isdigit() else 0
        if not wf:
            continue
        toks = r[si].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ex = int(r[ei])
        key = (".".join(op.split(".")[:2]), round(wf / max(ex, 1)))
        t = tab[key]
        t[0] += 1; t[1] += ex; t[2] += wf; t[3] += int(r[xi]) if r[xi].isdigit() else 0
    tot = sum(t[2] for t in tab.values())
    print("shared-memory wavefronts: %d = %.0f per unit" % (tot, tot / units))
    for (op, per), t in sorted(tab.items(), key=lambda kv: -kv[1][2]):
        print("  %-8s %2d wf/exec  %3d static  %7.1f exec/unit  %7.1f wf/unit  %7.1f excess/unit" % (op, per, t[0], t[1] / units, t[2] / units, t[3] / units))
    s = sum(stall.values())
    print("stall samples by reason:", ", ".join("%s %.1f%%" % (c[6:], 100.0 * v / s) for c, v in stall.most_common(8)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 65536)
