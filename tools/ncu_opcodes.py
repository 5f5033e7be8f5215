"""Opcode histogram (warp instructions executed, share, stall samples) of one kernel from the source page of an .ncu-rep:
python tools/ncu_opcodes.py report.ncu-rep [top]."""
import collections
import csv
import io
import subprocess
import sys


def main(path, top=28):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    k = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[k]
    si, ei, st = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    cnt, stall = collections.Counter(), collections.Counter()
    for r in rows[k + 1:]:
        if len(r) <= ei or not r[ei].isdigit():
            continue
        toks = r[si].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "SHFL", "MUFU")) else op.split(".")[0]
        cnt[op] += int(r[ei])
        stall[op] += int(r[st]) if r[st].isdigit() else 0
    tot, stot = sum(cnt.values()), max(1, sum(stall.values()))
    print("total warp instructions %d, stall samples %d" % (tot, stot))
    for op, c in cnt.most_common(int(top)):
        print("  %-14s %14d  %5.1f %%   stall samples %5.1f %%" % (op, c, 100.0 * c / tot, 100.0 * stall[op] / stot))


if __name__ == "__main__":
    main(*sys.argv[1:])
