"""Aggregates the ncu source page (SASS view) of a report: stall samples by opcode and by stall reason."""
import csv, io, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
si, ci, ei = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by_op = collections.Counter(); cnt_op = collections.Counter(); by_reason = collections.Counter(); tot = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    op = r[si].split()[0] if not r[si].strip().startswith("@") else r[si].split()[1]
    s = int(r[ci] or 0); tot += s
    by_op[op] += s; cnt_op[op] += int(r[ei] or 0)
    for h in reasons:
        by_reason[h] += int(r[hdr.index(h)] or 0)
print("total samples", tot)
print("by stall reason:", [(k, v) for k, v in by_reason.most_common(8)])
print("by opcode (samples, share, executed warp-instr):")
for op, s in by_op.most_common(14):
    print("  %-22s %8d %5.1f%% %12d" % (op, s, 100.0 * s / max(tot, 1), cnt_op[op]))
