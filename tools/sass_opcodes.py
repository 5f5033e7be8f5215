"""Static SASS opcode histogram of the hot kernels in nd4js_b200/libnd4b.so (cuobjdump, no GPU needed): the evidence for which
hardware paths a kernel uses — DMMA (FP64 tensor pipe), FFMA2 / FMUL2 (packed FP32), UTMALDG / UTMASTG (TMA tile copies through tensor
maps), UBLKCP (TMA bulk copies), LDGSTS (cp.async).
python tools/sass_opcodes.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOT = ["matmul32_kernel", "gemm_pipe_kernel", "gemm_bulk_kernel", "chol16_kernel", "qr64x32_blocked_kernel", "svd64_pre32_kernel",
       "svd64_ortho_kernel", "svd64cb_kernel", "trisolve16_kernel", "qr_lstsq32_kernel", "svd_lstsq_kernel"]
out = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "nd4js_b200", "libnd4b.so")], capture_output=True, text=True).stdout
name, counts = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "SHFL", "MUFU", "LDGSTS", "UBLKCP", "UTMALDG", "UTMASTG", "DMMA")) else op.split(".")[0]
        counts[name][op] += 1
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
for n, c in counts.items():
    if not any(h in n for h in HOT):
        continue
    tot = sum(c.values())
    key = {k: c[k] for k in ("DMMA.8", "DFMA", "FFMA2", "FMUL2", "UTMALDG.2D", "UTMASTG.2D", "UBLKCP", "UBLKCP.S", "LDGSTS.E", "LDGSTS", "SYNCS", "FSEL", "MOV", "IMAD") if c.get(k)}
    print("== %s\n   %d instructions; %s" % (demangle(n)[:150], tot, ", ".join("%s %d" % kv for kv in key.items())))
    print("   top: " + ", ".join("%s %d" % kv for kv in c.most_common(10)))
