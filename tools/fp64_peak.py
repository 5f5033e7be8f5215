"""Measures the FP64 roofline denominators on the current GPU (MEASURED_PEAKS.json has no fp64 entry):
DFMA vector-pipe peak, DMMA.8x8x4 tensor-pipe peak, and a device copy bandwidth cross-check."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import nd4js_b200 as nd  # noqa: E402

nd.init([0])
lib = nd.load()
out = {}
ms = C.c_float(0)
for which, name, per in ((0, "dfma", lambda it, bl, th: 2.0 * 16 * it * bl * th), (1, "dmma884", lambda it, bl, th: 512.0 * 8 * it * bl * (th // 32))):
    best = 0.0
    for threads in (128, 256, 512, 1024):
        for mult in (1, 2, 4, 8):
            blocks, iters = 148 * mult, 8192
            for _ in range(2):
                assert lib.nd4b_probe_fp64(0, which, iters, blocks, threads, C.byref(ms)) == 0
            tf = per(iters, blocks, threads) / (ms.value * 1e-3) / 1e12
            out["%s_t%d_b%d" % (name, threads, blocks)] = round(tf, 2)
            best = max(best, tf)
    out[name + "_peak_tflops"] = round(best, 2)
a = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
b = torch.empty_like(a)
best = 0.0
for _ in range(10):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best = max(best, 2 * a.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
out["copy_gbs"] = round(best, 1)
print(json.dumps(out))
