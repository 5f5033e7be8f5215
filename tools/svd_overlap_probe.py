"""Does the FP32 stage of one part of a C5 batch overlap with the FP64 stage of another when the parts are issued on different
streams?  (The two stages are bound by different pipes: one FP64 CTA of 255 registers and two FP32 CTAs of 128 fit one SM.)
Times svd_jac_1sided [16384,64,64] as 1, 2, 4, 8 parts on as many streams."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import nd4js_b200 as nd  # noqa: E402

nd.init([0])
lib = nd.load()
total = 16384
f64 = dict(dtype=torch.float64, device="cuda")
g = torch.Generator(device="cuda").manual_seed(7)
a = torch.rand(total, 64, 64, generator=g, **f64) * 2 - 1
u, sv, v = torch.empty(total, 64, 64, **f64), torch.empty(total, 64, **f64), torch.empty(total, 64, 64, **f64)
ws = lib.nd4b_dev_svd_workspace(total, 64, 64)
work = torch.empty(ws // 8 + 64, **f64)
p = lambda t: C.c_void_p(t.data_ptr())


def run(parts, streams):
    n = total // parts
    wper = (ws // parts) // 8
    main = torch.cuda.current_stream()
    ev0 = torch.cuda.Event()
    ev0.record(main)
    for i in range(parts):
        st = streams[i % len(streams)]
        st.wait_event(ev0)
        sl = slice(i * n, (i + 1) * n)
        rc = lib.nd4b_dev_svd_jac1_f64(0, C.c_void_p(st.cuda_stream), p(a[sl]), p(u[sl]), p(sv[sl]), p(v[sl]), n, 64, 64, None,
                                       C.c_void_p(work.data_ptr() + i * wper * 8), wper * 8)
        assert rc == 0, lib.nd4b_last_error()
    for st in streams:
        e = torch.cuda.Event()
        e.record(st)
        main.wait_event(e)


for parts, ns in ((1, 1), (2, 2), (4, 2), (4, 4), (8, 2), (8, 4), (8, 8), (16, 4)):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    run(parts, streams)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run(parts, streams)
    e1.record()
    torch.cuda.synchronize()
    print("parts %2d on %d streams: %.2f ms per batch" % (parts, ns, e0.elapsed_time(e1) / 5))
