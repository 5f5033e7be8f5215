"""Times one device-resident nd4b_dev_matmul_f64 for a few (I, K, J): python tools/gemm_sweep.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import nd4js_b200 as nd  # noqa: E402

nd.init([0])
lib = nd.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
for (i, k, j) in [(512, 512, 512), (512, 2048, 512), (512, 8192, 512), (1024, 512, 1024), (1024, 1024, 1024), (2048, 2048, 2048), (4096, 4096, 4096)]:
    a = torch.rand(i, k, dtype=torch.float64, device="cuda")
    b = torch.rand(k, j, dtype=torch.float64, device="cuda")
    c = torch.empty(i, j, dtype=torch.float64, device="cuda")
    reps = 200 if i * j * k < 2 ** 31 else 20
    for _ in range(5):
        lib.nd4b_dev_matmul_f64(0, st, p(a), 0, p(b), 0, p(c), 1, i, k, j)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        lib.nd4b_dev_matmul_f64(0, st, p(a), 0, p(b), 0, p(c), 1, i, k, j)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    print("%5d x %5d x %5d: %9.2f us  %6.2f TFLOP/s" % (i, k, j, us, 2.0 * i * k * j / us / 1e6))
# launch period of an empty-ish kernel through the same path (1x1x1 product)
a = torch.rand(1, 1, dtype=torch.float64, device="cuda"); c = torch.empty(1, 1, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(500):
    lib.nd4b_dev_matmul_f64(0, st, p(a), 0, p(a), 0, p(c), 1, 1, 1, 1)
e1.record()
torch.cuda.synchronize()
print("1x1x1 launch period: %.2f us" % (e0.elapsed_time(e1) * 1e3 / 500))
