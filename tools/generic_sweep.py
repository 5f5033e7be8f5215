"""Device-resident timing of the generic (any-shape) kernels for a few shapes: python tools/generic_sweep.py"""
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
f64 = dict(dtype=torch.float64, device="cuda")


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (b, r, c) in [(16384, 32, 32), (16384, 48, 24), (8192, 128, 32), (4096, 32, 64)]:
    a = torch.rand(b, r, c, **f64) * 2 - 1
    l = min(r, c)
    q, rr = torch.empty(b, r, l, **f64), torch.empty(b, l, c, **f64)
    wb = lib.nd4b_dev_qr_workspace(b, r, c)
    w = torch.empty(wb // 8 + 1, **f64)
    ms = timeit(lambda: lib.nd4b_dev_qr_f64(0, st, p(a), p(q), p(rr), b, r, c, p(w), C.c_size_t(wb)))
    print("qr      [%6d,%3d,%3d]: %8.3f ms  %9.0f matrices/s  %6.1f GB/s" % (b, r, c, ms, b / ms * 1e3, (2 * r * c + l * c) * 8 * b / ms / 1e6))
for (b, r, c) in [(4096, 32, 32), (2048, 64, 32), (2048, 48, 48)]:
    a = torch.rand(b, r, c, **f64) * 2 - 1
    l = min(r, c)
    u, sv, v = torch.empty(b, r, l, **f64), torch.empty(b, l, **f64), torch.empty(b, l, c, **f64)
    wb = lib.nd4b_dev_svd_workspace(b, r, c)
    w = torch.empty(wb // 8 + 1, **f64)
    ms = timeit(lambda: lib.nd4b_dev_svd_jac1_f64(0, st, p(a), p(u), p(sv), p(v), b, r, c, None, p(w), C.c_size_t(wb)))
    print("svd     [%6d,%3d,%3d]: %8.3f ms  %9.0f matrices/s" % (b, r, c, ms, b / ms * 1e3))
for (b, m, n, l) in [(65536, 64, 32, 1), (16384, 32, 32, 4)]:
    a, y = torch.rand(b, m, n, **f64) * 2 - 1, torch.rand(b, m, l, **f64)
    r, qy = torch.empty_like(a), torch.empty_like(y)
    ms = timeit(lambda: lib.nd4b_dev_qr_inplace_f64(0, st, p(a), p(y), p(r), p(qy), b, m, n, l))
    print("qr_inpl [%6d,%3d,%3d]+%d: %8.3f ms  %9.0f matrices/s  %6.1f GB/s" % (b, m, n, l, ms, b / ms * 1e3, 2 * (m * n + m * l) * 8 * b / ms / 1e6))
for (b, n) in [(65536, 32), (16384, 64), (262144, 8)]:
    g = torch.rand(b, n, n, **f64) * 2 - 1
    s = torch.baddbmm(float(n) * torch.eye(n, **f64).expand(b, n, n), g, g.transpose(1, 2))
    out = torch.empty_like(s)
    ms = timeit(lambda: lib.nd4b_dev_cholesky_f64(0, st, p(s), p(out), b, n, None))
    print("chol    [%6d,%3d,%3d]: %8.3f ms  %9.0f matrices/s  %6.1f GB/s" % (b, n, n, ms, b / ms * 1e3, 2 * n * n * 8 * b / ms / 1e6))
for (b, i, k, j) in [(1048576, 8, 8, 8), (262144, 16, 16, 16), (262144, 16, 32, 8), (65536, 32, 32, 32), (65536, 24, 40, 24), (16384, 64, 64, 64), (4096, 128, 128, 128),
                     (1048576, 4, 4, 4), (1048576, 3, 3, 3), (1048576, 4, 4, 1), (1048576, 3, 3, 1), (1048576, 5, 5, 5), (552336, 9, 9, 9), (154807, 17, 17, 17), (84573, 23, 23, 23)]:
    a, bb = torch.rand(b, i, k, **f64), torch.rand(b, k, j, **f64)
    c = torch.empty(b, i, j, **f64)
    ms = timeit(lambda: lib.nd4b_dev_matmul_f64(0, st, p(a), i * k, p(bb), k * j, p(c), b, i, k, j))
    print("matmul  [%7d,%3d,%3d]x[%3d,%3d]: %8.3f ms  %10.0f matrices/s  %6.1f GB/s  %7.1f GFLOP/s" % (b, i, k, k, j, ms, b / ms * 1e3, (i * k + k * j + i * j) * 8 * b / ms / 1e6, 2.0 * i * k * j * b / ms / 1e6))
