"""One launch each of the shape-generic kernels added late in round 1 (for ncu): python tools/new_kernels_case.py"""
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
for (b, i, k, j) in [(1048576, 4, 4, 4), (1048576, 8, 8, 8)]:
    a, bb, c = torch.rand(b, i, k, **f64), torch.rand(b, k, j, **f64), torch.empty(b, i, j, **f64)
    lib.nd4b_dev_matmul_f64(0, st, p(a), i * k, p(bb), k * j, p(c), b, i, k, j)
b, n = 65536, 32
g = torch.rand(b, n, n, **f64) * 2 - 1
s = torch.baddbmm(float(n) * torch.eye(n, **f64).expand(b, n, n), g, g.transpose(1, 2))
out = torch.empty_like(s)
lib.nd4b_dev_cholesky_f64(0, st, p(s), p(out), b, n, None)
t = torch.rand(b, n, n, **f64) + 4 * torch.eye(n, **f64)
y = torch.rand(b, n, 1, **f64)
x = torch.empty_like(y)
lib.nd4b_dev_tri_solve_f64(0, st, 2, p(t), n * n, p(y), n, p(x), b, n, 1)
a = torch.rand(16384, 32, 32, **f64) * 2 - 1
q, r = torch.empty(16384, 32, 32, **f64), torch.empty(16384, 32, 32, **f64)
lib.nd4b_dev_qr_f64(0, st, p(a), p(q), p(r), 16384, 32, 32, None, C.c_size_t(0))
torch.cuda.synchronize()
print("done")
