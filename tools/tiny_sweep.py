"""Device-resident timing of QR / SVD / Cholesky on tiny and small matrices: python tools/tiny_sweep.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import torch, nd4js_b200 as nd
nd.init([0]); lib = nd.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
f64 = dict(dtype=torch.float64, device="cuda")
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n in (3, 4, 8, 12, 16, 24):
    b = 262144 if n <= 8 else 8192
    a = torch.rand(b, n, n, **f64) * 2 - 1
    q, r = torch.empty(b, n, n, **f64), torch.empty(b, n, n, **f64)
    ms = timeit(lambda: lib.nd4b_dev_qr_f64(0, st, p(a), p(q), p(r), b, n, n, None, C.c_size_t(0)))
    print("qr   %dx%d x%d: %.3f ms  %.1f ns/matrix" % (n, n, b, ms, ms * 1e6 / b))
    u, sv, v = torch.empty(b, n, n, **f64), torch.empty(b, n, **f64), torch.empty(b, n, n, **f64)
    ms = timeit(lambda: lib.nd4b_dev_svd_jac1_f64(0, st, p(a), p(u), p(sv), p(v), b, n, n, None, None, C.c_size_t(0)))
    print("svd  %dx%d x%d: %.3f ms  %.1f ns/matrix" % (n, n, b, ms, ms * 1e6 / b))
    s = torch.baddbmm(float(n) * torch.eye(n, **f64).expand(b, n, n), a, a.transpose(1, 2)); out = torch.empty_like(s)
    ms = timeit(lambda: lib.nd4b_dev_cholesky_f64(0, st, p(s), p(out), b, n, None))
    print("chol %dx%d x%d: %.3f ms  %.1f ns/matrix" % (n, n, b, ms, ms * 1e6 / b))
