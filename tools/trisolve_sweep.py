"""Device-resident timing of the 16x16 triangular solves for 1-8 right-hand sides: python tools/trisolve_sweep.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.getcwd())
import torch, nd4js_b200 as nd
nd.init([0]); lib = nd.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
f64 = dict(dtype=torch.float64, device="cuda")
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
b = 262144
for op in (0, 1, 2):
  for J in (1, 2, 4, 8):
    t = torch.rand(b, 16, 16, **f64) + 4 * torch.eye(16, **f64)
    y = torch.rand(b, 16, J, **f64); x = torch.empty_like(y)
    ms = timeit(lambda: lib.nd4b_dev_tri_solve_f64(0, st, op, p(t), 256, p(y), 16 * J, p(x), b, 16, J))
    print("op %d J %d: %.4f ms  %.0f GB/s" % (op, J, ms, b * (2048 + 256 * J) / ms / 1e6))
print("generic sizes (cholesky_solve, op 2):")
for (b, m, J) in [(262144, 8, 1), (1048576, 4, 1), (1048576, 3, 1), (65536, 32, 1), (65536, 32, 4), (16384, 64, 1), (262144, 12, 2)]:
    t = torch.rand(b, m, m, **f64) + 4 * torch.eye(m, **f64)
    y = torch.rand(b, m, J, **f64); x = torch.empty_like(y)
    ms = timeit(lambda: lib.nd4b_dev_tri_solve_f64(0, st, 2, p(t), m * m, p(y), m * J, p(x), b, m, J))
    print("M %2d J %d batch %7d: %.4f ms  %.0f GB/s" % (m, J, b, ms, b * (m * m + 2 * m * J) * 8 / ms / 1e6))
