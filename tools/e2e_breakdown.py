"""Where the time of one nd.la.matmul2 call on C2 goes: the raw C ABI on caller-owned pinned buffers, the same with a fresh
pooled result block per call, and the operator (nd4js_b200.la.matmul2).  Prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nd4js_b200 as nd  # noqa: E402
from nd4js_b200 import _lib, la  # noqa: E402

nd.init([0])
L = nd.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
rng = np.random.default_rng(0)
a, b = nd.pinned_array(rng.uniform(-1, 1, (n, 32, 32))), nd.pinned_array(rng.uniform(-1, 1, (n, 32, 32)))
shp = np.asarray((n, 32, 32), np.int32)
sp = C.c_void_p(shp.ctypes.data)
p = lambda x: C.c_void_p(x.ctypes.data)
out = {}


def timed(name, fn, reps=8):
    fn()
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    out[name] = round(1e3 * (time.perf_counter() - t0) / reps, 3)


c_fixed = nd.pinned_empty(n * 1024)
timed("abi_fixed_output_ms", lambda: L.nd4b_matmul_f64(p(a.data), sp, 3, p(b.data), sp, 3, p(c_fixed), sp, 3))


def fresh():
    c = nd.pinned_empty(n * 1024)
    L.nd4b_matmul_f64(p(a.data), sp, 3, p(b.data), sp, 3, p(c), sp, 3)
    return c


timed("abi_pooled_output_ms", fresh)
timed("pinned_empty_only_ms", lambda: nd.pinned_empty(n * 1024))
timed("operator_ms", lambda: la.matmul2(a, b))
keep = []
timed("operator_keep_result_ms", lambda: keep.append(la.matmul2(a, b)) or (len(keep) > 1 and keep.pop(0)))
print(json.dumps(out))
