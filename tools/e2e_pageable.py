"""End-to-end timing of the host-buffer path from PAGEABLE numpy memory (what nd.la.* sees from an ordinary
Float64Array) vs pinned memory, C2 workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import nd4js_b200 as nd
nd.init([0])
n = 65536
rng = np.random.default_rng(0)
a = rng.uniform(-1, 1, (n, 32, 32)); b = rng.uniform(-1, 1, (n, 32, 32))
for rep in range(3):
    t0 = time.perf_counter(); c = nd.la.matmul2(a, b); dt = time.perf_counter() - t0
    print("pageable C2 matmul2: %.1f ms  (%.2f M matrices/s, %.1f GB/s moved)" % (1e3 * dt, n / dt / 1e6, 1.5 * 2**30 / dt / 1e9))
assert np.abs(c.numpy()[::997] - a[::997] @ b[::997]).max() < 1e-12
g = rng.uniform(-1, 1, (262144, 16, 16)); s = g @ g.transpose(0, 2, 1) + 16 * np.eye(16)
for rep in range(2):
    t0 = time.perf_counter(); l = nd.la.cholesky_decomp(s); dt = time.perf_counter() - t0
    print("pageable C3 cholesky: %.1f ms" % (1e3 * dt))
print(nd.stats())
