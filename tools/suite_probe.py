"""Runs the reference's SVD suites (tests/ref_suites.py) through the GPU path and reports every item that fails, with its
sweep count — a debugging aid for tests/test_gpu_parity.py::test_reference_suite_*."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nd4js_b200 as nd  # noqa: E402
import ref_suites as rs  # noqa: E402

nd.init([0])
la = nd.la


def run(name, items, check):
    worst = 0
    for k, a in enumerate(items):
        try:
            u, sv, v = (x.numpy() for x in la.svd_jac_1sided(a))
            sw = nd.stats()["last_sweeps"]
            worst = max(worst, sw)
            check(a, u, sv, v)
        except Exception as e:  # noqa: BLE001
            print("FAIL %s item %d shape %s sweeps %s: %s" % (name, k, a.shape, nd.stats()["last_sweeps"], str(e)[:200]))
            np.save(os.path.join(ROOT, "gpurun_out", "fail_%s_%d.npy" % (name, k)), a)
    print("%s: max sweeps %d" % (name, worst))


run("rankdef_examples", rs.rank_deficient_examples(150), rs.check_ndarray)
run("sparse_examples", rs.sparse_examples(300), rs.check_ndarray)
run("sparse_matrices", rs.sparse_matrices(512), rs.check_matrix)
for dr, dc in ((0, 0), (0, 1), (1, 0)):
    run("rankdef_matrices_%d%d" % (dr, dc), rs.rank_deficient_matrices(dr, dc), rs.check_matrix)
