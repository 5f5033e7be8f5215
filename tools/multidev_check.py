"""One process, one context over several GPUs (nd.init([0..n-1])): every host-buffer entry point must give the same bits as
the one-device context — batches are split into contiguous shards (nd4b_api.cu run_pipeline), broadcast operands are
replicated per device, a single large matmul is split by row panels, and the cholesky failure index is the global one.
Prints one JSON line.  Usage: python tools/multidev_check.py [n_devices]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nd4js_b200 as nd  # noqa: E402
from nd4js_b200 import _lib, la  # noqa: E402


def cases():
    rng = np.random.default_rng(42)
    u = lambda *s: rng.uniform(-1, 1, s)
    g = u(3001, 16, 16)
    spd = g @ g.transpose(0, 2, 1) + 16 * np.eye(16)
    bad = spd.copy()
    bad[2500, 3, 3] = -1.0          # fails in the last shard of a 2-, 4- or 8-way split
    bad[2900, 0, 0] = np.nan
    low = np.linalg.cholesky(spd)
    a64 = u(515, 64, 64)
    out = {
        "matmul_c2": lambda: la.matmul2(cases.a2, cases.b2).numpy(),
        "matmul_c2_pinned": lambda: la.matmul2(cases.a2p, cases.b2p).numpy(),
        "matmul_broadcast": lambda: la.matmul2(cases.a2, cases.b2[:1]).numpy(),
        "matmul_lead_broadcast": lambda: la.matmul2(cases.a2.reshape(7, 1, 143, 32, 32)[:, :, :9], cases.b2[:45].reshape(1, 5, 9, 32, 32)).numpy(),
        "matmul_row_panels_1024": lambda: la.matmul2(cases.big_a, cases.big_b).numpy(),
        "matmul_row_panels_pinned": lambda: la.matmul2(cases.big_ap, cases.big_bp).numpy(),
        "cholesky": lambda: la.cholesky_decomp(spd).numpy(),
        "qr_c4": lambda: np.concatenate([t.numpy().reshape(-1) for t in la.qr_decomp(cases.a4)]),
        "qr_generic": lambda: np.concatenate([t.numpy().reshape(-1) for t in la.qr_decomp(cases.a4[:333, :40, :20])]),
        "svd_c5": lambda: np.concatenate([t.numpy().reshape(-1) for t in la.svd_jac_1sided(a64)]),
        "svd_generic": lambda: np.concatenate([t.numpy().reshape(-1) for t in la.svd_jac_1sided(a64[:77, :20, :30])]),
        "cholesky_solve": lambda: la.cholesky_solve(low, cases.y16).numpy(),
        "cholesky_solve_bcast": lambda: la.cholesky_solve(low[0], cases.y16).numpy(),
        "svd_lstsq": lambda: la.svd_lstsq(cases.u5, cases.s5, cases.v5, cases.y5).numpy(),
    }
    cases.a2, cases.b2 = u(1001, 32, 32), u(1001, 32, 32)
    cases.a2p, cases.b2p = nd.pinned_array(np.tile(cases.a2, (3, 1, 1))), nd.pinned_array(np.tile(cases.b2, (3, 1, 1)))
    cases.big_a, cases.big_b = u(1024, 768), u(768, 1030)
    cases.big_ap, cases.big_bp = nd.pinned_array(cases.big_a), nd.pinned_array(cases.big_b)
    cases.a4 = u(777, 64, 32)
    cases.y16 = u(3001, 16, 2)
    cases.u5, cases.s5, cases.v5, cases.y5 = u(301, 20, 8), np.abs(u(301, 8)) + 0.1, u(301, 8, 8), u(301, 20, 3)
    return out, bad


def first_bad(s):
    try:
        la.cholesky_decomp(s)
    except nd.Nd4bError as e:
        return e.code, e.first_bad, str(e)
    return None


def main():
    import torch
    have = torch.cuda.device_count()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else have
    if have < n or n < 2:
        print(json.dumps({"skipped": "needs %d GPUs, %d visible" % (n, have)}))
        return
    fns, bad = cases()
    nd.init([0])
    one = {k: f() for k, f in fns.items()}
    one_bad = first_bad(bad)
    _lib.check(_lib.load().nd4b_shutdown())
    nd.init(list(range(n)))
    assert nd.stats()["n_devices"] == n
    res = {"n_devices": n, "equal_bits": {}, "ms": {}}
    for k, f in fns.items():
        t0 = time.perf_counter()
        got = f()
        res["ms"][k] = round(1e3 * (time.perf_counter() - t0), 2)
        res["equal_bits"][k] = bool(got.shape == one[k].shape and (got.view(np.int64) == one[k].view(np.int64)).all())
    res["cholesky_first_bad_one_device"] = one_bad
    res["cholesky_first_bad_multi_device"] = first_bad(bad)
    res["nccl_gather_equal_bits"] = gather_check(n)
    res["all_equal"] = (all(res["equal_bits"].values()) and res["cholesky_first_bad_multi_device"] == one_bad and one_bad[1] == 2500
                        and res["nccl_gather_equal_bits"])
    print(json.dumps(res))


def gather_check(n):
    """nd4b_dev_all_gather_f64: the SVD of a batch sharded over the context's devices through the dev entry points, U / sv / V
    gathered on every device with NCCL; compared bit for bit with the one-shot host-buffer call."""
    import ctypes as C
    import torch
    L = _lib.load()
    total = 1000 + 3                                   # not divisible by the device count: unequal shards
    rng = np.random.default_rng(11)
    a = rng.uniform(-1, 1, (total, 64, 64))
    want = [t.numpy() for t in la.svd_jac_1sided(a)]
    spans = [(total * d // n, total * (d + 1) // n) for d in range(n)]
    f64 = torch.float64
    outs, fulls, streams = [], [], []
    for d, (b0, b1) in enumerate(spans):
        dev = torch.device("cuda", d)
        with torch.cuda.device(dev):
            da = torch.from_numpy(a[b0:b1]).to(dev)
            cnt = b1 - b0
            u, sv, v = (torch.empty(cnt, 64, 64, dtype=f64, device=dev), torch.empty(cnt, 64, dtype=f64, device=dev),
                        torch.empty(cnt, 64, 64, dtype=f64, device=dev))
            ws = L.nd4b_dev_svd_workspace(cnt, 64, 64)
            work = torch.empty(ws // 8 + 2, dtype=f64, device=dev)
            st = torch.cuda.current_stream(dev).cuda_stream
            rc = L.nd4b_dev_svd_jac1_f64(d, C.c_void_p(st), C.c_void_p(da.data_ptr()), C.c_void_p(u.data_ptr()), C.c_void_p(sv.data_ptr()),
                                         C.c_void_p(v.data_ptr()), cnt, 64, 64, None, C.c_void_p(work.data_ptr()), ws)
            assert rc == 0, _lib.last_error()
            outs.append((u, sv, v, work, da))
            fulls.append((torch.empty(total, 64, 64, dtype=f64, device=dev), torch.empty(total, 64, dtype=f64, device=dev),
                          torch.empty(total, 64, 64, dtype=f64, device=dev)))
            streams.append(st)
    ok = True
    ptrs = lambda seq: (C.c_void_p * n)(*[t.data_ptr() for t in seq])
    st_arr = (C.c_void_p * n)(*streams)
    for k, per in enumerate((4096, 64, 4096)):
        counts = (C.c_int64 * n)(*[(b1 - b0) * per for b0, b1 in spans])
        rc = L.nd4b_dev_all_gather_f64(ptrs([o[k] for o in outs]), counts, ptrs([f[k] for f in fulls]), st_arr)
        assert rc == 0, _lib.last_error()
    for d in range(n):
        torch.cuda.synchronize(d)
        for k in range(3):
            ok = ok and bool((fulls[d][k].cpu().numpy().view(np.int64) == want[k].view(np.int64)).all())
    return ok


if __name__ == "__main__":
    main()
