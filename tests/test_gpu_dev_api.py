"""Device-resident entry points (`nd4b_dev_*`, include/nd4b.h): the forms the bench times and a caller composes on its
own stream.  Inputs live in HBM (torch is only the allocator here), results are checked against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

from util import spd, uniform

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def dev(la):
    import torch
    from nd4js_b200 import _lib
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream

    class Dev:
        pass

    d = Dev()
    d.torch, d.lib, d.stream = torch, lib, C.c_void_p(stream)
    d.up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    d.p = lambda t: C.c_void_p(t.data_ptr())

    def ok(rc):
        assert rc == 0, lib.nd4b_last_error().decode()
        torch.cuda.synchronize()

    d.ok = ok
    return d


def _sign_normalise(q, r):
    sg = np.where(np.diagonal(r, axis1=-2, axis2=-1) < 0, -1.0, 1.0)
    return q * sg[..., None, :], r * sg[..., :, None]


def test_dev_matmul_strides_and_broadcast(dev, ref):
    a, b = uniform(1, (37, 32, 32)), uniform(2, (37, 32, 32))
    da, db = dev.up(a), dev.up(b)
    out = dev.torch.empty(37, 32, 32, dtype=dev.torch.float64, device="cuda")
    dev.ok(dev.lib.nd4b_dev_matmul_f64(0, dev.stream, dev.p(da), 1024, dev.p(db), 1024, dev.p(out), 37, 32, 32, 32))
    want = ref.matmul2(a, b)
    assert np.max(np.abs(out.cpu().numpy() - want) / (np.abs(a) @ np.abs(b))) <= TOL
    # stride 0: one B for the whole batch (matmul.js:59-67 with a batch dim of 1)
    dev.ok(dev.lib.nd4b_dev_matmul_f64(0, dev.stream, dev.p(da), 1024, dev.p(db), 0, dev.p(out), 37, 32, 32, 32))
    want = ref.matmul2(a, b[:1])
    assert np.max(np.abs(out.cpu().numpy() - want) / (np.abs(a) @ np.abs(b[:1]))) <= TOL
    # a general shape through the tiled kernels
    a2, b2 = uniform(3, (5, 70, 33)), uniform(4, (5, 33, 18))
    out2 = dev.torch.empty(5, 70, 18, dtype=dev.torch.float64, device="cuda")
    da2, db2 = dev.up(a2), dev.up(b2)
    dev.ok(dev.lib.nd4b_dev_matmul_f64(0, dev.stream, dev.p(da2), 70 * 33, dev.p(db2), 33 * 18, dev.p(out2), 5, 70, 33, 18))
    assert np.max(np.abs(out2.cpu().numpy() - ref.matmul2(a2, b2)) / (np.abs(a2) @ np.abs(b2))) <= TOL


def test_dev_cholesky_and_failure_key(dev, ref):
    s = spd(5, (100,), 16)
    s[41, 7, 7] = -1.0  # not positive definite: the factorisation of matrix 41 fails
    ds = dev.up(s)
    out = dev.torch.empty_like(ds)
    info = dev.torch.full((1,), 2 ** 62, dtype=dev.torch.int64, device="cuda")
    dev.ok(dev.lib.nd4b_dev_cholesky_f64(0, dev.stream, dev.p(ds), dev.p(out), 100, 16, dev.p(info)))
    assert int(info.item()) >> 1 == 41 and int(info.item()) & 1 == 1  # key = 2*index + (singular ? 1 : 0)
    good = np.delete(np.arange(100), 41)
    assert (out.cpu().numpy()[good] == ref.cholesky_decomp(s[good])).all()


@pytest.mark.parametrize("offset", [0, 1])
def test_dev_qr_64x32_aligned_and_unaligned(dev, ref, offset):
    # offset 1 shifts every pointer by 8 bytes: the blocked kernel must leave its 16-byte TMA bulk path (A, Q) for plain loads
    a = uniform(6, (33, 64, 32))
    buf_a = dev.torch.empty(a.size + 2, dtype=dev.torch.float64, device="cuda")
    buf_q = dev.torch.empty(a.size + 2, dtype=dev.torch.float64, device="cuda")
    buf_r = dev.torch.empty(33 * 32 * 32 + 2, dtype=dev.torch.float64, device="cuda")
    da, dq, dr = buf_a[offset:offset + a.size], buf_q[offset:offset + a.size], buf_r[offset:offset + 33 * 1024]
    da.copy_(dev.up(a).reshape(-1))
    assert da.data_ptr() % 16 == 8 * offset
    dev.ok(dev.lib.nd4b_dev_qr_f64(0, dev.stream, dev.p(da), dev.p(dq), dev.p(dr), 33, 64, 32, None, 0))
    q, r = dq.cpu().numpy().reshape(33, 64, 32), dr.cpu().numpy().reshape(33, 32, 32)
    qn, rn = _sign_normalise(*ref.qr_decomp(a))
    assert (np.tril(r, -1) == 0).all() and np.max(np.abs(q - qn)) <= TOL and np.max(np.abs(r - rn)) <= TOL


def test_dev_qr_generic_shapes_and_workspace(dev, ref):
    # shapes that fit in shared memory need no workspace; larger ones work in a global scratch copy the caller provides
    for (b, rows, cols) in [(9, 20, 7), (3, 300, 100)]:
        a = uniform(7, (b, rows, cols))
        need = dev.lib.nd4b_dev_qr_workspace(b, rows, cols)
        assert (need == 0) == (rows * cols < 10000) and dev.lib.nd4b_dev_qr_workspace(b, 64, 32) == 0
        da = dev.up(a)
        dq = dev.torch.empty(b, rows, cols, dtype=dev.torch.float64, device="cuda")
        dr = dev.torch.empty(b, cols, cols, dtype=dev.torch.float64, device="cuda")
        if need:
            assert dev.lib.nd4b_dev_qr_f64(0, dev.stream, dev.p(da), dev.p(dq), dev.p(dr), b, rows, cols, None, 0) != 0  # no workspace
        work = dev.torch.empty(need // 8 + 1, dtype=dev.torch.float64, device="cuda")
        dev.ok(dev.lib.nd4b_dev_qr_f64(0, dev.stream, dev.p(da), dev.p(dq), dev.p(dr), b, rows, cols, dev.p(work), C.c_size_t(need)))
        qn, rn = _sign_normalise(*ref.qr_decomp(a))
        assert np.max(np.abs(dq.cpu().numpy() - qn)) <= TOL and np.max(np.abs(dr.cpu().numpy() - rn)) <= 20 * TOL


def test_dev_qr_inplace(dev, ref):
    a, y = uniform(8, (12, 10, 4)), uniform(9, (12, 10, 3))
    dr = dev.torch.empty(12, 10, 4, dtype=dev.torch.float64, device="cuda")
    dy = dev.torch.empty(12, 10, 3, dtype=dev.torch.float64, device="cuda")
    da, dyin = dev.up(a), dev.up(y)
    dev.ok(dev.lib.nd4b_dev_qr_inplace_f64(0, dev.stream, dev.p(da), dev.p(dyin), dev.p(dr), dev.p(dy), 12, 10, 4, 3))
    r, qty = dr.cpu().numpy(), dy.cpu().numpy()
    rref, qref = ref.qr_decomp_inplace(a, y)
    sg = np.sign(np.diagonal(rref, axis1=-2, axis2=-1))
    sg[sg == 0] = 1.0
    assert (np.tril(r, -1) == 0).all()
    assert np.max(np.abs(r[:, :4] - sg[..., None] * rref[:, :4])) <= TOL
    assert np.max(np.abs(qty[:, :4] - sg[..., None] * qref[:, :4])) <= TOL


def test_dev_svd_sweeps_and_sweep_counter(dev, ref):
    a = uniform(10, (24, 64, 64))
    da = dev.up(a)
    u = dev.torch.empty(24, 64, 64, dtype=dev.torch.float64, device="cuda")
    sv = dev.torch.empty(24, 64, dtype=dev.torch.float64, device="cuda")
    v = dev.torch.empty(24, 64, 64, dtype=dev.torch.float64, device="cuda")
    sweeps = dev.torch.zeros(1, dtype=dev.torch.int32, device="cuda")
    total = dev.torch.zeros(1, dtype=dev.torch.int64, device="cuda")
    dev.ok(dev.lib.nd4b_dev_svd_sweep_counter(0, dev.p(total)))
    try:
        for _ in range(2):
            dev.ok(dev.lib.nd4b_dev_svd_jac1_f64(0, dev.stream, dev.p(da), dev.p(u), dev.p(sv), dev.p(v), 24, 64, 64, dev.p(sweeps), None, 0))
    finally:
        dev.ok(dev.lib.nd4b_dev_svd_sweep_counter(0, None))
    smax, ssum = int(sweeps.item()), int(total.item())
    assert 5 <= smax <= 15 and 2 * 24 * 5 <= ssum <= 2 * 24 * smax and ssum % 2 == 0  # two identical launches were counted
    dev.ok(dev.lib.nd4b_dev_svd_jac1_f64(0, dev.stream, dev.p(da), dev.p(u), dev.p(sv), dev.p(v), 24, 64, 64, None, None, 0))
    assert int(total.item()) == ssum  # counter switched off
    _, sref, _ = ref.svd_jac_2sided(a)
    s = sv.cpu().numpy()
    assert np.max(np.abs(s - sref)) <= TOL * sref.max()
    assert np.max(np.abs((u.cpu().numpy() * s[:, None, :]) @ v.cpu().numpy() - a)) <= 64 * TOL


def test_ieee_fast_path_helpers_match_the_compiler(dev):
    """The bit-exact Cholesky / triangular-solve kernels inline nvcc's own sqrt and division fast paths (with a range
    flag instead of a branch).  2^28 pseudo-random operands over all exponents, plus values at the range limits: every
    result the helpers mark as valid must equal sqrt() / `/` bit for bit."""
    counts = (C.c_ulonglong * 4)()
    dev.ok(dev.lib.nd4b_selfcheck_ieee(0, 1 << 28, 20261018, counts))
    bad_sqrt, bad_div, fast_sqrt, fast_div = (int(c) for c in counts)
    assert bad_sqrt == 0 and bad_div == 0, (bad_sqrt, bad_div)
    assert fast_sqrt > (1 << 27) and fast_div > (1 << 26), (fast_sqrt, fast_div)


def test_dev_svd_preconditioned_pipeline_against_the_plain_iteration(dev, ref):
    """64x64 with a workspace: FP32 Jacobi -> DMMA hand-over (V1 orthogonalised, G1 = A V1) -> FP64 Jacobi (csrc/svd_pre.cu);
    without one: the plain FP64 iteration.  Both must satisfy the contract, agree on the singular values, and the
    preconditioned run must need far fewer FP64 sweeps; a NaN matrix and a zero matrix in the batch take the hand-over's
    fallback (A, I) without disturbing their neighbours."""
    t = dev.torch
    a = uniform(21, (40, 64, 64))
    a[7] = 0.0
    a[11, :, 5] = a[11, :, 3]                     # rank deficient
    a[13] *= 2.0 ** 300
    da = dev.up(a)
    f64 = dict(dtype=t.float64, device="cuda")
    outs = []
    for with_ws in (True, False):
        u, sv, v = t.empty(40, 64, 64, **f64), t.empty(40, 64, **f64), t.empty(40, 64, 64, **f64)
        sweeps = t.zeros(1, dtype=t.int32, device="cuda")
        ws = dev.lib.nd4b_dev_svd_workspace(40, 64, 64)
        assert ws == 40 * (4096 * 4 + 2 * 4096 * 8)
        work = t.empty(ws // 8, **f64)
        dev.ok(dev.lib.nd4b_dev_svd_jac1_f64(0, dev.stream, dev.p(da), dev.p(u), dev.p(sv), dev.p(v), 40, 64, 64, dev.p(sweeps),
                                             dev.p(work) if with_ws else None, ws if with_ws else 0))
        outs.append((u.cpu().numpy(), sv.cpu().numpy(), v.cpu().numpy(), int(sweeps.item())))
    (u1, s1, v1, sw1), (u0, s0, v0, sw0) = outs
    assert sw1 <= 4 < sw0
    for u, s, v in ((u1, s1, v1), (u0, s0, v0)):
        rec = (u * s[:, None, :]) @ v
        scale = np.maximum(np.sqrt(np.sum(a * a, axis=(1, 2))), 1e-300)
        assert np.max(np.sqrt(np.sum((rec - a) ** 2, axis=(1, 2))) / scale) <= TOL
        assert np.max(np.abs(np.swapaxes(u, 1, 2) @ u - np.eye(64))) <= TOL and np.max(np.abs(v @ np.swapaxes(v, 1, 2) - np.eye(64))) <= TOL
        assert (np.diff(s, axis=-1) <= 0).all() and (s >= 0).all()
    assert np.max(np.abs(s1 - s0) / np.maximum(s0[:, :1], 1e-300)) <= TOL
    assert (s1[7] == 0).all() and s1[11, -1] <= 1e-12 * s1[11, 0]
    _, sref, _ = ref.svd_jac_2sided(a[:4])
    assert np.max(np.abs(s1[:4] - sref) / sref[:, :1]) <= TOL
    an = a.copy()
    an[3, 2, 2] = np.nan
    dan = dev.up(an)
    u, sv, v = t.empty(40, 64, 64, **f64), t.empty(40, 64, **f64), t.empty(40, 64, 64, **f64)
    work = t.empty(dev.lib.nd4b_dev_svd_workspace(40, 64, 64) // 8, **f64)
    dev.ok(dev.lib.nd4b_dev_svd_jac1_f64(0, dev.stream, dev.p(dan), dev.p(u), dev.p(sv), dev.p(v), 40, 64, 64, None, dev.p(work), work.numel() * 8))
    sn = sv.cpu().numpy()
    assert np.isnan(sn[3]).any()                                   # NaN in, NaN out for that matrix, as in the reference
    assert (np.delete(sn, 3, axis=0) == np.delete(s1, 3, axis=0)).all()   # the neighbours are untouched, bit for bit


def test_dev_matmul_bulk_copy_pipeline(dev, ref):
    """Full 128x128x32 tiles on every SM: the GEMM fed by the TMA engine (gemm_bulk_kernel: producer warp, mbarrier ring,
    cp.async.bulk per tile row)."""
    t = dev.torch
    a, b = uniform(31, (2048, 256)), uniform(32, (256, 2048))
    da, db = dev.up(a), dev.up(b)
    out = t.empty(2048, 2048, dtype=t.float64, device="cuda")
    dev.ok(dev.lib.nd4b_dev_matmul_f64(0, dev.stream, dev.p(da), 0, dev.p(db), 0, dev.p(out), 1, 2048, 256, 2048))
    rows = np.r_[0:8, 1020:1030, 2040:2048]
    want = ref.matmul2(a[rows], b)
    assert np.max(np.abs(out.cpu().numpy()[rows] - want) / (np.abs(a[rows]) @ np.abs(b))) <= TOL
    got = out.cpu().numpy()
    assert np.max(np.abs(got - a @ b)) <= 1e-11                    # every tile, against LAPACK-grade numpy


def test_dev_all_gather_one_device_context(dev):
    """nd4b_dev_all_gather_f64 with a one-device context is a device copy (no NCCL needed); the multi-device form is covered by
    tools/multidev_check.py (tests/test_gpu_multi.py) on boxes with >= 2 GPUs."""
    t = dev.torch
    if dev.lib.nd4b_device_count() != 1:
        pytest.skip("the session's context spans several devices")
    x = t.arange(1000, dtype=t.float64, device="cuda")
    y = t.zeros(1000, dtype=t.float64, device="cuda")
    shards, fulls = (C.c_void_p * 1)(x.data_ptr()), (C.c_void_p * 1)(y.data_ptr())
    counts, streams = (C.c_int64 * 1)(1000), (C.c_void_p * 1)(dev.stream.value)
    dev.ok(dev.lib.nd4b_dev_all_gather_f64(shards, counts, fulls, streams))
    assert (y == x).all()
