"""ctypes loader for the CPU oracle (oracle/libnd4ref.so).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under nd4js_b200/ may import this module.
See oracle/nd4ref.h for which nd4js file:line each entry point restates and for the
pinning status of the oracle.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libnd4ref.so")

OK, E_SINGULAR, E_NAN_INPUT = 0, 1, -7
MESSAGES = {
    -1: "A must be at least 2D.",
    -2: "B must be at least 2D.",
    -3: "The last dimension of A and the 2nd to last dimension of B do not match.",
    -4: "Shapes are not broadcast-compatible.",
    -5: "Result shape mismatch.",
    -6: "Last two dimensions must be quadratic.",
    -7: "Assertion failed.",
    1: "Matrix contains NaNs or is (near) singular.",
}


def build(force=False):
    src = os.path.join(_HERE, "nd4ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libnd4ref.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, ip, i64 = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_int64
        L.nd4ref_matmul_shape.argtypes = [ip, C.c_int, ip, C.c_int, ip, C.POINTER(C.c_int)]
        L.nd4ref_matmul_f64.argtypes = [dp, ip, C.c_int, dp, ip, C.c_int, dp, ip, C.c_int]
        L.nd4ref_cholesky_f64.argtypes = [dp, dp, i64, C.c_int, C.POINTER(i64)]
        L.nd4ref_qr_f64.argtypes = [dp, dp, dp, i64, C.c_int, C.c_int]
        L.nd4ref_qr_full_f64.argtypes = [dp, dp, dp, i64, C.c_int, C.c_int]
        L.nd4ref_qr_full_f64.restype = C.c_int
        L.nd4ref_qr_inplace_f64.argtypes = [dp, dp, dp, dp, i64, C.c_int, C.c_int, C.c_int]
        L.nd4ref_svd_jac2_f64.argtypes = [dp, dp, dp, dp, i64, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.nd4ref_tri_solve_f64.argtypes = [C.c_int, dp, ip, C.c_int, dp, ip, C.c_int, dp, ip, C.c_int]
        L.nd4ref_tri_solve_f64.restype = C.c_int
        L.nd4ref_svd_rank_f64.argtypes = [dp, ip, i64, C.c_int]
        L.nd4ref_svd_rank_f64.restype = C.c_int
        L.nd4ref_svd_lstsq_shape.argtypes = [ip, C.c_int, ip, C.c_int, ip, C.c_int, ip, C.c_int, ip, C.POINTER(C.c_int)]
        L.nd4ref_svd_lstsq_shape.restype = C.c_int
        L.nd4ref_svd_lstsq_f64.argtypes = [dp, ip, C.c_int, dp, ip, C.c_int, dp, ip, C.c_int, dp, ip, C.c_int, dp, ip, C.c_int]
        L.nd4ref_svd_lstsq_f64.restype = C.c_int
        L.nd4ref_giv_rot_qr.argtypes = [C.c_double, C.c_double, dp]
        L.nd4ref_svd_jac_angles.argtypes = [C.c_double] * 4 + [dp]
        L.nd4ref_frobenius.argtypes = [dp, i64]
        L.nd4ref_frobenius.restype = C.c_double
        for f in ("matmul_shape", "matmul_f64", "cholesky_f64", "qr_f64", "svd_jac2_f64"):
            getattr(L, "nd4ref_" + f).restype = C.c_int
        _lib = L
    return _lib


class RefError(Exception):
    def __init__(self, code, first_bad=-1):
        super().__init__(MESSAGES.get(code, "error %d" % code))
        self.code, self.first_bad = code, first_bad


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def matmul_shape(a_shape, b_shape):
    a_s, b_s = np.asarray(a_shape, np.int32), np.asarray(b_shape, np.int32)
    out = np.zeros(max(len(a_s), len(b_s), 2), np.int32)
    nd = C.c_int(0)
    rc = lib().nd4ref_matmul_shape(_ip(a_s), len(a_s), _ip(b_s), len(b_s), _ip(out), C.byref(nd))
    if rc:
        raise RefError(rc)
    return tuple(int(x) for x in out[: nd.value])


def matmul2(a, b):
    a, b = _f64(a), _f64(b)
    shape = matmul_shape(a.shape, b.shape)
    c = np.empty(shape, np.float64)
    a_s, b_s, c_s = (np.asarray(x.shape, np.int32) for x in (a, b, c))
    rc = lib().nd4ref_matmul_f64(_dp(a), _ip(a_s), a.ndim, _dp(b), _ip(b_s), b.ndim, _dp(c), _ip(c_s), c.ndim)
    if rc:
        raise RefError(rc)
    return c


def cholesky_decomp(s):
    s = _f64(s)
    if s.ndim < 2 or s.shape[-1] != s.shape[-2]:
        raise RefError(-6)
    n = s.shape[-1]
    out = np.empty_like(s)
    bad = C.c_int64(-1)
    rc = lib().nd4ref_cholesky_f64(_dp(s), _dp(out), s.size // (n * n), n, C.byref(bad))
    if rc:
        raise RefError(rc, bad.value)
    return out


def qr_decomp(a):
    a = _f64(a)
    rows, cols = a.shape[-2:]
    l = min(rows, cols)
    batch = a.size // (rows * cols)
    q = np.empty(a.shape[:-2] + (rows, l))
    r = np.empty(a.shape[:-2] + (l, cols))
    rc = lib().nd4ref_qr_f64(_dp(a), _dp(q), _dp(r), batch, rows, cols)
    if rc:
        raise RefError(rc)
    return q, r


def qr_decomp_full(a):
    """src/la/qr.js:27-77: complete QR, Q [...,rows,rows], R [...,rows,cols], any shape."""
    a = _f64(a)
    rows, cols = a.shape[-2:]
    q = np.empty(a.shape[:-2] + (rows, rows))
    r = np.empty(a.shape)
    rc = lib().nd4ref_qr_full_f64(_dp(a), _dp(q), _dp(r), a.size // (rows * cols), rows, cols)
    if rc:
        raise RefError(rc)
    return q, r


def qr_decomp_inplace(a, y):
    """_qr_decomp_inplace (src/la/qr.js:147-183) over a batch: returns (R, Q^T y)."""
    a, y = _f64(a), _f64(y)
    m, n = a.shape[-2:]
    l = y.shape[-1]
    batch = a.size // (m * n)
    assert y.shape[-2] == m and y.size // (m * l) == batch
    r, qty = np.empty_like(a), np.empty_like(y)
    rc = lib().nd4ref_qr_inplace_f64(_dp(a), _dp(y), _dp(r), _dp(qty), batch, m, n, l)
    if rc:
        raise RefError(rc)
    return r, qty


def svd_jac_2sided(a, return_sweeps=False):
    a = _f64(a)
    rows, cols = a.shape[-2:]
    l = min(rows, cols)
    batch = a.size // (rows * cols)
    u = np.empty(a.shape[:-2] + (rows, l))
    sv = np.empty(a.shape[:-2] + (l,))
    v = np.empty(a.shape[:-2] + (l, cols))
    sw = C.c_int(0)
    rc = lib().nd4ref_svd_jac2_f64(_dp(a), _dp(u), _dp(sv), _dp(v), batch, rows, cols, C.byref(sw))
    if rc:
        raise RefError(rc)
    return (u, sv, v, sw.value) if return_sweeps else (u, sv, v)


def _tri_solve(op, t, y):
    t, y = _f64(t), _f64(y)
    lead = np.broadcast_shapes(t.shape[:-2], y.shape[:-2])
    x = np.empty(lead + y.shape[-2:], np.float64)
    t_s, y_s, x_s = (np.asarray(a.shape, np.int32) for a in (t, y, x))
    rc = lib().nd4ref_tri_solve_f64(op, _dp(t), _ip(t_s), t.ndim, _dp(y), _ip(y_s), y.ndim, _dp(x), _ip(x_s), x.ndim)
    if rc:
        raise RefError(rc)
    return x


def tril_solve(l, y):
    return _tri_solve(0, l, y)


def triu_solve(u, y):
    return _tri_solve(1, u, y)


def cholesky_solve(l, y):
    return _tri_solve(2, l, y)


def qr_lstsq(q, r, y):
    """src/la/qr.js:186-273: x = triu_solve(R, Q^T y) with Q^T y accumulated k-ascending from 0, unfused."""
    q, r, y = _f64(q), _f64(r), _f64(y)
    n, m = q.shape[-2:]
    i_, j_ = r.shape[-1], y.shape[-1]
    l = min(m, i_)
    lead = np.broadcast_shapes(q.shape[:-2], r.shape[:-2], y.shape[:-2])
    qb, rb, yb = (np.broadcast_to(a, lead + a.shape[-2:]) for a in (q, r, y))
    x = np.zeros(lead + (i_, j_))
    for ix in np.ndindex(*lead):
        qy = np.zeros((l, j_))
        for k in range(n):  # x[i][j] += Q[k][i] * y[k][j], k ascending (qr.js:246-249)
            qy += qb[ix][k, :l, None] * yb[ix][k][None, :]
        x[ix][:l] = _tri_solve(1, np.ascontiguousarray(rb[ix][:l, :l]), qy)
    return x


def svd_rank(sv):
    """src/la/svd.js:31-58; result shape sv.shape[:-1] (a 1-D sv gives shape ())."""
    sv = _f64(sv)
    n = sv.shape[-1]
    r = np.zeros(sv.shape[:-1], np.int32)
    rc = lib().nd4ref_svd_rank_f64(_dp(sv), r.ctypes.data_as(C.POINTER(C.c_int32)), sv.size // n, n)
    if rc:
        raise RefError(rc)
    return r


def svd_lstsq(u, sv, v, y):
    """src/la/svd.js:103-226."""
    u, sv, v, y = _f64(u), _f64(sv), _f64(v), _f64(y)
    shp = [np.asarray(a.shape, np.int32) for a in (u, sv, v, y)]
    xs = np.zeros(max(u.ndim, sv.ndim + 1, v.ndim, y.ndim, 2), np.int32)
    nd = C.c_int(0)
    rc = lib().nd4ref_svd_lstsq_shape(_ip(shp[0]), u.ndim, _ip(shp[1]), sv.ndim, _ip(shp[2]), v.ndim, _ip(shp[3]), y.ndim, _ip(xs), C.byref(nd))
    if rc:
        raise RefError(rc)
    x = np.empty(tuple(int(t) for t in xs[: nd.value]))
    xs = np.asarray(x.shape, np.int32)
    rc = lib().nd4ref_svd_lstsq_f64(_dp(u), _ip(shp[0]), u.ndim, _dp(sv), _ip(shp[1]), sv.ndim, _dp(v), _ip(shp[2]), v.ndim,
                                    _dp(y), _ip(shp[3]), y.ndim, _dp(x), _ip(xs), x.ndim)
    if rc:
        raise RefError(rc)
    return x


def giv_rot_qr(a, b):
    out = np.empty(3)
    lib().nd4ref_giv_rot_qr(a, b, _dp(out))
    return tuple(out)


def svd_jac_angles(spp, spq, sqp, sqq):
    out = np.empty(4)
    lib().nd4ref_svd_jac_angles(spp, spq, sqp, sqq, _dp(out))
    return tuple(out)


def frobenius(x):
    x = _f64(x).ravel()
    return lib().nd4ref_frobenius(_dp(x), x.size)
