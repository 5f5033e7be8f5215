"""nd.la.* — host-side mirror of the nd4js operator interface for the batched dense-LA hot path.

Same names, argument meaning, result shapes and error texts as the reference:
  matmul2(a,b), matmul(*matrices)      src/la/matmul.js:91-147, :150-236
  cholesky_decomp(S)                   src/la/cholesky.js:50-72
  qr_decomp(A) -> (Q,R)                src/la/qr.js:80-145
  svd_jac_1sided(A) -> (U,sv,V)        contract of src/la/svd_jac_2sided.js:30-144 (new export)
Argument handling (asarray, upcasts, shape checks) stays on the host as it stays in JS in the
reference; all arithmetic happens in libnd4b.so on the GPU.  Float64 only: other result dtypes raise
(the reference's int32/float32/complex paths are outside the hot path and there is no CPU fallback).
"""
import ctypes as C

import numpy as np

from . import _lib
from .nd_array import NDArray, asarray


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def _f64(nd, who):
    if nd.dtype not in ("float64", "int32"):
        raise TypeError("%s: only float64 (or int32, upcast) input is supported on the GPU path, got %s" % (who, nd.dtype))
    d = nd.data
    return d if d.dtype == np.float64 else d.astype(np.float64)  # Float64Array.from(int32 data), qr.js:93


def _new(shape):
    return np.empty(int(np.prod(shape, dtype=np.int64)), np.float64)


def matmul2(a, b):
    a, b = asarray(a), asarray(b)
    if a.ndim < 2:
        raise ValueError("A must be at least 2D.")
    if b.ndim < 2:
        raise ValueError("B must be at least 2D.")
    L = _lib.load()
    a_s = np.ascontiguousarray(a.shape, np.int32)
    b_s = np.ascontiguousarray(b.shape, np.int32)
    c_s = np.zeros(max(a.ndim, b.ndim), np.int32)
    nd = C.c_int(0)
    rc = L.nd4b_matmul_shape(_ptr(a_s), a.ndim, _ptr(b_s), b.ndim, _ptr(c_s), C.byref(nd))
    if rc:
        raise ValueError(_lib.last_error())
    ad, bd = _f64(a, "matmul2"), _f64(b, "matmul2")
    c = _new(c_s)
    _lib.check(L.nd4b_matmul_f64(_ptr(ad), _ptr(a_s), a.ndim, _ptr(bd), _ptr(b_s), b.ndim, _ptr(c), _ptr(c_s), nd.value))
    return NDArray(c_s, c)


def matmul(*matrices):
    """Matrix-chain product; the flop-optimal parenthesisation is found on the host exactly as in
    matmul.js:159-235 (broadcast-aware flop counts), the leaves are matmul2 calls on the GPU."""
    ms = [asarray(m) for m in matrices]
    if len(ms) == 1:
        return ms[0]
    if len(ms) == 2:
        return matmul2(*ms)

    def n_ops(sa, sb):
        i, k = sa[-2], sa[-1]
        j = sb[-1]
        if sb[-2] != k:
            raise ValueError("Shape mismatch.")
        ndim = max(len(sa), len(sb))
        shape = [1] * ndim
        shape[-2], shape[-1] = i, j
        for shp in (sa, sb):
            ii, jj = ndim - 2, len(shp) - 2
            while ii > 0 and jj > 0:
                ii -= 1
                jj -= 1
                if shape[ii] == 1:
                    shape[ii] = shp[jj]
                elif shape[ii] != shp[jj] and shp[jj] != 1:
                    raise ValueError("Shapes are not broadcast-compatible.")
        return float(np.prod(shape, dtype=np.float64)) * k, shape

    n = len(ms)
    op = [[None] * n for _ in range(n)]
    for i in range(n):
        op[i][i] = (0.0, [int(s) for s in ms[i].shape])
    for length in range(2, n + 1):
        for i in range(0, n - length + 1):
            best, best_shape = float("inf"), None
            for j in range(1, length):
                lf, ls = op[i][i + j - 1]
                rf, rs = op[i + j][i + length - 1]
                f, shape = n_ops(ls, rs)
                f += lf + rf
                if f < best:
                    best, best_shape = f, shape
            if best_shape is None:
                raise OverflowError("Integer overflow (too many FLOPs).")
            op[i][i + length - 1] = (best, best_shape)

    def product(lo, hi):
        if lo == hi:
            return ms[lo]
        best, idx = float("inf"), None
        for i in range(lo, hi):
            lf, ls = op[lo][i]
            rf, rs = op[i + 1][hi]
            f, _ = n_ops(ls, rs)
            f += lf + rf
            if f < best:
                best, idx = f, i
        return matmul2(product(lo, idx), product(idx + 1, hi))

    return product(0, n - 1)


def cholesky_decomp(S):
    S = asarray(S)
    if S.ndim < 2:
        raise ValueError("Last two dimensions must be quadratic.")
    n, m = int(S.shape[-2]), int(S.shape[-1])
    if n != m:
        raise ValueError("Last two dimensions must be quadratic.")
    sd = _f64(S, "cholesky_decomp")
    out = _new(S.shape)
    bad = C.c_int64(-1)
    rc = _lib.load().nd4b_cholesky_f64(_ptr(sd), _ptr(out), sd.size // (n * n), n, C.byref(bad))
    if rc:
        raise _lib.Nd4bError(rc, _lib.last_error(), bad.value)
    return NDArray(S.shape, out)


def qr_decomp(A):
    A = asarray(A)
    if A.ndim < 2:
        raise ValueError("qr_decomp(A): A.ndim must be at least 2.")
    rows, cols = int(A.shape[-2]), int(A.shape[-1])
    l = min(rows, cols)
    ad = _f64(A, "qr_decomp")
    q_shape = np.array(A.shape, np.int32)
    q_shape[-1] = l
    r_shape = np.array(A.shape, np.int32)
    r_shape[-2] = l
    q, r = _new(q_shape), _new(r_shape)
    _lib.check(_lib.load().nd4b_qr_f64(_ptr(ad), _ptr(q), _ptr(r), ad.size // (rows * cols), rows, cols))
    return NDArray(q_shape, q), NDArray(r_shape, r)


def svd_jac_1sided(A):
    A = asarray(A)
    if A.dtype.startswith("complex"):
        raise TypeError("svd_jac_1sided(A): A.dtype must be float.")
    if A.ndim < 2:
        raise ValueError("svd_jac_1sided(A): A.ndim must be at least 2.")
    rows, cols = int(A.shape[-2]), int(A.shape[-1])
    l = min(rows, cols)
    ad = _f64(A, "svd_jac_1sided")
    u_shape = np.array(A.shape, np.int32)
    u_shape[-1] = l
    s_shape = np.array(A.shape[:-1], np.int32)
    s_shape[-1] = l
    v_shape = np.array(A.shape, np.int32)
    v_shape[-2] = l
    u, sv, v = _new(u_shape), _new(s_shape), _new(v_shape)
    sweeps = C.c_int(0)
    _lib.check(_lib.load().nd4b_svd_jac1_f64(_ptr(ad), _ptr(u), _ptr(sv), _ptr(v), ad.size // (rows * cols), rows, cols,
                                             C.byref(sweeps)))
    return NDArray(u_shape, u), NDArray(s_shape, sv), NDArray(v_shape, v)
