"""nd.la.* — host-side mirror of the nd4js operator interface for the batched dense-LA hot path.

Same names, argument meaning, result shapes and error texts as the reference:
  matmul2(a,b), matmul(*matrices)      src/la/matmul.js:91-147, :150-236
  cholesky_decomp(S)                   src/la/cholesky.js:50-72
  qr_decomp(A) -> (Q,R)                src/la/qr.js:80-145
  svd_jac_1sided(A) -> (U,sv,V)        contract of src/la/svd_jac_2sided.js:30-144 (new export)
  tril_solve(L,Y), triu_solve(U,Y)     src/la/tri.js:156-293
  cholesky_solve(L,y)                  src/la/cholesky.js:75-144
  qr_lstsq(Q,R,y) / qr_lstsq((Q,R),y)  src/la/qr.js:186-273
  svd_rank(sv), svd_lstsq(U,sv,V,y), svd_solve(U,sv,V,y)   src/la/svd.js:31-226 (rank cut sqrt(eps)*sv[0])
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
    """Storage of a result NDArray.  Large results live in page-locked memory from the library's cache (nd4b_host_alloc),
    so that the D2H copies of the call are DMA'd straight into them and a later nd.la call on the result is DMA'd straight
    out of them; the block goes back to the cache when the array is garbage collected."""
    return _lib.pinned_empty(int(np.prod(shape, dtype=np.int64)))


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


def _chain_plan(shapes):
    """The parenthesisation of matmul.js:159-235 for operands of the given shapes (broadcast-aware flop counts, ties broken
    like the reference: first minimum wins), as a postfix plan: i >= 0 pushes operand i, -1 multiplies the two topmost
    items.  Returns (plan, result_shape).  Host logic only."""
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

    n = len(shapes)
    op = [[None] * n for _ in range(n)]
    for i in range(n):
        op[i][i] = (0.0, [int(s) for s in shapes[i]])
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

    plan = []

    def product(lo, hi):
        if lo == hi:
            plan.append(lo)
            return
        best, idx = float("inf"), None
        for i in range(lo, hi):
            lf, ls = op[lo][i]
            rf, rs = op[i + 1][hi]
            f, _ = n_ops(ls, rs)
            f += lf + rf
            if f < best:
                best, idx = f, i
        product(lo, idx)
        product(idx + 1, hi)
        plan.append(-1)

    product(0, n - 1)
    return plan, op[0][n - 1][1]


def matmul(*matrices):
    """Matrix-chain product (matmul.js:150-236): the flop-optimal parenthesisation is found on the host exactly as the
    reference does; the products run on the GPU with every intermediate kept in HBM (nd4b_matmul_plan_f64)."""
    ms = [asarray(m) for m in matrices]
    if len(ms) == 1:
        return ms[0]
    if len(ms) == 2:
        return matmul2(*ms)
    for i, m in enumerate(ms):
        if m.ndim < 2:
            raise ValueError("A must be at least 2D." if i == 0 else "B must be at least 2D.")
    plan, shape = _chain_plan([m.shape for m in ms])
    data = [_f64(m, "matmul") for m in ms]
    n = len(ms)
    ptrs = (C.c_void_p * n)(*[d.ctypes.data for d in data])
    shapes = [np.ascontiguousarray(m.shape, dtype=np.int32) for m in ms]
    sptrs = (C.c_void_p * n)(*[sh.ctypes.data for sh in shapes])
    ndims = (C.c_int * n)(*[m.ndim for m in ms])
    plan_arr = np.asarray(plan, np.int32)
    c_shape = np.asarray(shape, np.int32)
    out = _new(c_shape)
    _lib.check(_lib.load().nd4b_matmul_plan_f64(n, ptrs, sptrs, ndims, _ptr(plan_arr), len(plan), _ptr(out), _ptr(c_shape), len(c_shape)))
    return NDArray(c_shape, out)


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


def _qr_decomp_inplace(A, Y):
    """Batched form of nd4js `_qr_decomp_inplace(M,N,L, A,A_off, Y,Y_off)` (src/la/qr.js:147-183): returns `(R, QtY)` with
    R[...,M,N] upper trapezoidal and QtY[...,M,L] = Q^T Y, without forming Q.  The reference mutates flat arrays in place;
    NDArrays are values here, so the results are new arrays (same shapes as A and Y)."""
    A, Y = asarray(A), asarray(Y)
    if A.ndim < 2 or Y.ndim < 2:
        raise ValueError("Assertion failed.")
    m, n, l = int(A.shape[-2]), int(A.shape[-1]), int(Y.shape[-1])
    ad, yd = _f64(A, "_qr_decomp_inplace"), _f64(Y, "_qr_decomp_inplace")
    if int(Y.shape[-2]) != m or tuple(A.shape[:-2]) != tuple(Y.shape[:-2]):
        raise ValueError("Assertion failed.")
    r, qty = _new(np.array(A.shape, np.int32)), _new(np.array(Y.shape, np.int32))
    _lib.check(_lib.load().nd4b_qr_inplace_f64(_ptr(ad), _ptr(yd), _ptr(r), _ptr(qty), ad.size // (m * n), m, n, l))
    return NDArray(np.array(A.shape, np.int32), r), NDArray(np.array(Y.shape, np.int32), qty)


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


def _tri_solve(op, T, Y, err_t, err_y):
    T, Y = asarray(T), asarray(Y)
    if T.ndim < 2:
        raise ValueError(err_t)
    if Y.ndim < 2:
        raise ValueError(err_y)
    td, yd = _f64(T, "tri_solve"), _f64(Y, "tri_solve")
    L = _lib.load()
    t_s = np.ascontiguousarray(T.shape, np.int32)
    y_s = np.ascontiguousarray(Y.shape, np.int32)
    ndim = max(T.ndim, Y.ndim)
    try:
        lead = np.broadcast_shapes(tuple(T.shape[:-2]), tuple(Y.shape[:-2]))
    except ValueError:
        lead = (1,) * (ndim - 2)  # let the library produce the reference's error text
    x_s = np.array(tuple(lead) + (int(Y.shape[-2]), int(Y.shape[-1])), np.int32)
    x = _new(x_s)
    rc = L.nd4b_tri_solve_f64(op, _ptr(td), _ptr(t_s), T.ndim, _ptr(yd), _ptr(y_s), Y.ndim, _ptr(x), _ptr(x_s), ndim)
    if rc in (_lib.E_INNER, _lib.E_NOT_SQUARE, _lib.E_BROADCAST, _lib.E_A_NDIM, _lib.E_B_NDIM):
        raise ValueError(_lib.last_error())
    _lib.check(rc)
    return NDArray(x_s, x)


def tril_solve(L, Y):
    return _tri_solve(0, L, Y, "tril_solve(L,Y): L.ndim must be at least 2.", "tril_solve(L,Y): Y.ndim must be at least 2.")


def triu_solve(U, Y):
    return _tri_solve(1, U, Y, "triu_solve(U,Y): U.ndim must be at least 2.", "triu_solve(U,Y): Y.ndim must be at least 2.")


def cholesky_solve(L, y):
    return _tri_solve(2, L, y, "L must be at least 2D.", "y must be at least 2D.")


def qr_lstsq(Q, R, y=None):
    """x = argmin |Q R x - y| (qr.js:186-273).  Thin factors (at most 32 columns) with one common batch run in the fused,
    bit-exact kernel (nd4b_qr_lstsq_f64); broadcast operands and wider factors are composed from matmul2 (Q^T y on the GEMM
    path: a few ulp from the reference's sequential sum, qr.js:246-249) and the bit-exact back substitution."""
    if y is None:
        y = R
        Q, R = Q
    Q, R, y = asarray(Q), asarray(R), asarray(y)
    if Q.ndim < 2:
        raise ValueError("qr_lstsq(Q,R,y): Q.ndim must be at least 2.")
    if R.ndim < 2:
        raise ValueError("qr_lstsq(Q,R,y): R.ndim must be at least 2.")
    if y.ndim < 2:
        raise ValueError("qr_lstsq(Q,R,y): y.ndim must be at least 2.")
    n, m = int(Q.shape[-2]), int(Q.shape[-1])
    i_, j_ = int(R.shape[-1]), int(y.shape[-1])
    l = min(m, i_)
    if n != int(y.shape[-2]):
        raise ValueError("qr_lstsq(Q,R,y): Q and y don't match.")
    if m != int(R.shape[-2]):
        raise ValueError("qr_lstsq(Q,R,y): Q and R don't match.")
    if i_ > n:
        raise ValueError("qr_lstsq(Q,R,y): Under-determined systems not supported. Use rrqr instead.")
    try:
        np.broadcast_shapes(tuple(Q.shape[:-2]), tuple(R.shape[:-2]), tuple(y.shape[:-2]))
    except ValueError:
        raise ValueError("Q, R, y are not broadcast-compatible.")
    if m <= 32 and i_ <= 32 and tuple(Q.shape[:-2]) == tuple(R.shape[:-2]) == tuple(y.shape[:-2]):
        qd, rd, yd = _f64(Q, "qr_lstsq"), _f64(R, "qr_lstsq"), _f64(y, "qr_lstsq")
        x_shape = np.array(tuple(Q.shape[:-2]) + (i_, j_), np.int32)
        x = _new(x_shape)
        _lib.check(_lib.load().nd4b_qr_lstsq_f64(_ptr(qd), _ptr(rd), _ptr(yd), _ptr(x), qd.size // (n * m), n, m, i_, j_))
        return NDArray(x_shape, x)
    qty = matmul2(Q.T, y).numpy()                                   # [..., m, j]
    r = R.numpy()
    x_top = triu_solve(np.ascontiguousarray(r[..., :l, :l]), np.ascontiguousarray(qty[..., :l, :])).numpy()
    if l == i_:
        return NDArray(np.asarray(x_top.shape, np.int32), x_top.reshape(-1))
    x = np.zeros(x_top.shape[:-2] + (i_, j_))
    x[..., :l, :] = x_top
    return NDArray(np.asarray(x.shape, np.int32), x.reshape(-1))


def svd_rank(sv):
    """Numerical rank per matrix: index of the first singular value <= sqrt(eps)*|sv[0]| (svd.js:31-58), computed on the
    device (nd4b_svd_rank_f64).  Result shape sv.shape[:-1] (int32); like the reference, a non-finite entry raises only
    when the scan meets it before the cut."""
    sv = asarray(sv)
    if sv.ndim < 1:
        raise ValueError("svd_rank(sv): sv.ndim must be at least 1.")
    n = int(sv.shape[-1])
    d = _f64(sv, "svd_rank")
    batch = d.size // n
    r = np.zeros(batch, np.int32)
    rc = _lib.load().nd4b_svd_rank_f64(_ptr(d), _ptr(r), batch, n)
    if rc == _lib.E_NAN_INPUT:
        raise ValueError(_lib.last_error())
    _lib.check(rc)
    return NDArray(np.array(sv.shape[:-1], np.int32), r)   # a 1-D sv gives the 0-d NDArray of the reference (shape [])


def svd_lstsq(U, sv=None, V=None, y=None):
    """x = V^T diag(1/sv[:rank]) U^T y with the reference's rank cut and broadcasting (svd.js:103-226): one fused kernel
    behind nd4b_svd_lstsq_f64, bit-identical with the reference's loops."""
    if y is None:
        if V is not None:
            raise ValueError("svd_lstsq(Q,R,P, y): Either 2 ([Q,R,P], y) or 4 arguments (Q,R,P, y) expected.")
        y = sv
        U, sv, V = U
    U, sv, V, y = asarray(U), asarray(sv), asarray(V), asarray(y)
    L = _lib.load()
    shp = [np.ascontiguousarray(a.shape, np.int32) for a in (U, sv, V, y)]
    x_s = np.zeros(max(U.ndim, sv.ndim + 1, V.ndim, y.ndim, 2), np.int32)
    nd = C.c_int(0)
    rc = L.nd4b_svd_lstsq_shape(_ptr(shp[0]), U.ndim, _ptr(shp[1]), sv.ndim, _ptr(shp[2]), V.ndim, _ptr(shp[3]), y.ndim,
                                _ptr(x_s), C.byref(nd))
    if rc:
        raise ValueError(_lib.last_error())
    x_s = np.ascontiguousarray(x_s[: nd.value])
    ud, sd, vd, yd = (_f64(a, "svd_lstsq") for a in (U, sv, V, y))
    x = _new(x_s)
    rc = L.nd4b_svd_lstsq_f64(_ptr(ud), _ptr(shp[0]), U.ndim, _ptr(sd), _ptr(shp[1]), sv.ndim, _ptr(vd), _ptr(shp[2]), V.ndim,
                              _ptr(yd), _ptr(shp[3]), y.ndim, _ptr(x), _ptr(x_s), nd.value)
    if rc == _lib.E_NAN_INPUT:
        raise ValueError(_lib.last_error())
    _lib.check(rc)
    return NDArray(x_s, x)


def svd_solve(U, sv=None, V=None, y=None):
    """svd_lstsq for square systems (svd.js:61-100).  The reference means to throw SingularMatrixSolveError for a
    rank-deficient system, but its scan `for( let r; r < N; r++ )` starts from an undefined r and never runs, so it
    returns the least-squares solution for every input; this mirror does the same."""
    if y is None:
        if V is not None:
            raise ValueError("svd_lstsq(Q,R,P, y): Either 2 ([Q,R,P], y) or 4 arguments (Q,R,P, y) expected.")
        y = sv
        U, sv, V = U
    U, sv, V = asarray(U), asarray(sv), asarray(V)
    if int(U.shape[-2]) != int(V.shape[-1]):
        raise ValueError("rrqr_solve(Q,R,P, y): System not square.")
    return svd_lstsq(U, sv, V, y)
