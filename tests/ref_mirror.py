"""A second, independent restatement of the reference's QR and two-sided Jacobi SVD — pure Python / NumPy, written from the
JavaScript sources without looking at oracle/nd4ref.c — so that the C oracle is not a single unreviewed transcription
(SURVEY.md §7.1, §8c).  tests/test_oracle.py compares the two BIT FOR BIT on seeded inputs.

Follows, statement by statement:
  _giv_rot_qr, _giv_rot_rows, _giv_rot_cols   src/la/_giv_rot.js:20-87
  _transpose_inplace                          src/la/transpose_inplace.js:22-31
  qr_decomp_full, qr_decomp                   src/la/qr.js:27-77, 80-145
  _svd_jac_angles, _svd_jac_post(_skip1)      src/la/_svd_jac_utils.js:72-114, 123-188
  svd_jac_2sided                              src/la/svd_jac_2sided.js:30-144
Arithmetic is IEEE binary64 with every product and sum rounded separately (Python floats; NumPy element-wise ufuncs do not
fuse); atan2 / cos / sin come from the platform libm, as in the C oracle.  Matrices are flat row-major Python lists of
floats or NumPy vectors, exactly like the reference's Float64Arrays.
"""
import math

import numpy as np

EPS = 2.220446049250313e-16   # Number.EPSILON, src/dt/index.js:33-43


def giv_rot_qr(a_ii, a_ji):
    mx = max(abs(a_ii), abs(a_ji))
    if mx == 0:
        return 1.0, 0.0, 0.0
    a_ii /= mx
    a_ji /= mx
    norm = math.sqrt(a_ii * a_ii + a_ji * a_ji)
    a_ii /= norm
    a_ji /= norm
    norm *= mx
    assert 0 <= norm
    return a_ii, a_ji, norm


def giv_rot_rows(w, n, i, j, c, s):
    if n <= 0:
        return
    wi, wj = w[i:i + n].copy(), w[j:j + n].copy()
    w[i:i + n] = c * wi + s * wj
    w[j:j + n] = c * wj - s * wi


def giv_rot_cols(w, n, i, j, c, s):
    wi, wj = w[i:i + n * n:n].copy(), w[j:j + n * n:n].copy()
    w[i:i + n * n:n] = c * wi - s * wj
    w[j:j + n * n:n] = c * wj + s * wi


def transpose_inplace(n, a, off):
    m = a[off:off + n * n].reshape(n, n)
    m[...] = m.T.copy()


def qr_decomp_full(a):
    a = np.asarray(a, dtype=np.float64)
    m, n = a.shape[-2:]
    blk = 8
    r = a.reshape(-1).copy()
    batch = r.size // (m * n)
    q = np.zeros(batch * m * m)
    for b in range(batch):
        q_off, r_off = b * m * m, b * m * n
        for i in range(m):
            q[q_off + m * i + i] = 1.0
        for jj_ in range(0, n, blk):
            for ii_ in range(jj_, m, blk):
                for i in range(ii_, min(ii_ + blk, m)):
                    for j in range(jj_, min(jj_ + blk, n, i)):
                        ij = r_off + n * i + j
                        r_ij = r[ij]
                        if r_ij == 0:
                            continue
                        jj = r_off + n * j + j
                        c, s, norm = giv_rot_qr(float(r[jj]), float(r_ij))
                        r[ij] = 0
                        if s == 0:
                            continue
                        r[jj] = norm
                        giv_rot_rows(r, n - 1 - j, jj + 1, ij + 1, c, s)
                        giv_rot_rows(q, 1 + i, q_off + m * j, q_off + m * i, c, s)
        transpose_inplace(m, q, q_off)
    return q.reshape(a.shape[:-2] + (m, m)), r.reshape(a.shape)


def qr_decomp(a):
    a = np.asarray(a, dtype=np.float64)
    n, m = a.shape[-2:]          # the reference calls the rows N and the columns M here
    if n <= m:
        return qr_decomp_full(a)
    q = a.reshape(-1).copy()
    batch = q.size // (n * m)
    r = np.zeros(batch * m * m)
    for b in range(batch):
        q_off, r_off = b * n * m, b * m * m
        for i in range(1, n):
            for j in range(min(i, m)):
                ij = q_off + m * i + j
                r_ij = q[ij]
                if r_ij == 0:
                    continue
                jj = q_off + m * j + j
                c, s, norm = giv_rot_qr(float(q[jj]), float(r_ij))
                if s != 0:
                    if c < 0:
                        c *= -1
                        s *= -1
                        norm *= -1
                    giv_rot_rows(q, m - 1 - j, jj + 1, ij + 1, c, s)
                    q[jj] = norm
                q[ij] = s
        for i in range(m):
            for j in range(i, m):
                r[r_off + m * i + j] = q[q_off + m * i + j]
                q[q_off + m * i + j] = 1.0 if i == j else 0.0
        for i in range(n - 1, 0, -1):
            for j in range(min(i, m) - 1, -1, -1):
                s = float(q[q_off + m * i + j])
                if s == 0:
                    continue
                q[q_off + m * i + j] = 0
                c = math.sqrt((1 - s) * (1 + s))
                giv_rot_rows(q, m - j, q_off + m * i + j, q_off + m * j + j, c, s)
    return q.reshape(a.shape), r.reshape(a.shape[:-2] + (m, m))


def svd_jac_angles(s_pp, s_pq, s_qp, s_qq):
    x = math.atan2(s_qp - s_pq, s_qq + s_pp)
    y = math.atan2(s_qp + s_pq, s_qq - s_pp)
    a, b = (x - y) / 2, (x + y) / 2
    ca, sa, cb, sb = math.cos(a), math.sin(a), math.cos(b), math.sin(b)
    x = cb * (sa * s_qp + ca * s_pp) - sb * (sa * s_qq + ca * s_pq)
    y = sb * (ca * s_qp - sa * s_pp) + cb * (ca * s_qq - sa * s_pq)
    if abs(x) < abs(y):
        sa, ca = ca, -sa
        cb, sb = sb, -cb
        x = y
    if x < 0:
        cb, sb = -cb, -sb
    return ca, sa, cb, sb


def svd_jac_post(n, u, s, v, uv_off, sv, sv_off, order):
    for i in range(n):
        sv[sv_off + i] = s[n * i + i]
    for i in range(n - 1, -1, -1):
        sv_i = float(sv[sv_off + i])
        if sv_i < 0 or (sv_i == 0 and math.copysign(1.0, sv_i) < 0):
            sv[sv_off + i] = -sv_i
            u[uv_off + n * i:uv_off + n * i + n] *= -1
    # Int32Array.prototype.sort with a comparator is a stable sort; the comparator's sign decides (NaN counts as equal)
    import functools

    def cmp(i, j):
        d = float(sv[sv_off + j]) - float(sv[sv_off + i])
        return -1 if d < 0 else (1 if d > 0 else 0)

    order[:] = sorted((int(t) for t in order), key=functools.cmp_to_key(cmp))
    for i in range(n):
        j = i
        while True:
            tmp = order[j]
            order[j] = j
            j = tmp
            if j <= i:
                break
            row_i, row_j = uv_off + order[j] * n, uv_off + j * n
            for w in (u, v):
                t = w[row_i:row_i + n].copy()
                w[row_i:row_i + n] = w[row_j:row_j + n]
                w[row_j:row_j + n] = t
            t = sv[sv_off + order[j]]
            sv[sv_off + order[j]] = sv[sv_off + j]
            sv[sv_off + j] = t
    transpose_inplace(n, u, uv_off)


def svd_jac_2sided(a):
    a = np.asarray(a, dtype=np.float64)
    n, m = a.shape[-2:]
    lead = a.shape[:-2]
    if n > m:
        q, r = qr_decomp(a)
        u, sv, v = svd_jac_2sided(r)
        return _matmul(q, u), sv, v
    if n < m:
        q, r = qr_decomp(np.ascontiguousarray(np.swapaxes(a, -1, -2)))
        u, sv, v = svd_jac_2sided(r)
        v = np.ascontiguousarray(np.swapaxes(v, -1, -2))           # transpose_inplace(V)
        return v, sv, np.ascontiguousarray(np.swapaxes(_matmul(q, u), -1, -2))
    tol = (n * EPS) ** 2
    blk = 8
    u = a.reshape(-1).copy()
    s = np.zeros(n * n)
    v = np.zeros(u.size)
    sv = np.zeros(u.size // n)
    order = list(range(n))
    if n == 1:
        for i in range(u.size):
            if u[i] < 0.0:
                u[i] *= -1.0
                sv[i] = -1.0
            else:
                sv[i] = 1.0
        return sv.reshape(a.shape), u.reshape(a.shape[:-1]), np.ones(a.shape)
    for b in range(u.size // (n * n)):
        uv_off, sv_off = b * n * n, b * n
        s[:] = u[uv_off:uv_off + n * n]
        u[uv_off:uv_off + n * n] = np.eye(n).reshape(-1)
        v[uv_off:uv_off + n * n] = np.eye(n).reshape(-1)
        finished = False
        while not finished:
            finished = True
            for q0 in range(0, n, blk):
                for p0 in range(0, q0 + 1, blk):
                    for q in range(q0, min(q0 + blk, n)):
                        for p in range(p0, min(p0 + blk, q)):
                            s_pp, s_pq, s_qp, s_qq = (float(s[n * p + p]), float(s[n * p + q]), float(s[n * q + p]), float(s[n * q + q]))
                            if not (s_pq * s_pq + s_qp * s_qp > abs(s_pp * s_qq) * tol):
                                continue
                            finished = False
                            ca, sa, cb, sb = svd_jac_angles(s_pp, s_pq, s_qp, s_qq)
                            giv_rot_rows(s, n, n * p, n * q, ca, sa)
                            giv_rot_cols(s, n, p, q, cb, sb)
                            s[n * p + q] = 0.0
                            s[n * q + p] = 0.0
                            giv_rot_rows(u, n, uv_off + n * p, uv_off + n * q, ca, sa)
                            giv_rot_rows(v, n, uv_off + n * p, uv_off + n * q, cb, -sb)
        svd_jac_post(n, u, s, v, uv_off, sv, sv_off, order)
    return u.reshape(a.shape), sv.reshape(lead + (n,)), v.reshape(a.shape)


def _matmul(a, b):
    """matmul2_RR (src/la/matmul.js:31-74) for operands with equal leading dims: C zeroed, then C[i,:] += A[i,k] * B[k,:]
    for k ascending, product and sum rounded separately."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    c = np.zeros(a.shape[:-1] + (b.shape[-1],))
    for k in range(a.shape[-1]):
        c += a[..., :, k, None] * b[..., k, None, :]
    return c
