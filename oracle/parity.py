"""Parity metrics of the hot path against the CPU oracle — TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke() and
bench.py's post-timing parity check import it; nothing under nd4js_b200/ may).

`check(op, inputs, outputs)` runs the oracle on `inputs` (numpy arrays, a sample small enough to finish in about a second)
and returns `(max_err, bar)` in the well-posed form of north_star's tolerances:
  matmul                       max |C - Cref|_ij / (|A||B|)_ij                                 bar 1e-12
  cholesky, *_solve, lstsq     number of entries whose bits differ from the reference's          bar 0 (bit-exact)
  qr                           max of |Q-Qref|, |R-Rref|/|A|max after sign normalisation,
                               |QR-A|_F/|A|_F, |Q^T Q - I|max                                    bar 1e-12
  qr_inplace                   the same for R and the first min(M,N) rows of Q^T y              bar 1e-12
  svd                          max of |sv-svref|/sv_max (restated svd_jac_2sided), |U S V - A|_F/|A|_F,
                               |U^T U - I|max, |V V^T - I|max                                   bar 1e-12
"""
import numpy as np

from . import nd4ref

BAR = 1e-12


def _fro(x):
    return np.sqrt(np.sum(np.square(x), axis=(-2, -1)))


def matmul(a, b, c):
    ref = nd4ref.matmul2(a, b)
    den = np.abs(a) @ np.abs(b)
    den = np.where(den == 0, 1.0, den)
    return float(np.max(np.abs(c - ref) / den)), BAR


def _bits(got, ref):
    got, ref = np.ascontiguousarray(got), np.ascontiguousarray(ref)
    if got.shape != ref.shape:
        return float("inf"), 0.0
    return float(np.count_nonzero(got.view(np.int64) != ref.view(np.int64))), 0.0


def cholesky(s, l):
    return _bits(l, nd4ref.cholesky_decomp(s))


def cholesky_solve(l, y, x):
    return _bits(x, nd4ref.cholesky_solve(l, y))


def qr_lstsq(q, r, y, x):
    return _bits(x, nd4ref.qr_lstsq(q, r, y))


def svd_lstsq(u, sv, v, y, x):
    return _bits(x, nd4ref.svd_lstsq(u, sv, v, y))


def qr(a, q, r):
    qref, rref = nd4ref.qr_decomp(a)
    sg = np.where(np.diagonal(rref, axis1=-2, axis2=-1) < 0, -1.0, 1.0)
    k = sg.shape[-1]
    qn, rn = qref[..., :, :k] * sg[..., None, :], rref * sg[..., :, None]
    amax = max(float(np.max(np.abs(a))), 1e-300)
    err = max(float(np.max(np.abs(q - qn))), float(np.max(np.abs(r - rn))) / amax,
              float(np.max(_fro(q @ r - a) / np.maximum(_fro(a), 1e-300))),
              float(np.max(np.abs(np.swapaxes(q, -1, -2) @ q - np.eye(k)))))
    if (np.tril(r, -1) != 0).any() or (np.diagonal(r, axis1=-2, axis2=-1) < 0).any():
        err = float("inf")
    return err, BAR


def qr_inplace(a, y, r, qty):
    rref, qref = nd4ref.qr_decomp_inplace(a, y)
    k = min(a.shape[-2:])
    sg = np.sign(np.diagonal(rref, axis1=-2, axis2=-1))
    sg[sg == 0] = 1.0
    err = max(float(np.max(np.abs(r[..., :k, :] - sg[..., None] * rref[..., :k, :]))),
              float(np.max(np.abs(qty[..., :k, :] - sg[..., None] * qref[..., :k, :]))))
    if (np.tril(r, -1) != 0).any():
        err = float("inf")
    return err, BAR


def svd(a, u, sv, v):
    _, sref, _ = nd4ref.svd_jac_2sided(a)
    k = u.shape[-1]
    smax = np.maximum(sref[..., :1], 1e-300)
    rec = (u * sv[..., None, :]) @ v
    err = max(float(np.max(np.abs(sv - sref) / smax)),
              float(np.max(_fro(rec - a) / np.maximum(_fro(a), 1e-300))),
              float(np.max(np.abs(np.swapaxes(u, -1, -2) @ u - np.eye(k)))),
              float(np.max(np.abs(v @ np.swapaxes(v, -1, -2) - np.eye(k)))))
    if (sv < 0).any() or np.signbit(sv).any() or (np.diff(sv, axis=-1) > 0).any():
        err = float("inf")
    return err, BAR


CHECKS = {"matmul": matmul, "cholesky": cholesky, "cholesky_solve": cholesky_solve, "qr_lstsq": qr_lstsq, "svd_lstsq": svd_lstsq,
          "qr": qr, "qr_inplace": qr_inplace, "svd": svd}


def check(op, *arrays):
    return CHECKS[op](*[np.asarray(x, dtype=np.float64) for x in arrays])
