"""Shared generators and parity metrics."""
import numpy as np

EPS = 2.220446049250313e-16


def uniform(seed, shape):
    return np.random.default_rng(seed).uniform(-1.0, 1.0, shape)


def spd(seed, batch_shape, n, shift=None):
    """S = G G^T + shift*I with G ~ U(-1,1): the C3 generator (shift = n keeps cond <~ 10)."""
    g = uniform(seed, tuple(batch_shape) + (n, n))
    return g @ np.swapaxes(g, -1, -2) + (n if shift is None else shift) * np.eye(n)


def qr_sign_normalise(q, r):
    """Row i of R and column i of Q times sign(R_ii) (the reference's tall branch leaves mixed signs)."""
    d = np.diagonal(r, axis1=-2, axis2=-1)
    s = np.where(d < 0, -1.0, 1.0)
    k = s.shape[-1]
    return q[..., :, :k] * s[..., None, :], r * s[..., :, None]


def matmul_componentwise_err(c, c_ref, a, b):
    """max |C - C_ref|_ij / (|A||B|)_ij — the well-posed form of 'entrywise 1e-12 relative'."""
    den = np.abs(a) @ np.abs(b)
    den = np.where(den == 0, 1.0, den)
    return float(np.max(np.abs(c - c_ref) / den))


def fro(x):
    return np.sqrt(np.sum(np.square(x), axis=(-2, -1)))


def svd_residuals(a, u, sv, v):
    rec = (u * sv[..., None, :]) @ v
    recon = float(np.max(fro(rec - a) / np.maximum(fro(a), 1e-300)))
    k = u.shape[-1]
    orth_u = float(np.max(np.abs(np.swapaxes(u, -1, -2) @ u - np.eye(k))))
    orth_v = float(np.max(np.abs(v @ np.swapaxes(v, -1, -2) - np.eye(k))))
    return recon, orth_u, orth_v
