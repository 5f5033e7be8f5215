"""The reference's own SVD test suites, restated: the item generators of src/la/_generic_test_svd_decomp.js:180-365 driven by
the restated TestRNG (oracle/alea.py), seeded — as the reference does, src/jasmine_utils.js:277 — with the spec's description
string, plus the two checkers `test_ndarray` (:62-111) and `test_matrix` (:114-163) with the reference's tolerances.
Used for the CPU oracle (tests/test_oracle.py) and for the GPU path (tests/test_gpu_parity.py)."""
import itertools

import numpy as np

from oracle.alea import TestRNG, tabulate

EPS = 2.220446049250313e-16


def _lead(rng):
    ndim = rng.int(0, 3)
    return [rng.int(1, 4) for _ in range(ndim)]


def diagonal_batches(limit):
    """:180-201 — yields (SV expected, A) with A = diag_mat(S) of a random signed, unsorted S."""
    rng = TestRNG(" correctly decomposes random batches of diagonal matrices")
    for _ in range(min(limit, 1024)):
        n = rng.int(1, 32)
        shape = _lead(rng) + [n]
        s = tabulate(shape, lambda: rng.uniform(-4, +4) * (1.0 if rng.uniform(0, 1) < 0.9 else 0.0))
        a = np.zeros(tuple(shape) + (n,))
        idx = np.arange(n)
        a[..., idx, idx] = s
        yield -np.sort(-np.abs(s), axis=-1), a


def _tab_fn(rng, zeros):
    if zeros:
        return lambda: rng.uniform(-4, +4) * (1.0 if rng.uniform(0, 1) < 0.9 else 0.0)
    return lambda: rng.uniform(-4, +4)


def random_examples(limit, zeros):
    """:219-236"""
    rng = TestRNG(" correctly decomposes random examples" + (" with occasional zeros" if zeros else ""))
    for _ in range(min(limit, 1024)):
        shape = _lead(rng)
        shape += [rng.int(1, 32), rng.int(1, 32)]
        yield tabulate(shape, _tab_fn(rng, zeros))


def rank_deficient_examples(limit):
    """:239-254"""
    rng = TestRNG(" correctly decomposes random rank-deficient examples")
    for _ in range(min(limit, 1024)):
        shape = _lead(rng)
        shape += [rng.int(1, 32), rng.int(1, 32)]
        yield rng.rank_def(*shape)[0]


def sparse_examples(limit):
    """:257-274"""
    rng = TestRNG(" correctly decomposes random sparse examples")
    for _ in range(min(limit, 733)):
        shape = _lead(rng)
        shape += [rng.int(1, 32), rng.int(1, 32)]
        sparseness = rng.uniform(0, 1)
        yield tabulate(shape, lambda: 0.0 if sparseness > rng.uniform(0, 1) else rng.uniform(-4, +4))


def _shapes(rng, runs):
    for m in range(1, 17):
        for n in range(1, 17):
            yield m, n
    for _ in range(runs):
        m = rng.int(1, 64)
        n = rng.int(1, 64)
        yield m, n


def random_matrices(limit, zeros, skip=0):
    """:277-308 — 16 x 16 grid of small shapes, then 337 shapes up to 63 x 63"""
    rng = TestRNG("accurately decomposes random matrices" + (" with occasional zeros" if zeros else ""))
    for k, (m, n) in enumerate(_shapes(rng, 337)):
        a = tabulate([m, n], _tab_fn(rng, zeros))
        if k >= skip:
            yield a
        if k + 1 >= skip + limit:
            return


def rank_deficient_matrices(dr, dc):
    """:311-337 — sizes 1..16, then round(2^(run/3)) for run = 12..23 (16 .. 203)"""
    rng = TestRNG("accurately decomposes random rank-deficient matrices of shape [N+%d,N+%d]" % (dr, dc))
    sizes = list(range(1, 17)) + [int(np.floor(2 ** (run / 3) + 0.5)) for run in range(12, 24)]
    for l in sizes:
        yield rng.rank_def(l + dr, l + dc)[0]


def sparse_matrices(limit, skip=0):
    """:340-365"""
    rng = TestRNG("accurately decomposes random sparse matrices")
    for k, (m, n) in enumerate(_shapes(rng, 256)):
        sparseness = rng.uniform(0, 1)
        a = tabulate([m, n], lambda: 0.0 if sparseness > rng.uniform(0, 1) else rng.uniform(-4, +4))
        if k >= skip:
            yield a
        if k + 1 >= skip + limit:
            return


def _common(a, u, sv, v):
    m, n = a.shape[-2:]
    l = min(m, n)
    assert u.shape == a.shape[:-2] + (m, l) and sv.shape == a.shape[:-2] + (l,) and v.shape == a.shape[:-2] + (l, n)
    assert (sv == -np.sort(-np.abs(sv), axis=-1)).all()          # sorted descending and non-negative (:85-89)
    eye = np.eye(l)
    u_tol, v_tol = EPS * 4 * m, EPS * 4 * n
    ut, vt = np.swapaxes(u, -1, -2), np.swapaxes(v, -1, -2)
    if m >= n:
        assert np.max(np.abs(ut @ u - eye)) <= u_tol, ("U^T U", a.shape, np.max(np.abs(ut @ u - eye)), u_tol)
        assert np.max(np.abs(vt @ v - eye)) <= v_tol, ("V^T V", a.shape, np.max(np.abs(vt @ v - eye)), v_tol)
    if m <= n:
        assert np.max(np.abs(u @ ut - eye)) <= u_tol, ("U U^T", a.shape, np.max(np.abs(u @ ut - eye)), u_tol)
        assert np.max(np.abs(v @ vt - eye)) <= v_tol, ("V V^T", a.shape, np.max(np.abs(v @ vt - eye)), v_tol)
    return (u * sv[..., None, :]) @ v


def check_ndarray(a, u, sv, v):
    """test_ndarray, :62-111: shapes, order, orthogonality to 4*M*eps / 4*N*eps, U diag(sv) V close to A (atol 1e-7, rtol 1e-5)."""
    rec = _common(a, u, sv, v)
    assert (np.abs(rec - a) <= 1e-7 + 1e-5 * np.abs(a)).all()


def check_matrix(a, u, sv, v):
    """test_matrix, :114-163: as above, and |A - U diag(sv) V|_F <= 48 max(M,N) eps |A|_F."""
    assert a.ndim == 2
    rec = _common(a, u, sv, v)
    m, n = a.shape
    assert np.sqrt(np.sum((a - rec) ** 2)) <= EPS * 48 * max(m, n) * np.sqrt(np.sum(a * a)), a.shape


def check_diagonal(sv_want, a, u, sv, v):
    """:202-215 with the svd_jac* tolerance {rtol: 0, atol: 0}: everything exact."""
    n = a.shape[-1]
    eye = np.eye(n)
    ut, vt = np.swapaxes(u, -1, -2), np.swapaxes(v, -1, -2)
    assert (sv == sv_want).all()
    for g in (u @ ut, ut @ u, v @ vt, vt @ v):
        assert (g == eye).all()
    assert ((u * sv[..., None, :]) @ v == a).all()
