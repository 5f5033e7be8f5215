"""CPU: the oracle against everything the reference's own tests pin for the hot path,
against an independent NumPy/LAPACK computation, and against the reference's property suites."""
import numpy as np
import pytest

from util import EPS, fro, qr_sign_normalise, spd, svd_residuals, uniform


def test_matmul_known_answers(ref):
    # src/la/matmul_test.js:32-43
    c = ref.matmul2([[1], [2]], [[30, 40, 50]])
    assert c.shape == (2, 3) and (c == [[30, 40, 50], [60, 80, 100]]).all()
    # :45-62
    c = ref.matmul2([[1, 2, 3], [4, 5, 6]], [[70, 80], [90, 100], [110, 120]])
    assert (c == [[1 * 70 + 2 * 90 + 3 * 110, 1 * 80 + 2 * 100 + 3 * 120],
                  [4 * 70 + 5 * 90 + 6 * 110, 4 * 80 + 5 * 100 + 6 * 120]]).all()
    # :64-78 chain [1,4].[4,3].[3,2]
    a = [[1, 2, 3, 4]]
    b = [[11, 12, 13], [21, 22, 23], [31, 32, 33], [41, 42, 43]]
    cc = [[5, 6], [7, 8], [9, 10]]
    assert (ref.matmul2(ref.matmul2(a, b), cc) == [[6760, 7720]]).all()
    assert (ref.matmul2(a, ref.matmul2(b, cc)) == [[6760, 7720]]).all()


def test_matmul_golden_file(ref):
    g = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "known_answers.npz"))
    assert (ref.matmul2(g["mm1_a"], g["mm1_b"]) == g["mm1_c"]).all()
    assert (ref.matmul2(g["mm2_a"], g["mm2_b"]) == g["mm2_c"]).all()
    assert (ref.cholesky_decomp(g["chol_s"]) == g["chol_l"]).all()


@pytest.mark.parametrize("seed", range(40))
def test_matmul_broadcast_vs_numpy(ref, seed):
    # src/la/matmul_test.js:80-139: ndim 2-5, leading dims 1-3, matrix dims 1-15
    rng = np.random.default_rng(seed)
    nd_a, nd_b = rng.integers(0, 4, 2)
    lead = [int(x) for x in rng.integers(1, 4, max(nd_a, nd_b))]
    la_ = [d if rng.random() < 0.7 else 1 for d in lead[len(lead) - nd_a:]]
    lb_ = [d if rng.random() < 0.7 else 1 for d in lead[len(lead) - nd_b:]]
    i, k, j = (int(x) for x in rng.integers(1, 16, 3))
    a = rng.uniform(-1, 1, la_ + [i, k])
    b = rng.uniform(-1, 1, lb_ + [k, j])
    c = ref.matmul2(a, b)
    want = np.matmul(a, b)
    assert c.shape == want.shape
    np.testing.assert_allclose(c, want, rtol=1e-5, atol=1e-8)  # jasmine_utils.js:142 defaults
    assert np.max(np.abs(c - want) / np.maximum(np.abs(a) @ np.abs(b), 1e-300)) <= k * EPS


def test_matmul_errors(ref):
    with pytest.raises(ref.RefError, match="A must be at least 2D"):
        ref.matmul_shape((3,), (3, 3))
    with pytest.raises(ref.RefError, match="B must be at least 2D"):
        ref.matmul_shape((3, 3), (3,))
    with pytest.raises(ref.RefError, match="do not match"):
        ref.matmul_shape((2, 3), (4, 2))
    with pytest.raises(ref.RefError, match="broadcast-compatible"):
        ref.matmul_shape((2, 4, 3), (3, 3, 2))
    assert ref.matmul_shape((5, 1, 4, 3), (2, 3, 6)) == (5, 2, 4, 6)


def test_cholesky_docstring_example(ref):
    # src/help.js:1876-1885
    l = ref.cholesky_decomp([[25, -50], [-50, 101]])
    assert (l == [[5, 0], [-10, 1]]).all()


@pytest.mark.parametrize("n", [1, 2, 3, 7, 16, 31])
def test_cholesky_properties(ref, n):
    # src/la/cholesky_test.js:72-98: cholesky_decomp(L L^T) ~ L, lower triangular exactly
    rng = np.random.default_rng(n)
    l0 = np.tril(rng.uniform(-1, 1, (3, 2, n, n)))
    idx = np.arange(n)
    l0[..., idx, idx] = rng.uniform(1, 2, (3, 2, n))
    s = l0 @ np.swapaxes(l0, -1, -2)
    l = ref.cholesky_decomp(s)
    assert (np.triu(l, 1) == 0).all()
    np.testing.assert_allclose(l @ np.swapaxes(l, -1, -2), s, rtol=1e-5, atol=1e-8)
    # only the lower triangle is read (cholesky.js:65-67)
    s2 = s + np.triu(rng.uniform(-9, 9, s.shape), 1)
    assert (ref.cholesky_decomp(s2) == l).all()


def test_cholesky_vs_lapack_on_c3_generator(ref):
    s = spd(5, (64,), 16)
    l = ref.cholesky_decomp(s)
    want = np.linalg.cholesky(s)
    m = np.tril(np.ones((16, 16), bool))
    assert np.max(np.abs(l - want)[..., m] / np.abs(want)[..., m]) <= 1e-12


def test_cholesky_failures(ref):
    with pytest.raises(ref.RefError, match="singular") as e:
        ref.cholesky_decomp(np.stack([np.eye(3), -np.eye(3), np.eye(3)]))
    assert e.value.first_bad == 1
    s = np.eye(3)
    s[2, 0] = np.nan
    with pytest.raises(ref.RefError, match="Assertion failed"):
        ref.cholesky_decomp(s)
    # a zero pivot in the LAST row raises nothing
    assert ref.cholesky_decomp([[1.0, 0], [0, 0]])[1, 1] == 0
    with pytest.raises(ref.RefError, match="quadratic"):
        ref.cholesky_decomp(np.ones((2, 3)))


@pytest.mark.parametrize("shape", [(4, 64, 32), (3, 8, 8), (2, 5, 9), (2, 9, 1), (1, 1, 1), (2, 1, 4)])
def test_qr_properties(ref, shape):
    # src/la/qr_test.js:169-187
    a = uniform(sum(shape), shape)
    q, r = ref.qr_decomp(a)
    l = min(shape[-2:])
    assert q.shape == shape[:-1] + (l,) and r.shape == shape[:-2] + (l, shape[-1])
    assert (np.tril(r, -1) == 0).all()
    np.testing.assert_allclose(q @ r, a, atol=1e-7)
    np.testing.assert_allclose(np.swapaxes(q, -1, -2) @ q, np.broadcast_to(np.eye(l), shape[:-2] + (l, l)), atol=1e-12)
    # against LAPACK after sign normalisation
    qn, rn = qr_sign_normalise(q, r)
    for b in range(shape[0]):
        ql, rl = np.linalg.qr(a[b])
        qln, rln = qr_sign_normalise(ql, rl)
        np.testing.assert_allclose(rn[b], rln, atol=5e-14)
        if shape[-2] >= shape[-1]:
            np.testing.assert_allclose(qn[b], qln, atol=5e-14)


def test_qr_tall_leaves_mixed_signs_and_handles_zero_rows_cols(ref):
    a = uniform(1, (6, 64, 32))
    _, r = ref.qr_decomp(a)
    neg = (np.diagonal(r, axis1=-2, axis2=-1) < 0).sum()
    assert 0 < neg < 6 * 32  # qr.js:111-115
    a[:, 3, :] = 0
    a[:, :, 5] = 0
    q, r = ref.qr_decomp(a)
    np.testing.assert_allclose(q @ r, a, atol=1e-12)
    assert (np.tril(r, -1) == 0).all()


def test_giv_rot_qr(ref):
    assert ref.giv_rot_qr(0.0, 0.0) == (1.0, 0.0, 0.0)
    c, s, n = ref.giv_rot_qr(3.0, 4.0)
    assert (c, s, n) == (0.6, 0.8, 5.0)
    c, s, n = ref.giv_rot_qr(-1e200, 1e200)
    assert np.isfinite(n) and abs(c * c + s * s - 1) < 4 * EPS


@pytest.mark.parametrize("shape", [(2, 64, 64), (3, 16, 16), (3, 9, 5), (3, 5, 9), (4, 2, 2), (2, 1, 1)])
def test_svd_jac2_properties(ref, shape):
    # src/la/_generic_test_svd_decomp.js:79-163
    a = uniform(7 + sum(shape), shape)
    u, sv, v = ref.svd_jac_2sided(a)
    m, n = shape[-2:]
    l = min(m, n)
    assert u.shape == shape[:-1] + (l,) and sv.shape == shape[:-2] + (l,) and v.shape == shape[:-2] + (l, n)
    assert (sv >= 0).all() and (np.diff(sv, axis=-1) <= 0).all()
    recon, ou, ov = svd_residuals(a, u, sv, v)
    assert ou <= 4 * max(m, n) * EPS and ov <= 4 * max(m, n) * EPS
    rec = (u * sv[..., None, :]) @ v
    assert (fro(rec - a) <= 48 * max(m, n) * EPS * fro(a)).all()
    np.testing.assert_allclose(sv, np.linalg.svd(a, compute_uv=False), rtol=0, atol=1e-13 * max(1.0, float(sv.max())))


def test_svd_jac2_diagonal_exact(ref):
    # _generic_test_svd_decomp.js:180-216: exact for svd_jac* (rtol = atol = 0)
    rng = np.random.default_rng(3)
    for n in (1, 2, 5, 17, 32):
        d = rng.uniform(-4, 4, (3, n)) * (rng.uniform(0, 1, (3, n)) < 0.9)
        a = np.zeros((3, n, n))
        a[:, np.arange(n), np.arange(n)] = d
        u, sv, v = ref.svd_jac_2sided(a)
        assert (sv == -np.sort(-np.abs(d), axis=-1)).all()
        eye = np.broadcast_to(np.eye(n), (3, n, n))
        assert (u @ np.swapaxes(u, -1, -2) == eye).all() and (v @ np.swapaxes(v, -1, -2) == eye).all()
        assert ((u * sv[:, None, :]) @ v == a).all()


def test_svd_hand_crafted_int_example(ref):
    # _generic_test_svd_decomp.js:167-177
    a = np.array([[1, 1], [1, 2], [1, 3], [1, 4], [1, 5]], float)
    u, sv, v = ref.svd_jac_2sided(a)
    recon, ou, ov = svd_residuals(a, u, sv, v)
    assert recon <= 48 * 5 * EPS and ou <= 20 * EPS and ov <= 20 * EPS


def test_svd_jac_angles_diagonalise(ref):
    rng = np.random.default_rng(0)
    for _ in range(100):
        spp, spq, sqp, sqq = rng.uniform(-1, 1, 4)
        ca, sa, cb, sb = ref.svd_jac_angles(spp, spq, sqp, sqq)
        ra = np.array([[ca, sa], [-sa, ca]])
        rb = np.array([[cb, sb], [-sb, cb]])
        d = ra @ np.array([[spp, spq], [sqp, sqq]]) @ rb
        assert abs(d[0, 1]) < 1e-15 and abs(d[1, 0]) < 1e-15
        assert d[0, 0] >= abs(d[1, 1]) - 1e-15  # s1 >= |s2|, s1 >= 0


def test_frobenius(ref):
    x = uniform(0, (1000,))
    assert abs(ref.frobenius(x) - np.linalg.norm(x)) <= 1e-13 * np.linalg.norm(x)
    assert ref.frobenius(np.zeros(4)) == 0.0


@pytest.mark.parametrize("op", ["tril_solve", "triu_solve", "cholesky_solve"])
def test_tri_solves_vs_numpy(ref, op):
    # cholesky_test.js:33-69 (2048 broadcast cholesky_solve cases) in miniature, plus tri.js public solves
    rng = np.random.default_rng(len(op))
    for _ in range(40):
        m, j = (int(v) for v in rng.integers(1, 12, 2))
        lead_t = [int(v) for v in rng.integers(1, 4, rng.integers(0, 3))]
        lead_y = [d if rng.random() < 0.6 else 1 for d in lead_t][-int(rng.integers(0, 3)):] if lead_t else []
        t = rng.uniform(-1, 1, lead_t + [m, m]) + 4 * np.eye(m)
        t = np.tril(t) if op != "triu_solve" else np.triu(t)
        y = rng.uniform(-1, 1, lead_y + [m, j])
        x = getattr(ref, op)(t, y)
        full = t @ np.swapaxes(t, -1, -2) if op == "cholesky_solve" else t
        np.testing.assert_allclose(full @ x, np.broadcast_to(y, x.shape), rtol=1e-9, atol=1e-11)


def test_qr_lstsq_vs_lapack(ref):
    rng = np.random.default_rng(11)
    a, y = rng.uniform(-1, 1, (3, 12, 5)), rng.uniform(-1, 1, (3, 12, 2))
    q, r = ref.qr_decomp(a)
    x = ref.qr_lstsq(q, r, y)
    for b in range(3):
        np.testing.assert_allclose(x[b], np.linalg.lstsq(a[b], y[b], rcond=None)[0], atol=1e-13)


@pytest.mark.parametrize("shape", [(3, 7, 4, 2), (2, 4, 4, 3), (2, 3, 6, 1), (1, 1, 1, 1)])
def test_qr_decomp_inplace_restatement(ref, shape):
    # src/la/qr.js:147-183: Givens rotations are orthogonal, so (Q^T y) and R solve the same least-squares problem as (A, y):
    # R upper trapezoidal with exact zeros, R^T R = A^T A, R^T (Q^T y) = A^T y, |Q^T y|_F = |y|_F, and the thin factor
    # agrees with qr_decomp's R up to row signs
    b, m, n, l = shape
    a, y = uniform(31, (b, m, n)), uniform(32, (b, m, l))
    r, qty = ref.qr_decomp_inplace(a, y)
    assert r.shape == a.shape and qty.shape == y.shape
    assert (np.tril(r, -1) == 0).all()
    at = np.swapaxes(a, -1, -2)
    rt = np.swapaxes(r, -1, -2)
    np.testing.assert_allclose(rt @ r, at @ a, atol=1e-13)
    np.testing.assert_allclose(rt @ qty, at @ y, atol=1e-13)
    np.testing.assert_allclose(np.linalg.norm(qty, axis=(-2, -1)), np.linalg.norm(y, axis=(-2, -1)), rtol=1e-14)
    k = min(m, n)
    _, r2 = ref.qr_decomp(a)
    np.testing.assert_allclose(np.abs(r[..., :k, :]), np.abs(r2), atol=1e-13)


def test_svd_rank_and_lstsq_restatement(ref):
    """src/la/svd.js:31-58 and :103-226 against LAPACK (minimum-norm least squares with the same rank cut), incl. broadcasting."""
    rng = np.random.default_rng(5)
    a, y = rng.uniform(-1, 1, (4, 9, 6)), rng.uniform(-1, 1, (4, 9, 2))
    a[1, :, 5] = a[1, :, 0]                       # rank 5
    u, s, vt = np.linalg.svd(a, full_matrices=False)
    assert list(ref.svd_rank(s)) == [6, 5, 6, 6]
    x = ref.svd_lstsq(u, s, vt, y)
    for b in range(4):
        np.testing.assert_allclose(x[b], np.linalg.lstsq(a[b], y[b], rcond=1e-8)[0], atol=1e-12)
    # operands broadcast independently: one factorisation, many right-hand sides and the other way round
    yb = rng.uniform(-1, 1, (3, 1, 9, 2))
    xb = ref.svd_lstsq(u, s, vt, yb)
    assert xb.shape == (3, 4, 6, 2)
    np.testing.assert_array_equal(xb[2, 1], ref.svd_lstsq(u[1], s[1], vt[1], yb[2, 0]))
    assert ref.svd_rank(np.array([3.0, 1e-9, np.nan])).shape == () and int(ref.svd_rank(np.array([3.0, 1e-9, np.nan]))) == 1
    with pytest.raises(ref.RefError):
        ref.svd_rank(np.array([3.0, np.inf, 1.0]))


# ------------------------------------------------------------ second mirror, seeded reference generators ----

def test_alea_known_answers():
    """The restated AleaRNG (src/rand/alea_rng.js:37-142) against the published vectors of the Alea generator it is
    (seedrandom's alea('hello.'): next, 53-bit double, int32)."""
    from oracle.alea import DIV53, AleaRNG, _i32, mash
    r = AleaRNG("hello.")
    assert r._next() == 0.4783254903741181
    assert r._next() + _i32(r._next() * 0x200000) * DIV53 == 0.8297006866124559
    assert _i32(r._next() * 4294967296.0) == 1076136327
    assert mash(" ", 0xefc8249d) == mash(" ", 0xefc8249d) and mash("ab", 1.0) != mash("ba", 1.0)
    u = AleaRNG(7)
    xs = [u.uniform(-4, 4) for _ in range(1000)]
    assert -4 <= min(xs) and max(xs) < 4 and abs(sum(xs) / 1000) < 0.4
    ks = [u.int(1, 4) for _ in range(300)]
    assert set(ks) == {1, 2, 3}
    q = u.ortho(3, 7, 4)
    assert np.max(np.abs(np.swapaxes(q, -1, -2) @ q - np.eye(4))) < 1e-14


@pytest.mark.parametrize("shape", [(3, 9, 5), (2, 5, 9), (4, 6, 6), (2, 1, 1), (1, 17, 17), (2, 20, 3), (1, 12, 30), (1, 64, 32),
                                   (1, 1, 7), (1, 7, 1), (2, 3, 8, 8)])
def test_numpy_mirror_of_qr_is_bit_identical_with_the_c_oracle(ref, shape):
    import ref_mirror
    a = uniform(900 + sum(shape), shape)
    if shape[-1] > 2:
        a.reshape(-1, *shape[-2:])[0, :, 1] = 0.0       # a zero column: the `continue` branches
    q1, r1 = ref_mirror.qr_decomp(a)
    q2, r2 = ref.qr_decomp(a)
    assert q1.shape == q2.shape and r1.shape == r2.shape
    assert (q1.view(np.int64) == q2.view(np.int64)).all() and (r1.view(np.int64) == r2.view(np.int64)).all()


@pytest.mark.parametrize("shape", [(3, 6, 6), (2, 9, 5), (2, 5, 9), (2, 1, 1), (1, 17, 17), (1, 2, 2), (1, 1, 4), (1, 4, 1), (1, 33, 20),
                                   (1, 40, 40)])
def test_numpy_mirror_of_svd_jac_2sided_is_bit_identical_with_the_c_oracle(ref, shape):
    import ref_mirror
    a = uniform(950 + sum(shape), shape)
    if min(shape[-2:]) > 2:
        a[0, :, 1] = a[0, :, 0]                           # rank deficient: zero singular values, sign flips of -0
    got = ref_mirror.svd_jac_2sided(a)
    want = ref.svd_jac_2sided(a)
    for g, w in zip(got, want):
        assert g.shape == w.shape and (np.ascontiguousarray(g).view(np.int64) == np.ascontiguousarray(w).view(np.int64)).all()


def test_oracle_passes_the_reference_svd_suites_on_the_reference_inputs(ref):
    """The restated svd_jac_2sided on the items the reference's own suites generate (same seeds, same generators), held to
    the reference's own tolerances — a bounded number of items per suite so that the CPU suite stays short."""
    import ref_suites as rs
    for sv_want, a in rs.diagonal_batches(60):
        rs.check_diagonal(sv_want, a, *ref.svd_jac_2sided(a))
    for zeros in (False, True):
        for a in rs.random_examples(40, zeros):
            rs.check_ndarray(a, *ref.svd_jac_2sided(a))
        for a in rs.random_matrices(40, zeros, skip=250):
            rs.check_matrix(a, *ref.svd_jac_2sided(a))
    for a in rs.rank_deficient_examples(30):
        rs.check_ndarray(a, *ref.svd_jac_2sided(a))
    for a in rs.sparse_examples(40):
        rs.check_ndarray(a, *ref.svd_jac_2sided(a))
    for a in rs.sparse_matrices(30, skip=250):
        rs.check_matrix(a, *ref.svd_jac_2sided(a))
    for k, a in enumerate(rs.rank_deficient_matrices(1, 0)):
        if k < 20:
            rs.check_matrix(a, *ref.svd_jac_2sided(a))
