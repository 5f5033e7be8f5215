"""Seeded inputs for the reference-run golden vectors (TEST INFRASTRUCTURE).

Each case: (name, op, [numpy inputs]).  `run_js` evaluates the case with the reference's own functions
(src/la/*.js executed by QJSEngine, see qjs.py) and returns the outputs, or the thrown message for the
failure cases.  The same list feeds `gen_golden.py` (fixed seed -> tests/golden/jsref_golden.npz) and the
live test (fresh seeds, only where the reference checkout and the engine exist).
"""
import numpy as np

from . import qjs


def _spd(rng, lead, n, shift=None):
    a = rng.standard_normal(lead + (n, n))
    return a @ np.swapaxes(a, -1, -2) + (n if shift is None else shift) * np.eye(n)


def _tri(rng, lead, n, lower, zero_other=True):
    t = rng.standard_normal(lead + (n, n)) + 3.0 * np.eye(n)
    return np.tril(t) if lower else np.triu(t)


def make_cases(seed, small=False):
    """small=True keeps the matrices tiny (live test, pure-JS interpreter speed)."""
    rng = np.random.default_rng(seed)
    N = rng.standard_normal
    cases = []
    add = lambda name, op, *ins: cases.append((name, op, [np.ascontiguousarray(x, dtype=np.float64) for x in ins]))

    # --- matmul2 (src/la/matmul.js:91-147): shapes, broadcasting, the BASELINE tile sizes
    add("mm_2d", "matmul2", N((3, 4)), N((4, 5)))
    add("mm_batch", "matmul2", N((2, 3, 4)), N((2, 4, 2)))
    add("mm_bcast_b", "matmul2", N((5, 6, 7)), N((7, 3)))
    add("mm_bcast_both", "matmul2", N((2, 1, 3, 3)), N((1, 4, 3, 2)))
    add("mm_lead_bcast", "matmul2", N((3, 1, 2, 5)), N((4, 5, 2)))
    add("mm_outer", "matmul2", N((7, 1)), N((1, 7)))
    add("mm_ints", "matmul2", rng.integers(-9, 10, (2, 4, 6)), rng.integers(-9, 10, (2, 6, 3)))
    add("mm_cancel", "matmul2", N((2, 8, 8)) * 1e8, N((2, 8, 8)) * 1e-8)
    if not small:
        add("mm_c2_tile", "matmul2", N((6, 32, 32)), N((6, 32, 32)))
        add("mm_c2_bcast", "matmul2", N((6, 32, 32)), N((1, 32, 32)))
        add("mm_c1_small", "matmul2", N((96, 80)), N((80, 72)))
    # --- matmul(...) chain ordering (matmul.js:150-236)
    add("chain3", "matmul", N((6, 2)), N((2, 9)), N((9, 3)))
    add("chain4", "matmul", N((2, 7, 3)), N((3, 8)), N((1, 8, 2)), N((2, 2, 5)))
    add("chain5_ties", "matmul", N((4, 4)), N((4, 4)), N((4, 4)), N((4, 4)), N((4, 4)))
    add("chain_vec_end", "matmul", N((5, 9)), N((9, 9)), N((9, 1)))

    # --- cholesky (src/la/cholesky.js:27-72, src/kahan_sum.js)
    add("chol_1", "cholesky_decomp", _spd(rng, (3,), 1))
    add("chol_5", "cholesky_decomp", _spd(rng, (2, 2), 5))
    add("chol_16", "cholesky_decomp", _spd(rng, (6,), 16))
    add("chol_16_illcond", "cholesky_decomp", _spd(rng, (4,), 16, shift=1e-9))
    s = _spd(rng, (3,), 16)
    s = np.triu(s) * 0 + np.tril(s) + np.triu(N((3, 16, 16)), 1)      # garbage above the diagonal is never read
    add("chol_16_lower_only", "cholesky_decomp", s)
    sp = _spd(rng, (3,), 16) * (rng.random((3, 16, 16)) < 0.3)
    sp = np.tril(sp) + np.swapaxes(np.tril(sp, -1), -1, -2) + 40 * np.eye(16)
    add("chol_16_sparse", "cholesky_decomp", sp)
    add("chol_16_scaled", "cholesky_decomp", np.stack([_spd(rng, (), 16) * 1e300, _spd(rng, (), 16) * 1e-300]))
    if not small:
        add("chol_33", "cholesky_decomp", _spd(rng, (2,), 33))
        add("chol_64", "cholesky_decomp", _spd(rng, (1,), 64))
    bad = _spd(rng, (4,), 16)
    bad[2, 7, 7] = -1.0
    add("chol_fail_not_pd", "cholesky_decomp", bad)
    bad = _spd(rng, (4,), 16)
    bad[1, 5, 3] = np.nan
    add("chol_fail_nan", "cholesky_decomp", bad)
    add("chol_fail_shape", "cholesky_decomp", N((2, 3, 4)))

    # --- triangular solves (src/la/tri.js:45-292, cholesky.js:75-140)
    add("tril_16", "tril_solve", _tri(rng, (5,), 16, True), N((5, 16, 3)))
    add("triu_16", "triu_solve", _tri(rng, (5,), 16, False), N((5, 16, 1)))
    add("tril_bcast", "tril_solve", _tri(rng, (), 7, True), N((3, 7, 2)))
    add("tril_fail_bcast", "tril_solve", _tri(rng, (7,), 7, True), N((3, 7, 2)))
    add("triu_fail_rows", "triu_solve", _tri(rng, (2,), 6, False), N((2, 5, 1)))
    add("triu_bcast", "triu_solve", _tri(rng, (3, 1), 6, False), N((1, 4, 6, 2)))
    add("tril_garbage_upper", "tril_solve", _tri(rng, (2,), 9, True) + np.triu(N((2, 9, 9)), 1), N((2, 9, 2)))
    lch = np.linalg.cholesky(_spd(rng, (5,), 16))
    add("chol_solve_16", "cholesky_solve", lch, N((5, 16, 1)))
    add("chol_solve_16_4rhs", "cholesky_solve", lch, N((5, 16, 4)))
    add("chol_solve_unit_rhs", "cholesky_solve", lch[:2], np.broadcast_to(np.eye(16)[:, :2], (2, 16, 2)))
    if not small:
        add("chol_solve_40", "cholesky_solve", np.linalg.cholesky(_spd(rng, (2,), 40)), N((2, 40, 2)))

    # --- QR (src/la/qr.js:27-145, _giv_rot.js)
    add("qr_tall", "qr_decomp", N((3, 7, 4)))
    add("qr_square", "qr_decomp", N((3, 5, 5)))
    add("qr_wide", "qr_decomp", N((2, 4, 7)))
    add("qr_1x1", "qr_decomp", N((3, 1, 1)))
    add("qr_col", "qr_decomp", N((2, 6, 1)))
    z = N((3, 8, 4))
    z[0, :, 2] = 0.0
    z[1, :, 1] = z[1, :, 0] * 2
    z[2] = 0.0
    add("qr_rankdef_zero", "qr_decomp", z)
    sp = N((3, 12, 6)) * (rng.random((3, 12, 6)) < 0.35)
    add("qr_sparse", "qr_decomp", sp)
    add("qr_graded", "qr_decomp", N((2, 10, 5)) * (10.0 ** -np.arange(0, 25, 5)))
    add("qr_full_tall", "qr_decomp_full", N((3, 7, 4)))
    add("qr_full_wide", "qr_decomp_full", N((2, 4, 7)))
    add("qr_full_square", "qr_decomp_full", N((2, 9, 9)))
    if not small:
        add("qr_c4", "qr_decomp", N((4, 64, 32)))
        add("qr_c4_sparse", "qr_decomp", N((2, 64, 32)) * (rng.random((2, 64, 32)) < 0.2))
        add("qr_full_c4", "qr_decomp_full", N((1, 64, 32)))
        add("qr_48x24", "qr_decomp", N((2, 48, 24)))
    add("qr_inplace", "qr_decomp_inplace", N((3, 9, 5)), N((3, 9, 2)))
    if not small:
        add("qr_inplace_c4", "qr_decomp_inplace", N((3, 64, 32)), N((3, 64, 1)))
    a = N((3, 9, 5))
    q, r = np.linalg.qr(a)
    add("qr_lstsq", "qr_lstsq", q, r, N((3, 9, 2)))
    add("qr_lstsq_bcast", "qr_lstsq", q[:1], r[:1], N((4, 9, 1)))
    if not small:
        q, r = np.linalg.qr(N((3, 64, 32)))
        add("qr_lstsq_c4", "qr_lstsq", q, r, N((3, 64, 1)))

    # --- svd_jac_2sided (src/la/svd_jac_2sided.js, _svd_jac_utils.js) and the solvers on its factors (src/la/svd.js)
    add("svd_square", "svd_jac_2sided", N((3, 6, 6)))
    add("svd_tall", "svd_jac_2sided", N((3, 7, 4)))
    add("svd_wide", "svd_jac_2sided", N((2, 4, 7)))
    add("svd_1x1", "svd_jac_2sided", np.array([[[-3.0]], [[0.0]], [[2.5]]]))
    add("svd_diag", "svd_jac_2sided", np.stack([np.diag(rng.integers(-5, 6, 6).astype(float)) for _ in range(3)]))
    z = N((3, 8, 8))
    z[0, :, 3] = 0.0
    z[1, 5] = z[1, 2]
    z[2] = 0.0
    add("svd_rankdef", "svd_jac_2sided", z)
    add("svd_sparse", "svd_jac_2sided", N((3, 10, 10)) * (rng.random((3, 10, 10)) < 0.25))
    add("svd_graded", "svd_jac_2sided", N((2, 8, 8)) * (10.0 ** -np.arange(0, 32, 4)))
    add("svd_repeated", "svd_jac_2sided", np.stack([np.linalg.qr(N((8, 8)))[0] * 2.0, np.eye(8)]))
    if not small:
        add("svd_16", "svd_jac_2sided", N((3, 16, 16)))
        add("svd_32", "svd_jac_2sided", N((2, 32, 32)))
        add("svd_c5", "svd_jac_2sided", N((2, 64, 64)))
        add("svd_40x30", "svd_jac_2sided", N((2, 40, 30)))
        add("svd_c5_sparse", "svd_jac_2sided", N((1, 64, 64)) * (rng.random((1, 64, 64)) < 0.1))
    sv = np.sort(np.abs(N((4, 9))), axis=-1)[..., ::-1].copy()
    sv[1, 5:] *= 1e-17
    sv[2, :] = 0.0
    sv[3, 8] = 0.0
    add("svd_rank", "svd_rank", sv)
    add("svd_rank_1d", "svd_rank", sv[1])
    u, s, v = np.linalg.svd(N((3, 8, 8)))
    add("svd_lstsq_sq", "svd_lstsq", u, s, v, N((3, 8, 2)))
    add("svd_solve_sq", "svd_solve", u, s, v, N((3, 8, 1)))
    s2 = s.copy()
    s2[1, 5:] = s2[1, 0] * 1e-18
    add("svd_lstsq_rankcut", "svd_lstsq", u, s2, v, N((3, 8, 2)))
    u, s, v = np.linalg.svd(N((2, 9, 5)), full_matrices=False)
    add("svd_lstsq_tall", "svd_lstsq", u, s, v, N((2, 9, 3)))
    u, s, v = np.linalg.svd(N((2, 5, 9)), full_matrices=False)
    add("svd_lstsq_wide", "svd_lstsq", u, s, v, N((2, 5, 1)))
    add("svd_lstsq_bcast", "svd_lstsq", u[:1], s[:1], v[:1], N((3, 5, 2)))
    if not small:
        u, s, v = np.linalg.svd(N((3, 64, 64)))
        add("svd_lstsq_c5", "svd_lstsq", u, s, v, N((3, 64, 1)))
    return cases


_MODS = {
    "matmul2": ("la/matmul.js", "matmul2"), "matmul": ("la/matmul.js", "matmul"),
    "cholesky_decomp": ("la/cholesky.js", "cholesky_decomp"), "cholesky_solve": ("la/cholesky.js", "cholesky_solve"),
    "tril_solve": ("la/tri.js", "tril_solve"), "triu_solve": ("la/tri.js", "triu_solve"),
    "qr_decomp": ("la/qr.js", "qr_decomp"), "qr_decomp_full": ("la/qr.js", "qr_decomp_full"), "qr_lstsq": ("la/qr.js", "qr_lstsq"),
    "svd_jac_2sided": ("la/svd_jac_2sided.js", "svd_jac_2sided"),
    "svd_rank": ("la/svd.js", "svd_rank"), "svd_lstsq": ("la/svd.js", "svd_lstsq"), "svd_solve": ("la/svd.js", "svd_solve"),
}


def run_js(eng, op, ins):
    """Outputs of the reference for one case as a list of numpy arrays, or ('error', message)."""
    nda = eng.module("nd_array.js")
    try:
        if op == "qr_decomp_inplace":      # src/la/qr.js:148-183 works on flat typed arrays, one matrix per call
            a, y = ins
            m, n = a.shape[-2:]
            l = y.shape[-1]
            qr = eng.module("la/qr.js")
            out = eng.call("(function(){ var A=%s, Y=%s; for (var b=0; b<%d; b++) %s._qr_decomp_inplace(%d,%d,%d, A,b*%d, Y,b*%d); return [A,Y]; })()"
                           % (qjs.js_f64(a), qjs.js_f64(y), a.size // (m * n), qr, m, n, l, m * n, m * l))
            return [out[0].reshape(a.shape), out[1].reshape(y.shape)]
        mod, fn = _MODS[op]
        out = eng.call("%s.%s(%s)" % (eng.module(mod), fn, ", ".join(qjs.js_nd(nda, x) for x in ins)))
    except qjs.JSError as e:
        return ("error", str(e))
    return list(out) if isinstance(out, (list, tuple)) else [out]
