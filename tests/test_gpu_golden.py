"""The GPU path against what the REFERENCE ITSELF returned (tests/golden/jsref_golden.npz, see test_jsref_golden.py).

Bars: bit-exact where the kernels replay the reference's per-entry operation sequence (cholesky, the
triangular solves, svd_rank / svd_lstsq / svd_solve, the fused qr_lstsq); entrywise 1e-12 of (|A||B|) for matmul (FMA,
other k order); 1e-12 for QR after sign normalisation and for the singular values, reconstruction and orthogonality of the
SVD (different algorithms: Householder for Givens, one-sided for two-sided Jacobi); the thrown messages verbatim.
"""
import numpy as np
import pytest

from jsref_golden import bits_equal, chain_by_plan, golden
from util import fro, matmul_componentwise_err, qr_sign_normalise, svd_residuals

pytestmark = pytest.mark.gpu
TOL = 1e-12
G = golden()
VALUE = [c for c in G.cases if "error" not in c and c["op"] != "qr_decomp_full"]
ERROR = [c for c in G.cases if "error" in c]
BIT_EXACT = {"cholesky_decomp", "tril_solve", "triu_solve", "cholesky_solve", "svd_rank", "svd_lstsq", "svd_solve"}


def _np(x):
    return [t.numpy() for t in x] if isinstance(x, (tuple, list)) else [x.numpy()]


@pytest.mark.parametrize("case", VALUE, ids=[c["name"] for c in VALUE])
def test_gpu_against_reference_run_vectors(la, case):
    op, ins, want = case["op"], G.ins(case), G.outs(case)
    if op in BIT_EXACT:
        got = _np(getattr(la, op)(*ins))
        for g, w in zip(got, want):
            assert bits_equal(g.reshape(w.shape), w), case["name"]
    elif op == "matmul2":
        (c,) = _np(la.matmul2(*ins))
        assert c.shape == want[0].shape
        assert matmul_componentwise_err(c, want[0], *ins) <= TOL
        if case["name"] == "mm_ints":
            assert np.array_equal(c, want[0])
    elif op == "matmul":
        (c,) = _np(la.matmul(*ins))
        assert c.shape == want[0].shape
        # same parenthesisation as the reference (host logic, pinned bit for bit in test_jsref_golden.py): the error bound is
        # that of the chain of |.| products
        bound = chain_by_plan([np.abs(m) for m in ins], np.matmul, la._chain_plan)
        assert np.max(np.abs(c - want[0]) / np.where(bound == 0, 1.0, bound)) <= TOL
    elif op == "qr_decomp":
        a = ins[0]
        q, r = _np(la.qr_decomp(a))
        qn, rn = qr_sign_normalise(*want)
        assert q.shape == qn.shape and r.shape == rn.shape
        assert (np.tril(r, -1) == 0).all() and (np.diagonal(r, axis1=-2, axis2=-1) >= 0).all()
        assert np.max(fro(q @ r - a) / np.maximum(fro(a), 1e-300)) <= TOL
        assert np.max(np.abs(np.swapaxes(q, -1, -2) @ q - np.eye(q.shape[-1]))) <= TOL
        # rows of R / columns of Q up to the first numerically zero pivot are unique (up to the sign fixed above); past it
        # the reference's own factors are one arbitrary choice among many (the direction is set by rounding noise)
        amax = max(float(np.max(np.abs(a))), 1e-300)
        qtol = 1e-9 if "graded" in case["name"] else TOL
        for ix in np.ndindex(*a.shape[:-2]):
            d = np.abs(np.diagonal(rn[ix]))
            small = np.nonzero(d <= 1e-8 * amax)[0]
            k0 = int(small[0]) if small.size else d.size
            assert np.max(np.abs(r[ix][:k0] - rn[ix][:k0]), initial=0.0) / amax <= TOL, (case["name"], ix)
            assert np.max(np.abs(q[ix][:, :k0] - qn[ix][:, :k0]), initial=0.0) <= qtol, (case["name"], ix)
    elif op == "qr_decomp_inplace":
        a, y = ins
        r, qty = _np(la._qr_decomp_inplace(a, y))
        rref, qref = want
        k = min(a.shape[-2:])
        sg = np.sign(np.diagonal(rref, axis1=-2, axis2=-1))
        sg[sg == 0] = 1.0
        assert np.max(np.abs(r[:, :k] - sg[..., None] * rref[:, :k])) <= TOL
        assert np.max(np.abs(qty[:, :k] - sg[..., None] * qref[:, :k])) <= TOL
    elif op == "qr_lstsq":
        (x,) = _np(la.qr_lstsq(*ins))
        q, r, y = ins
        if q.shape[:-2] == r.shape[:-2] == y.shape[:-2]:
            assert bits_equal(x, want[0]), case["name"]          # fused kernel: the reference's own sequence
        else:
            assert np.max(np.abs(x - want[0])) <= TOL * max(1.0, float(np.max(np.abs(want[0]))))
    elif op == "svd_jac_2sided":
        a = ins[0]
        u, sv, v = _np(la.svd_jac_1sided(a))
        uref, sref, vref = want
        assert u.shape == uref.shape and sv.shape == sref.shape and v.shape == vref.shape
        assert (sv >= 0).all() and not np.signbit(sv).any() and (np.diff(sv, axis=-1) <= 0).all()
        smax = np.maximum(sref[..., :1], 1e-300)
        assert np.max(np.abs(sv - sref) / smax) <= TOL
        recon, ou, ov = svd_residuals(a, u, sv, v)
        assert recon <= TOL and ou <= TOL and ov <= TOL, (recon, ou, ov)
        if case["name"] == "svd_diag":                            # exact on diagonal input, like the reference
            assert np.array_equal(sv, sref)
    else:
        raise AssertionError("unhandled op " + op)


@pytest.mark.parametrize("case", ERROR, ids=[c["name"] for c in ERROR])
def test_gpu_path_throws_the_references_messages(la, case):
    with pytest.raises(Exception) as ei:
        getattr(la, case["op"])(*G.ins(case))
    assert case["error"] in str(ei.value), (case["error"], str(ei.value))
