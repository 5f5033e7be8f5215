"""The JS shim (nd4js_b200/js/index.js) EXECUTED — in Qt's QJSEngine, next to the reference's own modules.

There is no Node in the image, so the N-API addon cannot be loaded; its place is taken by a mock whose entry points have the
addon's signatures (addon/nd4b_napi.cc) and compute with the reference's own functions.  What is under test is therefore the
shim's JavaScript: argument handling, dtype upcasts, result shapes, the matrix-chain ordering (it must pick the reference's
parenthesisation: results are compared bit for bit), the composed qr_lstsq path, the argument juggling of svd_lstsq / svd_solve,
and the error texts it raises itself.  Skipped where the engine or the reference checkout is missing.
"""
import json
import os

import numpy as np
import pytest

from oracle.jsref import qjs

pytestmark = pytest.mark.skipif(not qjs.available(), reason="needs /root/reference and the Qt 6 QJSEngine shipped with Nsight Compute")
SHIM = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nd4js_b200", "js", "index.js")

MOCK = r"""
var REF = {nda: %(nda)s, mm: %(mm)s, ch: %(ch)s, qr: %(qr)s, svd: %(svd)s, svs: %(svs)s, tri: %(tri)s};
var __calls = {};
function __nd(shape, data) { return new REF.nda.NDArray(Int32Array.from(shape), data); }
function __count(n) { __calls[n] = (__calls[n] || 0) + 1; }
var __mock_addon = {
  pinnedFloat64Array: function(n) { __count('pinned'); return new Float64Array(n); },
  init: function() {}, stats: function() { return __calls; },
  matmulShape: function(as, bs, out) { __count('matmulShape');
    var c = REF.mm.matmul2(__nd(as, new Float64Array(as.reduce(function(x,y){return x*y;},1))), __nd(bs, new Float64Array(bs.reduce(function(x,y){return x*y;},1))));
    if (c.shape.length !== out.length) throw new Error('mock: ndim'); out.set(c.shape); },
  matmul: function(a, as, b, bs, c, cs) { __count('matmul'); c.set(REF.mm.matmul2(__nd(as, a), __nd(bs, b)).data); },
  matmulPlan: function(datas, shapes, plan, c, shape) { __count('matmulPlan');
    var st = [];
    for (var i = 0; i < plan.length; i++) { if (plan[i] >= 0) st.push(__nd(shapes[plan[i]], datas[plan[i]])); else { var b = st.pop(), a = st.pop(); st.push(REF.mm.matmul2(a, b)); } }
    if (st.length !== 1 || st[0].data.length !== c.length) throw new Error('mock: plan');
    c.set(st[0].data); },
  cholesky: function(s, L, batch, N) { __count('cholesky'); L.set(REF.ch.cholesky_decomp(__nd([batch, N, N], s)).data); },
  qr: function(a, Q, R, batch, N, M) { __count('qr'); var r = REF.qr.qr_decomp(__nd([batch, N, M], a)); Q.set(r[0].data); R.set(r[1].data); },
  qrInplace: function(a, y, R, QtY, batch, M, N, L) { __count('qrInplace'); var A = Float64Array.from(a), Y = Float64Array.from(y);
    for (var b = 0; b < batch; b++) REF.qr._qr_decomp_inplace(M, N, L, A, b*M*N, Y, b*M*L); R.set(A); QtY.set(Y); },
  svdJac1: function(a, U, sv, V, batch, N, M) { __count('svdJac1'); var r = REF.svd.svd_jac_2sided(__nd([batch, N, M], a)); U.set(r[0].data); sv.set(r[1].data); V.set(r[2].data); },
  triSolve: function(op, t, ts, y, ys, X, xs) { __count('triSolve');
    var f = [REF.tri.tril_solve, REF.tri.triu_solve, REF.ch.cholesky_solve][op], r = f(__nd(ts, t), __nd(ys, y));
    if (r.data.length !== X.length) throw new Error('mock: triSolve shape'); X.set(r.data); },
  qrLstsq: function(q, r, y, X, batch, N, M, I, J) { __count('qrLstsq'); X.set(REF.qr.qr_lstsq(__nd([batch, N, M], q), __nd([batch, M, I], r), __nd([batch, N, J], y)).data); },
  svdRank: function(d, r, N) { __count('svdRank'); r.set(REF.svs.svd_rank(__nd([d.length / N, N], d)).data); },
  svdLstsqShape: function(us, ss, vs, ys, out) { __count('svdLstsqShape');
    var z = function(s) { return __nd(s, new Float64Array(Array.from(s).reduce(function(x,y){return x*y;},1)).fill(1)); };
    var x = REF.svs.svd_lstsq(z(us), z(ss), z(vs), z(ys)); out.set(x.shape); return x.shape.length; },
  svdLstsq: function(u, us, s, ss, v, vs, y, ys, X, xs) { __count('svdLstsq'); X.set(REF.svs.svd_lstsq(__nd(us, u), __nd(ss, s), __nd(vs, v), __nd(ys, y)).data); }
};
var SHIM = (function() {
  var module = {exports: {}}, exports = module.exports;
  var require = function(name) { if (name === 'nd4js') return {NDArray: REF.nda.NDArray, asarray: REF.nda.asarray}; if (/nd4b\.node$/.test(name)) return __mock_addon; throw new Error('require ' + name); };
  %(src)s
  return module.exports;
})();
'ok'
"""


@pytest.fixture(scope="module")
def js():
    eng = qjs.engine()
    mods = {"nda": eng.module("nd_array.js"), "mm": eng.module("la/matmul.js"), "ch": eng.module("la/cholesky.js"), "qr": eng.module("la/qr.js"),
            "svd": eng.module("la/svd_jac_2sided.js"), "svs": eng.module("la/svd.js"), "tri": eng.module("la/tri.js")}
    assert eng.run(MOCK % dict(mods, src=open(SHIM, encoding="utf-8").read().replace("'use strict';", ""))) == "ok"
    return eng, mods["nda"]


def _same(js, shim_expr, ref_expr):
    eng, _ = js
    a, b = eng.call(shim_expr), eng.call(ref_expr)
    a = a if isinstance(a, list) else [a]
    b = b if isinstance(b, list) else [b]
    assert len(a) == len(b)
    for x, y in zip(a, b):
        x, y = np.asarray(x), np.asarray(y)
        assert x.shape == y.shape and x.dtype == y.dtype, (shim_expr, x.shape, y.shape, x.dtype, y.dtype)
        assert np.array_equal(x.view(np.uint8) if x.dtype.kind == "f" else x, y.view(np.uint8) if y.dtype.kind == "f" else y), shim_expr


def _err(js, expr):
    eng, _ = js
    with pytest.raises(qjs.JSError) as ei:
        eng.call(expr)
    return str(ei.value)


def test_shim_exports_the_path(js):
    eng, _ = js
    names = set(json.loads(eng.run("JSON.stringify(Object.keys(SHIM))")))
    assert names >= {"matmul2", "matmul", "cholesky_decomp", "qr_decomp", "_qr_decomp_inplace", "svd_jac_1sided", "tril_solve", "triu_solve",
                     "cholesky_solve", "qr_lstsq", "svd_rank", "svd_lstsq", "svd_solve", "init", "stats", "pinnedFloat64Array"}


def test_shim_matches_the_reference_through_the_mock_addon(js):
    eng, nda = js
    rng = np.random.default_rng(5)
    N = rng.standard_normal
    nd = lambda a: qjs.js_nd(nda, a)
    a, b = N((2, 1, 3, 4)), N((5, 4, 2))
    _same(js, "SHIM.matmul2(%s, %s)" % (nd(a), nd(b)), "REF.mm.matmul2(%s, %s)" % (nd(a), nd(b)))
    ai = "new REF.nda.NDArray(Int32Array.of(2,3), Int32Array.of(1,2,3,4,5,6))"          # int32 is upcast: float64 out, same values
    got = eng.call("SHIM.matmul2(%s, %s)" % (ai, nd(N((3, 2)))))
    assert got.dtype == np.float64 and got.shape == (2, 2)
    s = N((3, 6, 6))
    s = s @ s.transpose(0, 2, 1) + 6 * np.eye(6)
    _same(js, "SHIM.cholesky_decomp(%s)" % nd(s), "REF.ch.cholesky_decomp(%s)" % nd(s))
    for shape in [(2, 7, 4), (2, 4, 7), (3, 5, 5)]:
        x = N(shape)
        _same(js, "SHIM.qr_decomp(%s)" % nd(x), "REF.qr.qr_decomp(%s)" % nd(x))
        _same(js, "SHIM.svd_jac_1sided(%s)" % nd(x), "REF.svd.svd_jac_2sided(%s)" % nd(x))
    low = np.linalg.cholesky(s)
    y = N((3, 6, 2))
    _same(js, "SHIM.cholesky_solve(%s, %s)" % (nd(low), nd(y)), "REF.ch.cholesky_solve(%s, %s)" % (nd(low), nd(y)))
    _same(js, "SHIM.tril_solve(%s, %s)" % (nd(low[0]), nd(y)), "REF.tri.tril_solve(%s, %s)" % (nd(low[0]), nd(y)))
    _same(js, "SHIM.triu_solve(%s, %s)" % (nd(low.transpose(0, 2, 1)), nd(y[:1])), "REF.tri.triu_solve(%s, %s)" % (nd(low.transpose(0, 2, 1)), nd(y[:1])))
    q, r = np.linalg.qr(N((3, 9, 5)))
    yy = N((3, 9, 2))
    _same(js, "SHIM.qr_lstsq(%s, %s, %s)" % (nd(q), nd(r), nd(yy)), "REF.qr.qr_lstsq(%s, %s, %s)" % (nd(q), nd(r), nd(yy)))
    _same(js, "SHIM.qr_lstsq([%s, %s], %s)" % (nd(q), nd(r), nd(yy)), "REF.qr.qr_lstsq(%s, %s, %s)" % (nd(q), nd(r), nd(yy)))
    u, sv, v = np.linalg.svd(N((2, 6, 6)))
    y6 = N((2, 6, 1))
    args = "%s, %s, %s, %s" % (nd(u), nd(sv), nd(v), nd(y6))
    _same(js, "SHIM.svd_lstsq(%s)" % args, "REF.svs.svd_lstsq(%s)" % args)
    _same(js, "SHIM.svd_lstsq([%s, %s, %s], %s)" % (nd(u), nd(sv), nd(v), nd(y6)), "REF.svs.svd_lstsq(%s)" % args)
    _same(js, "SHIM.svd_solve(%s)" % args, "REF.svs.svd_solve(%s)" % args)
    _same(js, "SHIM.svd_rank(%s)" % nd(sv), "REF.svs.svd_rank(%s)" % nd(sv))
    # _qr_decomp_inplace keeps the reference's in-place contract on flat arrays with offsets
    A, Y = N(3 + 5 * 3), N(2 + 5 * 2)
    _same(js, "(function(){ var A=%s, Y=%s; SHIM._qr_decomp_inplace(5,3,2, A,3, Y,2); return [A,Y]; })()" % (qjs.js_f64(A), qjs.js_f64(Y)),
          "(function(){ var A=%s, Y=%s; REF.qr._qr_decomp_inplace(5,3,2, A,3, Y,2); return [A,Y]; })()" % (qjs.js_f64(A), qjs.js_f64(Y)))


def test_shim_qr_lstsq_composed_path(js):
    # broadcast operands take the matmul2 + triu_solve composition: same values as the reference to rounding
    eng, nda = js
    rng = np.random.default_rng(6)
    nd = lambda a: qjs.js_nd(nda, a)
    q, r = np.linalg.qr(rng.standard_normal((1, 9, 5)))
    y = rng.standard_normal((4, 9, 2))
    got = eng.call("SHIM.qr_lstsq(%s, %s, %s)" % (nd(q), nd(r), nd(y)))
    want = eng.call("REF.qr.qr_lstsq(%s, %s, %s)" % (nd(q), nd(r), nd(y)))
    assert got.shape == want.shape and np.max(np.abs(got - want)) <= 1e-13


@pytest.mark.parametrize("seed", range(12))
def test_shim_chain_takes_the_references_parenthesisation(js, seed):
    # random chains with broadcast leading dims; the mock multiplies in the order of the shim's plan with the reference's matmul2,
    # so a different parenthesisation (or tie-break) shows up as different bits
    eng, nda = js
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(3, 7))
    dims = [int(rng.integers(1, 9)) for _ in range(n + 1)]
    mats = []
    for i in range(n):
        lead = tuple(int(rng.choice([1, 2, 3])) if rng.random() < 0.5 else 1 for _ in range(int(rng.integers(0, 3))))
        lead = tuple(2 if d > 1 else 1 for d in lead)          # broadcast-compatible: every leading dim is 1 or 2
        mats.append(rng.standard_normal(lead + (dims[i], dims[i + 1])))
    args = ", ".join(qjs.js_nd(nda, m) for m in mats)
    _same(js, "SHIM.matmul(%s)" % args, "REF.mm.matmul(%s)" % args)


def test_shim_error_texts(js):
    eng, nda = js
    nd = lambda a: qjs.js_nd(nda, a)
    one = "new REF.nda.NDArray(Int32Array.of(3), Float64Array.of(1,2,3))"
    m23, m32, m33 = nd(np.ones((2, 3))), nd(np.ones((3, 2))), nd(np.eye(3))
    for shim, ref in [("SHIM.matmul2(%s, %s)" % (one, m33), "REF.mm.matmul2(%s, %s)" % (one, m33)),
                      ("SHIM.matmul2(%s, %s)" % (m33, one), "REF.mm.matmul2(%s, %s)" % (m33, one)),
                      ("SHIM.matmul2(%s, %s)" % (m23, m23), "REF.mm.matmul2(%s, %s)" % (m23, m23)),
                      ("SHIM.cholesky_decomp(%s)" % m23, "REF.ch.cholesky_decomp(%s)" % m23),
                      ("SHIM.cholesky_decomp(%s)" % nd(-np.eye(3)), "REF.ch.cholesky_decomp(%s)" % nd(-np.eye(3))),
                      ("SHIM.qr_lstsq(%s, %s, %s)" % (m32, m23, m23), "REF.qr.qr_lstsq(%s, %s, %s)" % (m32, m23, m23)),
                      ("SHIM.tril_solve(%s, %s)" % (one, m33), "REF.tri.tril_solve(%s, %s)" % (one, m33)),
                      ("SHIM.cholesky_solve(%s, %s)" % (m33, one), "REF.ch.cholesky_solve(%s, %s)" % (m33, one)),
                      ("SHIM.svd_solve(%s, %s, %s, %s)" % (m32, nd(np.ones(2)), m23 if False else nd(np.ones((2, 2))), m32),
                       "REF.svs.svd_solve(%s, %s, %s, %s)" % (m32, nd(np.ones(2)), nd(np.ones((2, 2))), m32))]:
        assert _err(js, shim) == _err(js, ref), shim
    c64 = "new REF.nda.NDArray(Int32Array.of(2,2), Float32Array.of(1,2,3,4))"
    assert "not supported by the GPU path" in _err(js, "SHIM.matmul2(%s, %s)" % (c64, c64))     # no CPU fallback: other dtypes throw


def test_the_mock_addon_was_on_the_path(js):
    eng, _ = js
    calls = json.loads(eng.run("JSON.stringify(SHIM.stats())"))
    for name in ("matmulShape", "matmul", "matmulPlan", "cholesky", "qr", "qrInplace", "svdJac1", "triSolve", "qrLstsq", "svdRank", "svdLstsqShape", "svdLstsq"):
        assert calls.get(name, 0) > 0, (name, calls)
