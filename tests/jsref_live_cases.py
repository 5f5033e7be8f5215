"""Live cross-check against the reference's JavaScript, executed here by QJSEngine (oracle/jsref/qjs.py).

Runs only where the reference checkout and Nsight Compute's Qt libraries exist (the build container; the GPU box has no
/root/reference) — everywhere else the committed vectors of test_jsref_golden.py stand in.  Fresh seeds, so this is not the
fixture over again; and the reference's own suite generators are compared with their restatement in ref_suites.py.
"""
import numpy as np
import pytest

from oracle.jsref import qjs

pytestmark = pytest.mark.skipif(not qjs.available(), reason="needs /root/reference and the Qt 6 QJSEngine shipped with Nsight Compute")


@pytest.fixture(scope="module")
def eng():
    return qjs.engine()


@pytest.mark.parametrize("seed", [11, 12])
def test_oracle_bit_identical_with_live_reference(eng, seed):
    from jsref_golden import bits_equal, oracle_run
    from oracle.jsref import cases
    n = 0
    for name, op, ins in cases.make_cases(seed, small=True):
        js = cases.run_js(eng, op, ins)
        try:
            got = oracle_run(op, ins)
        except ValueError:
            assert isinstance(js, tuple), name
            continue
        if isinstance(js, tuple):
            assert isinstance(got, tuple), (name, js)
            continue
        assert not isinstance(got, tuple), (name, got)
        for g, w in zip(got, js):
            assert bits_equal(np.asarray(g).reshape(np.asarray(w).shape), w), name
        n += 1
    assert n >= 50


def test_restated_suite_generators_produce_the_references_items(eng):
    """The first 20 items every spec of _generic_test_svd_decomp.js hands to the SVD under test, recorded from the running
    suite, equal the items of tests/ref_suites.py bit for bit (so the GPU tests that consume ref_suites run on the
    reference's own inputs)."""
    import ref_suites as rs
    from oracle.jsref import jasmine
    K = 20
    run = jasmine.Runner(eng)
    g = eng.module("la/_generic_test_svd_decomp.js")
    svd = eng.module("la/svd_jac_2sided.js")
    eng.run("""var __REC = {};
      function rec(A) { var k = __J.cur.name, l = (__REC[k] = __REC[k] || []);
        if (l.length >= %d) throw new Error('__STOP__');
        l.push({shape: Array.from(A.shape), dt: A.dtype, h: A.dtype === 'float64' ? __to_hex(A.data) : ''});
        return %s.svd_jac_2sided(A); }
      'ok'""" % (K, svd))
    n0 = int(eng.run("__J.specs.length"))
    eng.run("describe('rec', function(){ %s.generic_test_svd_decomp(rec); }); 'ok'" % g)
    n1 = int(eng.run("__J.specs.length"))
    names = run.names(range(n0, n1))
    want = {
        " correctly decomposes random batches of diagonal matrices": (a for _sv, a in rs.diagonal_batches(K)),
        " correctly decomposes random examples": rs.random_examples(K, False),
        " correctly decomposes random examples with occasional zeros": rs.random_examples(K, True),
        " correctly decomposes random rank-deficient examples": rs.rank_deficient_examples(K),
        " correctly decomposes random sparse examples": rs.sparse_examples(K),
        "accurately decomposes random matrices": rs.random_matrices(K, False),
        "accurately decomposes random matrices with occasional zeros": rs.random_matrices(K, True),
        "accurately decomposes random rank-deficient matrices of shape [N+0,N+0]": rs.rank_deficient_matrices(0, 0),
        "accurately decomposes random rank-deficient matrices of shape [N+0,N+1]": rs.rank_deficient_matrices(0, 1),
        "accurately decomposes random rank-deficient matrices of shape [N+1,N+0]": rs.rank_deficient_matrices(1, 0),
        "accurately decomposes random sparse matrices": rs.sparse_matrices(K),
    }
    seen = 0
    for i, name in zip(range(n0, n1), names):
        tail = name.split(" > ")[-1]
        if "[generic SVD tests]" not in name or tail not in want:
            continue
        run.run(i, budget_ms=60000)
        import json
        items = json.loads(eng.run("JSON.stringify(__REC[%s] || [])" % json.dumps(name)))
        assert len(items) == K, (name, len(items))
        for k, (item, a) in enumerate(zip(items, want[tail])):
            assert item["dt"] == "float64" and tuple(item["shape"]) == a.shape, (name, k, item["shape"], a.shape)
            got = np.frombuffer(bytes.fromhex(item["h"]), "<f8").reshape(a.shape)
            assert np.array_equal(got.view(np.uint64), np.ascontiguousarray(a).view(np.uint64)), (name, k)
        seen += 1
    assert seen == len(want)


def test_npy_wire_format_both_ways_against_the_live_reference(eng):
    """nd4js_b200.io against src/io/npy.js running in the engine: bytes written here are read there and vice versa."""
    from nd4js_b200 import io
    from nd4js_b200.nd_array import from_numpy
    npy = eng.module("io/npy.js")
    nda = eng.module("nd_array.js")
    rng = np.random.default_rng(3)
    for shape in [(3, 4), (2, 3, 5), (7,), (1, 1, 1), (2, 16, 16), (5, 1)]:
        a = rng.standard_normal(shape)
        js_bytes = bytes(eng.call("%s.npy_serialize(%s)" % (npy, qjs.js_nd(nda, a))).tolist())
        ours = bytes(io.npy_serialize(from_numpy(a)))
        assert ours == js_bytes, shape                                     # byte for byte the reference's file
        back = io.npy_deserialize(js_bytes)
        assert tuple(back.shape) == shape and np.array_equal(np.asarray(back.data).reshape(shape), a)
        got = eng.call("%s.npy_deserialize(__from_hex('%s', Uint8Array))" % (npy, ours.hex()))
        assert got.shape == shape and np.array_equal(got.view(np.uint64), a.view(np.uint64))
    i32 = np.arange(-6, 6, dtype=np.int32).reshape(3, 4)
    js_bytes = bytes(eng.call("%s.npy_serialize(new %s.NDArray(Int32Array.of(3,4), Int32Array.from([%s])))" % (npy, nda, ",".join(map(str, i32.ravel())))).tolist())
    assert bytes(io.npy_serialize(from_numpy(i32))) == js_bytes
