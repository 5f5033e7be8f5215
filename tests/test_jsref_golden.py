"""The oracle against the REFERENCE ITSELF.

tests/golden/jsref_golden.npz holds what nd4js v1.3.0's own JavaScript (src/la/*.js, src/rand/alea_rng.js, src/_test_rng.js)
returned for seeded inputs when executed in the build container by QJSEngine (oracle/jsref/gen_golden.py; the engine is
the Qt 6.6.3 one Nsight Compute ships).  These tests pin, bit for bit:
  * oracle/nd4ref.c (the C restatement every GPU parity test is measured against),
  * tests/ref_mirror.py (the independent NumPy restatement),
  * oracle/alea.py (AleaRNG / TestRNG, which regenerate the reference's own test inputs),
  * the host-side chain ordering of nd4js_b200.la (matmul.js:150-236),
and the thrown messages of the failure cases.  The GPU path is held to the same vectors in test_gpu_golden.py.
"""
import numpy as np
import pytest

from jsref_golden import bits_equal, golden, oracle_run

G = golden()
VALUE = [c for c in G.cases if "error" not in c]
ERROR = [c for c in G.cases if "error" in c]


def test_fixture_is_complete():
    ops = {c["op"] for c in G.cases}
    assert ops >= {"matmul2", "matmul", "cholesky_decomp", "cholesky_solve", "tril_solve", "triu_solve", "qr_decomp", "qr_decomp_full",
                   "qr_decomp_inplace", "qr_lstsq", "svd_jac_2sided", "svd_rank", "svd_lstsq", "svd_solve"}
    assert len(VALUE) >= 70 and len(ERROR) >= 5
    assert "QJSEngine" in G.manifest["reference"]


@pytest.mark.parametrize("case", VALUE, ids=[c["name"] for c in VALUE])
def test_c_oracle_is_bit_identical_with_the_reference(case):
    got = oracle_run(case["op"], G.ins(case))
    assert not isinstance(got, tuple), got
    want = G.outs(case)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert bits_equal(g, w), (case["name"], np.abs(np.asarray(g, float) - np.asarray(w, float)).max())


_ORACLE_TEXT = {  # messages the oracle's status codes stand for (oracle/nd4ref.py) vs the reference's texts
    "chol_fail_not_pd": "Matrix contains NaNs or is (near) singular.",
    "chol_fail_nan": "Assertion failed.",
    "chol_fail_shape": "Last two dimensions must be quadratic.",
}


@pytest.mark.parametrize("case", ERROR, ids=[c["name"] for c in ERROR])
def test_failure_cases_fail_in_the_oracle_too(case):
    try:
        got = oracle_run(case["op"], G.ins(case))
    except ValueError:            # numpy's broadcast check in the Python wrapper
        return
    assert isinstance(got, tuple) and got[0] == "error", "the reference throws %r" % case["error"]
    if case["name"] in _ORACLE_TEXT:
        assert case["error"] == _ORACLE_TEXT[case["name"]] == str(got[1])


@pytest.mark.parametrize("case", [c for c in VALUE if c["op"] in ("qr_decomp", "qr_decomp_full", "svd_jac_2sided") and max(G.ins(c)[0].shape[-2:]) <= 16],
                         ids=lambda c: c["name"])
def test_numpy_mirror_is_bit_identical_with_the_reference(case):
    import ref_mirror
    a = G.ins(case)[0]
    want = G.outs(case)
    fn = getattr(ref_mirror, case["op"])
    lead = a.shape[:-2]
    for ix in np.ndindex(*lead):
        got = fn(a[ix])
        for g, w in zip(got, want):
            assert bits_equal(np.asarray(g, float).reshape(w[ix].shape), w[ix]), (case["name"], ix)


def test_alea_restatement_matches_the_reference_generators():
    from oracle import alea
    z, rng = G.z, G.manifest["rng"]
    for key, info in rng.items():
        if key.startswith("alea"):
            for field, draw in (("uniform", lambda r: r.uniform(-2, 3)), ("int", lambda r: r.int(-7, 1000)), ("normal", lambda r: r.normal()),
                                ("bool", lambda r: 1.0 if r.bool() else 0.0)):
                r = alea.AleaRNG(info["seed"])
                got = np.array([draw(r) for _ in range(24)], dtype=np.float64)
                assert bits_equal(got, z["%s.%s" % (key, field)]), (key, field)
        else:
            r = alea.TestRNG(info["seed"])
            for i, shape in enumerate([(5, 7), (2, 6, 6), (4, 3)]):
                a, ranks = r.rank_def(*shape)
                assert bits_equal(a, z["%s.rankDef.%d" % (key, 2 * i)].reshape(a.shape)), (key, shape)
                assert np.array_equal(np.asarray(ranks).ravel(), z["%s.rankDef.%d" % (key, 2 * i + 1)].ravel())
            r = alea.TestRNG(info["seed"])
            for i, shape in enumerate([(4,), (2, 5, 3), (3, 6)]):
                q = r.ortho(*shape)
                assert bits_equal(q, z["%s.ortho.%d" % (key, i)].reshape(q.shape)), (key, shape)
