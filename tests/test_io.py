"""NPY wire format (nd4js src/io/npy.js) — host-side mirror nd4js_b200/io.py against the reference's own golden vectors
(tests/golden/npy_golden.json, extracted from src/io/npy_test_data.js by tests/golden/make_npy_golden.py), NumPy's
reader/writer, and the reference's error texts (src/io/npy_test.js:34-70)."""
import base64
import io as _io
import json
import os

import numpy as np
import pytest

from nd4js_b200 import io as ndio
from nd4js_b200.nd_array import NDArray, from_numpy

GOLDEN = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "npy_golden.json")))["items"]
NP = {"int32": np.int32, "float32": np.float32, "float64": np.float64}


def test_golden_vectors_of_the_reference_deserialize_exactly():
    assert len(GOLDEN) >= 100
    seen = set()
    for item in GOLDEN:
        raw = base64.b64decode(item["b64"])
        got = ndio.npy_deserialize(raw)
        assert got.dtype == item["dtype"]
        assert list(got.shape) == item["shape"]
        want = np.array([float(v) for v in item["data"]], dtype=np.float64).astype(NP[item["dtype"]])
        assert got.data.dtype == NP[item["dtype"]]
        assert (got.data == want).all()
        # NumPy reads the same bytes to the same array (the fixtures were produced by NumPy)
        ref = np.load(_io.BytesIO(raw))
        assert (np.ascontiguousarray(ref).reshape(-1) == got.data).all()
        seen.add((raw[8 + 2:].split(b"'descr': '")[1][:1], b"'fortran_order': True" in raw[:128]))
    assert {(b"<", False), (b">", False), (b"<", True), (b">", True)} <= seen  # both byte orders, C and Fortran order


@pytest.mark.parametrize("dtype", ["int32", "float32", "float64"])
def test_serialize_round_trip_and_numpy_interop(dtype):
    rng = np.random.default_rng(7)
    for _ in range(40):
        shape = tuple(int(s) for s in rng.integers(1, 24, size=rng.integers(1, 5)))
        a = (rng.uniform(-1e3, 1e3, shape) * (rng.uniform(0, 1, shape) > 0.1)).astype(NP[dtype])
        A = from_numpy(a)
        raw = ndio.npy_serialize(A)
        assert bytes(ndio.npy_serialize_gen(A)) == raw
        B = ndio.npy_deserialize(raw)
        assert B.dtype == dtype and (B.shape == A.shape).all() and (B.data == A.data).all()
        back = np.load(_io.BytesIO(raw))
        assert back.dtype == NP[dtype] and back.shape == shape and (back == a).all()
        # and the other direction: NumPy's own writer (version 1.0, single-quoted header)
        f = _io.BytesIO()
        np.save(f, a)
        C_ = ndio.npy_deserialize(f.getvalue())
        assert (C_.shape == A.shape).all() and (C_.data == A.data).all()
        f = _io.BytesIO()
        np.save(f, np.asfortranarray(a).astype(NP[dtype].__name__ and np.dtype(NP[dtype]).newbyteorder(">")))
        D = ndio.npy_deserialize(f.getvalue())
        assert D.dtype == dtype and (D.numpy() == a).all()


def test_header_bytes_are_the_references():
    # npy.js:52-76: double-quoted keys, "(3,4,)" with a trailing comma, blanks up to a multiple of 64, '\n', version 1.0
    raw = ndio.npy_serialize(from_numpy(np.zeros((3, 4))))
    header = '{"descr": "<f8", "fortran_order": False, "shape": (3,4,)}'
    assert raw[:6] == bytes((0x93,)) + b"NUMPY" and raw[6:8] == bytes((1, 0))
    total = ((len(header) + 11 + 63) >> 6) << 6  # npy.js:65
    assert total == 128 and int.from_bytes(raw[8:10], "little") == total - 10
    assert raw[10:10 + len(header)] == header.encode()
    assert raw[10 + len(header):total - 1] == b" " * (total - 11 - len(header)) and raw[total - 1:total] == b"\n"
    assert len(raw) == total + 12 * 8
    zero_d = NDArray(np.zeros(0, np.int32), np.array([2.5]))
    raw0 = ndio.npy_serialize(zero_d)
    assert b'"shape": ()}' in raw0 and ndio.npy_deserialize(raw0).data[0] == 2.5 and ndio.npy_deserialize(raw0).ndim == 0


def test_error_texts():
    good = ndio.npy_serialize(from_numpy(np.arange(6.0).reshape(2, 3)))
    with pytest.raises(ValueError, match="does not start with"):
        ndio.npy_deserialize(b"XNUMPY" + good[6:])
    with pytest.raises(ValueError, match="version 3.0 not supported"):
        ndio.npy_deserialize(good[:6] + bytes((3, 0)) + good[8:])
    with pytest.raises(ValueError, match="ended unexpectedly"):
        ndio.npy_deserialize(good[:-1])
    with pytest.raises(ValueError, match="not yet supported"):
        ndio.npy_deserialize(good.replace(b"<f8", b"<c8"))
    f = _io.BytesIO()
    np.save(f, np.zeros(3, np.complex128))
    with pytest.raises(ValueError, match="dtype '<c16' not yet supported"):
        ndio.npy_deserialize(f.getvalue())


@pytest.mark.gpu
def test_npy_fixture_feeds_the_gpu_path(la, ref):
    # the interchange path end to end: a batch serialised as .npy -> deserialised -> cholesky on the GPU -> .npy -> NumPy
    from util import spd
    s = spd(9, (37,), 16)
    S = ndio.npy_deserialize(ndio.npy_serialize(from_numpy(s)))
    L = la.cholesky_decomp(S)
    back = np.load(_io.BytesIO(ndio.npy_serialize(L)))
    assert (back == ref.cholesky_decomp(s)).all()
