"""CPU: the C-ABI library loads, exports every symbol include/nd4b.h declares, and its host-only
logic (shape inference, argument errors, loud failure without a device) matches the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "nd4b.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nd4b_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from nd4js_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.SYMBOLS) == names


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "nd4js_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".js")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "nd4ref" not in src and "oracle" not in src.lower(), f


def _shape_call(lib, a, b):
    a_s, b_s = np.asarray(a, np.int32), np.asarray(b, np.int32)
    out = np.zeros(max(len(a), len(b), 2), np.int32)
    nd = C.c_int(0)
    rc = lib.nd4b_matmul_shape(C.c_void_p(a_s.ctypes.data), len(a), C.c_void_p(b_s.ctypes.data), len(b),
                               C.c_void_p(out.ctypes.data), C.byref(nd))
    return rc, tuple(int(x) for x in out[: nd.value])


def test_matmul_shape_matches_oracle(ref):
    from nd4js_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for _ in range(300):
        a = [int(x) for x in rng.integers(1, 4, rng.integers(1, 6))]
        b = [int(x) for x in rng.integers(1, 4, rng.integers(1, 6))]
        if rng.random() < 0.6 and len(a) >= 2 and len(b) >= 2:
            b[-2] = a[-1]
        rc, shape = _shape_call(lib, a, b)
        try:
            want = ref.matmul_shape(a, b)
            assert rc == 0 and shape == want, (a, b)
        except ref.RefError as e:
            assert rc == e.code, (a, b)
            assert _lib.last_error() == str(e)


def test_error_texts_are_the_references():
    from nd4js_b200 import _lib
    lib = _lib.load()
    assert _shape_call(lib, [3], [3, 3])[0] == _lib.E_A_NDIM and _lib.last_error() == "A must be at least 2D."
    assert _shape_call(lib, [3, 3], [3])[0] == _lib.E_B_NDIM and _lib.last_error() == "B must be at least 2D."
    assert _shape_call(lib, [2, 3], [4, 2])[0] == _lib.E_INNER
    assert _lib.last_error() == "The last dimension of A and the 2nd to last dimension of B do not match."
    assert _shape_call(lib, [2, 4, 3], [3, 3, 2])[0] == _lib.E_BROADCAST
    assert _lib.last_error() == "Shapes are not broadcast-compatible."


def test_null_and_bad_arguments_are_rejected_before_any_device_work():
    from nd4js_b200 import _lib
    lib = _lib.load()
    assert lib.nd4b_cholesky_f64(None, None, 1, 4, None) == _lib.E_ARG
    x = np.zeros(16)
    p = C.c_void_p(x.ctypes.data)
    assert lib.nd4b_cholesky_f64(p, p, 0, 4, None) == _lib.E_ARG
    assert lib.nd4b_qr_f64(p, p, p, 1, 0, 4) == _lib.E_ARG
    assert lib.nd4b_svd_jac1_f64(p, p, p, p, 1, 4, -1, None) == _lib.E_ARG


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import nd4js_b200
    with pytest.raises(nd4js_b200.Nd4bError, match="no CPU fallback"):
        nd4js_b200.la.cholesky_decomp(np.eye(4))
    with pytest.raises(nd4js_b200.Nd4bError, match="no CPU fallback"):
        nd4js_b200.la.matmul2(np.eye(4), np.eye(4))
