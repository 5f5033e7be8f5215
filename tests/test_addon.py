"""CPU: the Node N-API addon compiles against the hand-declared N-API subset and links against libnd4b.so
with only napi_* symbols left for the Node runtime to supply (there is no Node in this image)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_addon_compiles_and_binds_the_c_abi(tmp_path):
    from nd4js_b200 import _lib
    _lib.load()
    src = os.path.join(ROOT, "nd4js_b200", "addon", "nd4b_napi.cc")
    obj, node = str(tmp_path / "nd4b_napi.o"), str(tmp_path / "nd4b.node")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fPIC", "-c", src, "-o", obj])
    subprocess.check_call(["g++", "-shared", "-o", node, obj, "-L" + os.path.join(ROOT, "nd4js_b200"), "-lnd4b",
                           "-Wl,--unresolved-symbols=ignore-all"])
    syms = subprocess.check_output(["nm", "-D", node], text=True)
    undefined = [l.split()[-1] for l in syms.splitlines() if " U " in l]
    for name in ("nd4b_matmul_f64", "nd4b_cholesky_f64", "nd4b_qr_f64", "nd4b_svd_jac1_f64", "nd4b_matmul_shape",
                 "nd4b_host_alloc", "nd4b_last_error"):
        assert name in undefined, name  # bound to libnd4b.so
    assert any(u.startswith("napi_") for u in undefined)


def test_js_shim_exports_the_reference_names():
    js = open(os.path.join(ROOT, "nd4js_b200", "js", "index.js")).read()
    for name in ("matmul2", "matmul", "cholesky_decomp", "qr_decomp", "svd_jac_1sided"):
        assert "function %s(" % name in js
    for text in ("A must be at least 2D.", "B must be at least 2D.", "Last two dimensions must be quadratic.",
                 "qr_decomp(A): A.ndim must be at least 2.", "svd_jac_1sided(A): A.dtype must be float."):
        assert text in js
