"""ctypes binding of libnd4b.so (include/nd4b.h).

There is no CPU fallback: if the library is missing it is built with nvcc; if it cannot be loaded,
or a call is made without a usable B200, the call raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("ND4B_LIB_PATH") or os.path.join(_HERE, "libnd4b.so")   # ND4B_LIB_PATH: an experimental build for A/B timing

# symbols declared in include/nd4b.h (tests check that every one is exported)
SYMBOLS = [
    "nd4b_init", "nd4b_shutdown", "nd4b_device_count", "nd4b_last_error", "nd4b_version",
    "nd4b_host_alloc", "nd4b_host_free", "nd4b_host_trim", "nd4b_set_chunk_bytes", "nd4b_get_stats", "nd4b_reset_stats",
    "nd4b_matmul_shape", "nd4b_matmul_f64", "nd4b_cholesky_f64", "nd4b_qr_f64", "nd4b_svd_jac1_f64",
    "nd4b_dev_matmul_f64", "nd4b_dev_cholesky_f64", "nd4b_dev_qr_f64", "nd4b_dev_svd_jac1_f64",
    "nd4b_dev_qr_workspace", "nd4b_dev_svd_workspace", "nd4b_probe_fp64", "nd4b_selfcheck_ieee", "nd4b_tri_solve_f64",
    "nd4b_dev_svd_sweep_counter", "nd4b_dev_svd_pre_sweep_counter", "nd4b_qr_inplace_f64", "nd4b_dev_qr_inplace_f64", "nd4b_matmul_plan_f64", "nd4b_dev_tri_solve_f64", "nd4b_qr_lstsq_f64", "nd4b_dev_qr_lstsq_f64",
    "nd4b_svd_rank_f64", "nd4b_svd_lstsq_shape", "nd4b_svd_lstsq_f64", "nd4b_dev_svd_lstsq_f64", "nd4b_dev_all_gather_f64",
]

OK, E_SINGULAR = 0, 1
E_A_NDIM, E_B_NDIM, E_INNER, E_BROADCAST, E_SHAPE, E_NOT_SQUARE, E_NAN_INPUT, E_ARG, E_CUDA, E_NO_CONVERGENCE = range(-1, -11, -1)


class Stats(C.Structure):
    _fields_ = [("calls", C.c_uint64), ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("staged_bytes", C.c_uint64), ("last_sweeps", C.c_int32),
                ("n_devices", C.c_int32)]


class Nd4bError(RuntimeError):
    """Raised for every non-zero status; str() is the reference's message text where one exists."""

    def __init__(self, code, message, first_bad=-1):
        super().__init__(message)
        self.code, self.first_bad = code, first_bad


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        from . import build as _build
        _build.build()
    L = C.CDLL(SO_PATH)
    dp, ip, i64, vp = C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p
    sig = {
        "nd4b_init": ([C.c_void_p, C.c_int], C.c_int),
        "nd4b_shutdown": ([], C.c_int),
        "nd4b_device_count": ([], C.c_int),
        "nd4b_last_error": ([], C.c_char_p),
        "nd4b_version": ([], C.c_char_p),
        "nd4b_host_alloc": ([C.c_size_t], C.c_void_p),
        "nd4b_host_free": ([C.c_void_p], None),
        "nd4b_host_trim": ([], None),
        "nd4b_set_chunk_bytes": ([C.c_size_t], C.c_int),
        "nd4b_get_stats": ([C.POINTER(Stats)], C.c_int),
        "nd4b_reset_stats": ([], C.c_int),
        "nd4b_matmul_shape": ([ip, C.c_int, ip, C.c_int, ip, C.POINTER(C.c_int)], C.c_int),
        "nd4b_matmul_f64": ([dp, ip, C.c_int, dp, ip, C.c_int, dp, ip, C.c_int], C.c_int),
        "nd4b_cholesky_f64": ([dp, dp, i64, C.c_int, C.POINTER(i64)], C.c_int),
        "nd4b_qr_f64": ([dp, dp, dp, i64, C.c_int, C.c_int], C.c_int),
        "nd4b_svd_jac1_f64": ([dp, dp, dp, dp, i64, C.c_int, C.c_int, C.POINTER(C.c_int)], C.c_int),
        "nd4b_dev_matmul_f64": ([C.c_int, vp, dp, i64, dp, i64, dp, i64, C.c_int, C.c_int, C.c_int], C.c_int),
        "nd4b_dev_cholesky_f64": ([C.c_int, vp, dp, dp, i64, C.c_int, vp], C.c_int),
        "nd4b_dev_qr_f64": ([C.c_int, vp, dp, dp, dp, i64, C.c_int, C.c_int, dp, C.c_size_t], C.c_int),
        "nd4b_dev_svd_jac1_f64": ([C.c_int, vp, dp, dp, dp, dp, i64, C.c_int, C.c_int, vp, dp, C.c_size_t], C.c_int),
        "nd4b_dev_qr_workspace": ([i64, C.c_int, C.c_int], C.c_size_t),
        "nd4b_dev_svd_workspace": ([i64, C.c_int, C.c_int], C.c_size_t),
        "nd4b_tri_solve_f64": ([C.c_int, dp, ip, C.c_int, dp, ip, C.c_int, dp, ip, C.c_int], C.c_int),
        "nd4b_dev_tri_solve_f64": ([C.c_int, vp, C.c_int, dp, i64, dp, i64, dp, i64, C.c_int, C.c_int], C.c_int),
        "nd4b_qr_lstsq_f64": ([dp, dp, dp, dp, i64, C.c_int, C.c_int, C.c_int, C.c_int], C.c_int),
        "nd4b_dev_qr_lstsq_f64": ([C.c_int, vp, dp, dp, dp, dp, i64, C.c_int, C.c_int, C.c_int, C.c_int], C.c_int),
        "nd4b_matmul_plan_f64": ([C.c_int, vp, vp, vp, ip, C.c_int, dp, ip, C.c_int], C.c_int),
        "nd4b_qr_inplace_f64": ([dp, dp, dp, dp, i64, C.c_int, C.c_int, C.c_int], C.c_int),
        "nd4b_dev_qr_inplace_f64": ([C.c_int, vp, dp, dp, dp, dp, i64, C.c_int, C.c_int, C.c_int], C.c_int),
        "nd4b_dev_svd_sweep_counter": ([C.c_int, vp], C.c_int),
        "nd4b_dev_svd_pre_sweep_counter": ([C.c_int, vp], C.c_int),
        "nd4b_dev_all_gather_f64": ([vp, vp, vp, vp], C.c_int),
        "nd4b_svd_rank_f64": ([dp, ip, i64, C.c_int], C.c_int),
        "nd4b_svd_lstsq_shape": ([ip, C.c_int, ip, C.c_int, ip, C.c_int, ip, C.c_int, ip, C.POINTER(C.c_int)], C.c_int),
        "nd4b_svd_lstsq_f64": ([dp, ip, C.c_int, dp, ip, C.c_int, dp, ip, C.c_int, dp, ip, C.c_int, dp, ip, C.c_int], C.c_int),
        "nd4b_dev_svd_lstsq_f64": ([C.c_int, vp, dp, dp, dp, dp, dp, i64, C.c_int, C.c_int, C.c_int, C.c_int, vp], C.c_int),
        "nd4b_probe_fp64": ([C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)], C.c_int),
        "nd4b_selfcheck_ieee": ([C.c_int, C.c_longlong, C.c_ulonglong, C.POINTER(C.c_ulonglong)], C.c_int),
    }
    for name, (args, res) in sig.items():
        f = getattr(L, name)
        f.argtypes, f.restype = args, res
    _lib = L
    return L


def last_error():
    return (load().nd4b_last_error() or b"").decode()


def check(rc, first_bad=-1):
    if rc != 0:
        raise Nd4bError(rc, last_error(), first_bad)


def init(devices=None):
    L = load()
    if devices:
        arr = (C.c_int * len(devices))(*devices)
        check(L.nd4b_init(arr, len(devices)))
    else:
        check(L.nd4b_init(None, 0))


PINNED_MIN_BYTES = 1 << 20   # smaller results are ordinary numpy arrays: the staging copy of a few KiB costs nothing


class _PinnedBlock:
    """Owner of one nd4b_host_alloc block; numpy arrays made from it keep it alive through __array_interface__."""
    __slots__ = ("ptr", "__array_interface__", "__weakref__")

    def __init__(self, ptr, n):
        self.ptr = ptr
        self.__array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}

    def __del__(self):
        p, self.ptr = self.ptr, None
        if p and _lib is not None:
            _lib.nd4b_host_free(p)


def pinned_empty(n, dtype=None):
    """Flat float64 array of n elements in page-locked memory (plain numpy below PINNED_MIN_BYTES)."""
    import numpy as np
    n = int(n)
    if n * 8 < PINNED_MIN_BYTES:
        return np.empty(n, np.float64)
    ptr = load().nd4b_host_alloc(n * 8)
    if not ptr:
        raise MemoryError(last_error())
    return np.asarray(_PinnedBlock(ptr, n))


def pinned_array(a):
    """Copy of a numpy-like array in page-locked memory, as an NDArray: inputs built this way skip the staging copy
    (the counterpart of the addon's pinnedFloat64Array for JS callers)."""
    import numpy as np
    from .nd_array import NDArray
    a = np.asarray(a, dtype=np.float64)
    flat = pinned_empty(a.size)
    flat[...] = a.reshape(-1)
    return NDArray(np.asarray(a.shape, np.int32), flat)


def host_trim():
    load().nd4b_host_trim()


def stats():
    s = Stats()
    check(load().nd4b_get_stats(C.byref(s)))
    return {k: getattr(s, k) for k, _ in Stats._fields_}
