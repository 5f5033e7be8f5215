"""nd4js_b200 — B200-native (sm_100a) batched Float64 dense linear algebra behind nd4js's nd.la API.

    from nd4js_b200 import la
    L = la.cholesky_decomp(S)          # S: NDArray / numpy-like [...,N,N]

The arithmetic lives in libnd4b.so (hand-written CUDA, C ABI in include/nd4b.h); this package is the
host-side mirror of the reference's operator interface plus the batch partitioner.
"""
from . import la  # noqa: F401
from ._lib import Nd4bError, host_trim, init, load, pinned_array, pinned_empty, stats  # noqa: F401
from .nd_array import NDArray, asarray, from_numpy  # noqa: F401

__all__ = ["la", "NDArray", "asarray", "from_numpy", "init", "load", "stats", "Nd4bError", "pinned_array", "pinned_empty",
           "host_trim"]
