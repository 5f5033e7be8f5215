"""NPY wire format — the interchange/fixture path of the nd.la hot path, mirroring nd4js
`npy_serialize`, `npy_serialize_gen` and `npy_deserialize` (src/io/npy.js:28-85 and :88-187): batches produced here
can be fed to the JS reference and vice versa, byte for byte.

Same header text as the reference writes (`{"descr": "<f8", "fortran_order": False, "shape": (3,4,)}`, padded with
blanks to a multiple of 64 bytes, version 1.0), same accepted inputs on the reading side (versions 1.0 and 2.0, both
byte orders, Fortran order, dtypes i4 / f4 / f8) and the same error texts.  complex128 ('c16') is outside the Float64
path and raises like an unsupported dtype.
"""
import ast
import sys

import numpy as np

from .nd_array import NDArray, asarray

MAGIC_STRING = bytes((0x93,)) + b"NUMPY"
IS_LITTLE_ENDIAN = sys.byteorder == "little"
_DESCR = {"int32": "i4", "float32": "f4", "float64": "f8"}
_DTYPE = {"i4": np.int32, "f4": np.float32, "f8": np.float64}


def npy_serialize_gen(A):
    """Yields the bytes of the .npy file one by one, as the reference's generator does (npy.js:34-85)."""
    yield from npy_serialize(A)


def npy_serialize(A):
    A = asarray(A)
    if A.dtype not in _DESCR:
        raise ValueError("nd_to_npy: A.dtype=%s not yet supported." % A.dtype)
    dt = ("<" if IS_LITTLE_ENDIAN else ">") + _DESCR[A.dtype]
    shape = ",".join(str(int(s)) for s in A.shape) + ("," if A.ndim > 0 else "")
    header = '{"descr": "%s", "fortran_order": False, "shape": (%s)}' % (dt, shape)
    header_len = ((len(header) + 11 + 63) >> 6) << 6
    if header_len > 0xFFFF:
        raise ValueError("nd_to_npy: Header too large.")
    out = bytearray(MAGIC_STRING)
    out += bytes((1, 0, (header_len - 10) & 255, ((header_len - 10) >> 8) & 255))
    out += header.encode("latin-1")
    out += b" " * (header_len - len(header) - 11) + b"\n"
    assert len(out) == header_len and header_len % 64 == 0
    out += np.ascontiguousarray(A.data).tobytes()
    return bytes(out)


def npy_deserialize(npy_bytes):
    buf = bytes(bytearray(npy_bytes))  # any iterable of byte values, like the reference
    pos = 0

    def take(n):
        nonlocal pos
        if pos + n > len(buf):
            raise ValueError("npy_to_nd: byte sequence ended unexpectedly.")
        pos += n
        return buf[pos - n:pos]

    if take(6) != MAGIC_STRING:
        raise ValueError("npy_to_nd: byte sequence does not start with '\\u0093NUMPY'.")
    major, minor = take(2)
    version = "%d.%d" % (major, minor)
    if version not in ("1.0", "2.0"):
        raise ValueError("npy_bytes: npy-file version %s not supported." % version)
    header_len = int.from_bytes(take(2 if version == "1.0" else 4), "little")
    header = ast.literal_eval(take(header_len).decode("latin-1"))  # the reference parses it with its PYON reader
    descr = header["descr"]
    if descr[0] not in "<>" or descr[1:] not in _DTYPE:
        raise ValueError("npy_to_nd: dtype '%s' not yet supported." % descr)
    dtype = np.dtype(_DTYPE[descr[1:]])
    shape = [int(s) for s in header["shape"]]
    count = 1
    for s in shape:
        count *= s
    data = np.frombuffer(take(count * dtype.itemsize), dtype=dtype.newbyteorder(descr[0])).astype(dtype)
    if header["fortran_order"] and len(shape) > 1:
        data = np.ascontiguousarray(data.reshape(shape[::-1]).transpose()).reshape(-1)
    return NDArray(np.asarray(shape, np.int32), data)
