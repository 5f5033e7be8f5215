"""NDArray — the value type at the nd.la boundary, mirroring nd4js src/nd_array.js:128-147:
`shape` is an Int32Array (numpy int32, read-only), `data` a flat typed array (numpy, C order),
every dim >= 1 and data.length == prod(shape)."""
import numpy as np

_DTYPES = {"int32": np.int32, "float32": np.float32, "float64": np.float64}


class NDArray:
    __slots__ = ("shape", "data")

    def __init__(self, shape, data):
        if not (isinstance(shape, np.ndarray) and shape.dtype == np.int32):
            raise ValueError("Shape must be Int32Array.")
        if (shape < 1).any():
            raise ValueError("Invalid shape: %s." % ",".join(map(str, shape)))
        if data.ndim != 1 or data.size != int(np.prod(shape, dtype=np.int64)):
            raise ValueError("Shape [%s] does not match array length of %d." % (",".join(map(str, shape)), data.size))
        shape = shape.copy()
        shape.setflags(write=False)  # Object.freeze(shape.buffer), nd_array.js:142
        self.shape, self.data = shape, data

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def dtype(self):
        for name, t in _DTYPES.items():
            if self.data.dtype == t:
                return name
        return "object"

    @property
    def T(self):  # nd_array.js:362-366 — a copy with the last two axes swapped
        a = self.numpy()
        if a.ndim < 2:
            return NDArray(self.shape, self.data.copy())
        return from_numpy(np.ascontiguousarray(np.swapaxes(a, -1, -2)))

    def numpy(self):
        return self.data.reshape(tuple(int(s) for s in self.shape))

    def __repr__(self):
        return "NDArray(shape=%s, dtype=%s)" % (list(self.shape), self.dtype)


def from_numpy(a):
    a = np.ascontiguousarray(a)
    if a.ndim == 0:
        raise ValueError("Invalid shape: scalars are not NDArrays.")
    return NDArray(np.asarray(a.shape, np.int32), a.reshape(-1))


def asarray(x):
    """nd_array.js:102-126: NDArrays pass through, nested sequences are copied into a new NDArray."""
    if isinstance(x, NDArray):
        return x
    a = np.asarray(x)  # numpy input is wrapped without a copy (inputs are never written); nested lists are copied
    if a.dtype.kind == "c":
        raise TypeError("complex128 is outside the float64 hot path")
    if a.dtype.kind in "iub":
        a = a.astype(np.int32, copy=False)
    elif a.dtype != np.float32:
        a = a.astype(np.float64, copy=False)
    return from_numpy(a)
