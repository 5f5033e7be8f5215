"""AleaRNG and TestRNG of the reference, restated — TEST INFRASTRUCTURE ONLY (tests/ and tools/ import it).

The reference's test suites draw every input from `new TestRNG(description)` (src/jasmine_utils.js:268-279), a seeded Alea
generator, so restating the generator reproduces the reference's own test inputs without a JavaScript engine:
  mash, AleaRNG.{constructor, __next, bool, int, uniform, normal, ortho}   src/rand/alea_rng.js:37-142, 168-227
  TestRNG.rankDef                                                           src/_test_rng.js:27-64
  tabulate (row-major visiting order)                                       src/tabulate.js:23-52
Pinned by the known answers of the Alea generator this class restates (Baagoe's Alea with one seed string has the same
seeding and recurrence; the `alea('hello.')` vectors of the seedrandom package: 0.4783254903741181, then the 53-bit draw
0.8297006866124559, then the int32 draw 1076136327) — tests/test_oracle.py::test_alea_known_answers.
JS semantics kept: doubles are Python floats, `x >>> 0` is ToUint32, `x | 0` is ToInt32.  `normal` uses the platform's
log (V8's may differ in the last bit; inputs drawn through `normal` are "the reference's generator", not its exact bits).
"""
import math

import numpy as np

MUL32, DIV32, DIV53 = 2.0 ** 32, 2.0 ** -32, 2.0 ** -53


def _u32(x):
    return int(x) % 4294967296           # ToUint32 (truncation, then modulo 2^32)


def _i32(x):
    v = int(x) % 4294967296              # ToInt32
    return v - 4294967296 if v >= 2147483648 else v


def mash(text, seed):
    for ch in str(text):
        seed += ord(ch)
        temp = 0.02519603282416938 * seed
        seed = _u32(temp)
        temp -= seed
        temp *= seed
        seed = _u32(temp)
        temp -= seed
        seed += temp * MUL32
    return seed


def _giv_rot_qr(a, b):   # src/la/_giv_rot.js:20-37
    mx = max(abs(a), abs(b))
    if mx == 0:
        return 1.0, 0.0, 0.0
    a /= mx
    b /= mx
    norm = math.sqrt(a * a + b * b)
    return a / norm, b / norm, norm * mx


class AleaRNG:
    def __init__(self, seed):
        if seed is None:
            raise ValueError("Assertion failed.")
        seed = str(seed)
        s0 = mash(" ", 0xefc8249d)
        s1 = mash(" ", s0)
        s2 = mash(" ", s1)
        t0 = mash(seed, s2)
        t1 = mash(seed, t0)
        t2 = mash(seed, t1)

        def fold(s, t):
            s = float(_u32(s) - _u32(t)) * DIV32
            return s + (1.0 if s < 0 else 0.0)

        self.s0, self.s1, self.s2, self.c = fold(s0, t0), fold(s1, t1), fold(s2, t2), 1
        self._next_normal = math.nan

    def _next(self):
        t = 2091639 * self.s0 + self.c * DIV32
        self.s0, self.s1 = self.s1, self.s2
        self.c = _i32(t)
        self.s2 = t - self.c
        return self.s2

    def uniform(self, lo=-1.0, hi=1.0):
        lo, hi = float(lo), float(hi)
        a = self._next()
        s = a + _i32(self._next() * 0x200000) * DIV53
        return lo * (1 - s) + s * hi

    def bool(self):
        return self.uniform() < 0.0

    def int(self, lo, until=None):
        if until is None:
            lo, until = 0, lo
        if not lo < until:
            raise ValueError("AlreaRNG::int(from,until): from must be less than until.")
        return math.floor(self.uniform(lo, until))

    def normal(self, mean=0.0, sigma=1.0):
        nxt = self._next_normal
        if not math.isnan(nxt):
            self._next_normal = math.nan
            return nxt * sigma + mean
        while True:                       # Marsaglia polar method, alea_rng.js:154-160
            x = self.uniform()
            y = self.uniform()
            r = x * x + y * y
            if not (r > 1 or r == 0):
                break
        z = math.sqrt(-2 * math.log(r) / r)
        self._next_normal = z * x
        return mean + z * y * sigma

    def ortho(self, *shape):
        """Batch of random (semi-)orthogonal matrices [..., M, N] (alea_rng.js:168-227)."""
        shape = [int(s) for s in shape]
        if len(shape) == 1:
            shape.append(shape[0])
        m, n = shape[-2:]
        k, l = max(m, n), min(m, n)
        out = np.zeros(int(np.prod(shape)))
        q = np.zeros(k * l)
        for u_off in range(out.size - m * n, -1, -m * n):
            for i in range(k - 1, -1, -1):
                for j in range(l - 1, -1, -1):
                    q[l * i + j] = 0.0 if i != j else (-1.0 if self.bool() else 1.0)
            for j in range(k):
                a_jj = self.normal()
                for i in range(j + 1, k):
                    a_ij = self.normal()
                    c, s, norm = _giv_rot_qr(a_jj, a_ij)
                    if s == 0:
                        continue
                    a_jj = norm
                    w = min(i + 1, l)
                    qi, qj = q[l * j:l * j + w].copy(), q[l * i:l * i + w].copy()   # _giv_rot_rows(Q, w, L*j, L*i, c, s)
                    q[l * j:l * j + w] = c * qi + s * qj
                    q[l * i:l * i + w] = c * qj - s * qi
            qm = q.reshape(k, l)
            out[u_off:u_off + m * n] = (qm.T if m < n else qm).reshape(-1)
        return out.reshape(shape)


def tabulate(shape, fn):
    """Row-major fill, one fn() call per element in index order (src/tabulate.js:36-50)."""
    shape = [int(s) for s in shape]
    n = int(np.prod(shape)) if shape else 1
    return np.array([fn() for _ in range(n)], dtype=np.float64).reshape(shape)


class TestRNG(AleaRNG):
    __test__ = False   # not a pytest class

    def rank_def(self, *shape):
        """[A, ranks]: A = U S V from random orthogonal factors and mostly rank-deficient spectra (src/_test_rng.js:29-63)."""
        shape = [int(s) for s in shape]
        n = shape.pop()
        m = shape.pop()
        l = min(m, n)
        cnt = int(np.prod(shape)) if shape else 1
        ranks = np.array([self.int(0, l + 1) for _ in range(cnt)], dtype=np.int32)
        u = self.ortho(*shape, m, l)
        v = self.ortho(*shape, l, n).reshape(cnt, l, n)
        for i in range(cnt - 1, -1, -1):
            for j in range(l - 1, -1, -1):
                scale = 0.0 if ranks[i] <= j else self.uniform(1e-4, 1e+4)
                v[i, j, :] *= scale
        from . import nd4ref
        a = nd4ref.matmul2(u, v.reshape(shape + [l, n]))
        return a, ranks.reshape(shape)
