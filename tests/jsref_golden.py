"""Access to tests/golden/jsref_golden.npz — outputs of the reference's own JavaScript (oracle/jsref/gen_golden.py) —
and the oracle-side evaluation of the same cases."""
import json
import os

import numpy as np

PATH = os.path.join(os.path.dirname(__file__), "golden", "jsref_golden.npz")


class Golden:
    def __init__(self, path=PATH):
        self.z = np.load(path)
        self.manifest = json.loads(bytes(self.z["manifest"]).decode())
        self.cases = self.manifest["cases"]

    def ins(self, c):
        return [self.z["%s.in%d" % (c["name"], i)] for i in range(c["n_in"])]

    def outs(self, c):
        return [self.z["%s.out%d" % (c["name"], i)] for i in range(c["n_out"])]


_golden = None


def golden():
    global _golden
    if _golden is None:
        _golden = Golden()
    return _golden


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f" and b.dtype.kind == "f":
        return np.array_equal(a.astype(np.float64).view(np.uint64), b.astype(np.float64).view(np.uint64))
    return np.array_equal(a, b)


def chain_by_plan(mats, matmul2, chain_plan):
    """Evaluate matmul(...matrices) with the host-side parenthesisation (la._chain_plan) and a given matmul2."""
    plan, _shape = chain_plan([m.shape for m in mats])
    stack = []
    for p in plan:
        if p >= 0:
            stack.append(mats[p])
        else:
            b = stack.pop()
            a = stack.pop()
            stack.append(matmul2(a, b))
    assert len(stack) == 1
    return stack[0]


def oracle_run(op, ins):
    """The C oracle (oracle/nd4ref.c) on one case: list of outputs, or ('error', RefError)."""
    from oracle import nd4ref
    from nd4js_b200 import la
    try:
        if op == "matmul":
            return [chain_by_plan(ins, nd4ref.matmul2, la._chain_plan)]
        if op == "qr_decomp_full":
            return list(nd4ref.qr_decomp_full(*ins))
        if op == "svd_solve":
            return [nd4ref.svd_lstsq(*ins)]
        if op == "qr_decomp_inplace":
            return list(nd4ref.qr_decomp_inplace(*ins))
        out = getattr(nd4ref, op)(*ins)
    except nd4ref.RefError as e:
        return ("error", e)
    return list(out) if isinstance(out, tuple) else [out]
