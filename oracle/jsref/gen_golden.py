"""Generate tests/golden/jsref_golden.npz: outputs of the reference's own JavaScript (src/la/*.js of nd4js v1.3.0,
executed in this container by QJSEngine — qjs.py) on the seeded inputs of cases.py.

    python -m oracle.jsref.gen_golden            # rewrites the fixture (needs /root/reference and Nsight Compute's Qt)

The fixture holds inputs and outputs bit for bit, the thrown messages of the failure cases, and known answers of
the reference's random generators (src/rand/alea_rng.js, src/_test_rng.js) that pin oracle/alea.py.
"""
import json
import os
import sys
import time

import numpy as np

from . import cases, qjs

SEED = 20261019
OUT = os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", "jsref_golden.npz")


def rng_vectors(eng):
    """Known answers of AleaRNG / TestRNG straight from the reference (pins oracle/alea.py)."""
    alea = eng.module("rand/alea_rng.js")
    trng = eng.module("_test_rng.js")
    out = {}
    for i, seed in enumerate(["hello.", "nd4b", 1337, "svd_jac_1sided"]):
        s = json.dumps(seed)
        out["alea%d" % i] = {"seed": seed,
                             "uniform": eng.call("(function(){ var r=new %s.AleaRNG(%s), o=new Float64Array(24); for (var i=0;i<24;i++) o[i]=r.uniform(-2,3); return o; })()" % (alea, s)),
                             "int": eng.call("(function(){ var r=new %s.AleaRNG(%s), o=new Float64Array(24); for (var i=0;i<24;i++) o[i]=r.int(-7,1000); return o; })()" % (alea, s)),
                             "normal": eng.call("(function(){ var r=new %s.AleaRNG(%s), o=new Float64Array(24); for (var i=0;i<24;i++) o[i]=r.normal(); return o; })()" % (alea, s)),
                             "bool": eng.call("(function(){ var r=new %s.AleaRNG(%s), o=new Float64Array(24); for (var i=0;i<24;i++) o[i]=r.bool()?1:0; return o; })()" % (alea, s))}
    for i, desc in enumerate(["works for random square matrices", "x"]):
        s = json.dumps(desc)
        out["trng%d" % i] = {"seed": desc,
                             "rankDef": eng.call("(function(){ var r=new %s.TestRNG(%s); return [r.rankDef(5,7), r.rankDef(2,6,6), r.rankDef(4,3)]; })()" % (trng, s)),
                             "ortho": eng.call("(function(){ var r=new %s.TestRNG(%s); return [r.ortho(4), r.ortho(2,5,3), r.ortho(3,6)]; })()" % (trng, s))}
    return out


def main():
    eng = qjs.engine()
    arrays, manifest = {}, {"reference": "nd4js v1.3.0 src/ executed by QJSEngine (Qt 6.6.3)", "seed": SEED, "cases": []}
    t0 = time.time()
    for name, op, ins in cases.make_cases(SEED):
        out = cases.run_js(eng, op, ins)
        entry = {"name": name, "op": op, "n_in": len(ins)}
        for i, x in enumerate(ins):
            arrays["%s.in%d" % (name, i)] = x
        if isinstance(out, tuple):
            entry["error"] = out[1]
        else:
            entry["n_out"] = len(out)
            for i, x in enumerate(out):
                arrays["%s.out%d" % (name, i)] = np.asarray(x)
        manifest["cases"].append(entry)
        print("%-22s %-18s %s" % (name, op, entry.get("error", "ok")), file=sys.stderr)
    rv = rng_vectors(eng)
    manifest["rng"] = {}
    for k, d in rv.items():
        manifest["rng"][k] = {"seed": d["seed"], "fields": [f for f in d if f != "seed"]}
        for f, v in d.items():
            if f == "seed":
                continue
            if isinstance(v, list):
                flat = []
                for x in v:                       # rankDef returns [A, ranks] pairs: stored flat, A0, ranks0, A1, ...
                    flat += x if isinstance(x, list) else [x]
                v = flat
                manifest["rng"][k][f + "_n"] = len(v)
                for i, x in enumerate(v):
                    arrays["%s.%s.%d" % (k, f, i)] = np.asarray(x)
            else:
                arrays["%s.%s" % (k, f)] = np.asarray(v)
    arrays["manifest"] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    np.savez_compressed(OUT, **arrays)
    print("wrote %s: %d cases, %.1f KiB, %.1f s" % (os.path.normpath(OUT), len(manifest["cases"]), os.path.getsize(OUT) / 1024, time.time() - t0), file=sys.stderr)


if __name__ == "__main__":
    main()
