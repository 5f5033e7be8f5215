"""The reference's OWN test suites as the judge of the GPU path — TEST INFRASTRUCTURE.

The reference cannot run where the GPU is, and the GPU is not where the reference is, so this goes in three steps:

  record   (build container)  the reference's Jasmine suites run in QJSEngine (jasmine.py) with hooks in place of
                              qr_decomp / cholesky_decomp / the SVD under test; every argument the suites pass in is
                              written to tests/golden/_suite_calls.npz (git-ignored; it travels to the GPU box)
  compute  (GPU box)          nd4js_b200.la.* on every recorded argument -> gpurun_out/suite_results.npz
  replay   (build container)  the same suites again, same seeds; the hooks check that each argument is the recorded one and
                              hand back the GPU's result, so every `expect(...)` of the reference's own test bodies —
                              shapes, triangularity, orthogonality, reconstruction, ranks, least-squares optimality with
                              the reference's own tolerances — is evaluated on what the CUDA kernels returned.
                              Result: profiles/r02_reference_suites_on_gpu_results.txt

    python -m oracle.jsref.suite_replay record | compute | replay  [groups: a b1 b2]

`Math.random` (cholesky_test.js, qr_test.js draw from it) is replaced by a seeded generator that is re-seeded at the start
of every spec, so that record and replay see the same items; items larger than MAX_ELEMS go to the reference's own function in
both passes; at most MAX_CALLS hooked calls per spec (then the spec's item loop is stopped, and reported as such).
The SVD under test is registered under the name `svd_jac_1sided` — the function north_star adds and the reference lacks.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CALLS = os.path.join(ROOT, "tests", "golden", "_suite_calls.npz")
RESULTS = os.path.join(ROOT, "gpurun_out", "suite_results.npz")
REPORT = os.path.join(ROOT, "profiles", "r02_reference_suites_on_gpu_results.txt")
MAX_ELEMS = 70000
MAX_CALLS = 100
MAX_BIG = 4        # items above MAX_ELEMS per spec that go to the reference's own function before the spec is stopped

HOOK_JS = r"""
var __HOOK = {mode: '%(mode)s', calls: [], n: 0, per_spec: 0, big_per_spec: 0, results: null, mismatches: 0};
(function(){
  var s = 1;
  Math.random = function() { s = (Math.imul(s, 1103515245) + 12345) >>> 0; var hi = s; s = (Math.imul(s, 1103515245) + 12345) >>> 0;
                             return ((hi >>> 6) * 67108864 + (s >>> 6)) / 9007199254740992; };
  __HOOK.reseed = function(k) { s = (k * 2654435761 + 1) >>> 0; };
})();
function __hook_arg(A) { return {shape: Array.from(A.shape), dt: A.dtype, h: __to_hex(A.dtype === 'float64' ? A.data : Float64Array.from(A.data))}; }
function __hook_calln(fn, args, orig, zeros) {      // args: array of array-likes; orig / zeros take the converted NDArrays
  args = args.map(function(A) { return %(nda)s.asarray(A); });
  var big = false, n_f64 = 0;
  args.forEach(function(A) { big = big || A.data.length > %(max_elems)d || !(A.dtype === 'float64' || A.dtype === 'int32') || A.ndim < 2; n_f64 += (A.dtype === 'float64'); });
  if (big || (args.length > 1 && n_f64 === 0)) {     // too large, other dtypes, or an all-int32 product (the reference keeps those int32)
    if (__HOOK.big_per_spec >= %(max_big)d) throw new Error('__STOP__');
    __HOOK.big_per_spec++;
    return orig.apply(null, args);
  }
  if (__HOOK.per_spec >= %(max_calls)d) throw new Error('__STOP__');
  __HOOK.per_spec++;
  var idx = __HOOK.n++, rec = args.map(__hook_arg);
  if (__HOOK.mode === 'record') { __HOOK.calls.push({fn: fn, spec: __J.cur ? __J.cur.name : '', args: rec}); return zeros.apply(null, args); }
  var r = __HOOK.results[idx];
  if (!r || r.fn !== fn || r.h !== rec.map(function(a){ return a.h; }).join('|')) { __HOOK.mismatches++; throw new Error('__MISMATCH__ call ' + idx + ' ' + fn); }
  if (r.err) throw new Error(r.err);
  return r.out.map(function(o) { return new %(nda)s.NDArray(Int32Array.from(o.shape), __from_hex(o.h, Float64Array)); });
}
function __hook_call(fn, A, orig, zeros) { return __hook_calln(fn, [A], orig, zeros); }
function __bshape(a, b, tail) {                      // broadcast leading dims of two shapes + tail
  var la = a.slice(0, -2), lb = b.slice(0, -2), n = Math.max(la.length, lb.length), out = [];
  for (var i = 0; i < n; i++) { var x = la[la.length - n + i] || 1, y = lb[lb.length - n + i] || 1; out.push(Math.max(x, y)); }
  return out.concat(tail);
}
function __zeros(shape) { var n = 1; for (var i = 0; i < shape.length; i++) n *= shape[i]; return new %(nda)s.NDArray(Int32Array.from(shape), new Float64Array(n)); }
function svd_jac_1sided(A) {
  return __hook_call('svd', A, %(svd)s.svd_jac_2sided, function(A) { var s = Array.from(A.shape), M = s[s.length-2], N = s[s.length-1], L = Math.min(M,N), b = s.slice(0,-2);
    return [__zeros(b.concat([M,L])), __zeros(b.concat([L])), __zeros(b.concat([L,N]))]; });
}
__HOOK.qr_decomp = function(A, orig) {
  return __hook_call('qr', A, orig, function(A) { var s = Array.from(A.shape), M = s[s.length-2], N = s[s.length-1], L = Math.min(M,N), b = s.slice(0,-2);
    return [__zeros(b.concat([M,L])), __zeros(b.concat([L,N]))]; });
};
__HOOK.cholesky_decomp = function(S, orig) {
  var r = __hook_call('chol', S, function(S) { return [orig(S)]; }, function(S) { return [__zeros(Array.from(S.shape))]; });
  return r[0];
};
'ok'
"""

HOOK_B1 = r"""
__HOOK.qr_decomp = null; __HOOK.cholesky_decomp = null;
__HOOK.matmul2 = function(a, b, orig) {
  return __hook_calln('matmul2', [a, b], function(a, b) { return [orig(a, b)]; },
                      function(a, b) { var sa = Array.from(a.shape), sb = Array.from(b.shape); return [__zeros(__bshape(sa, sb, [sa[sa.length-2], sb[sb.length-1]]))]; })[0];
};
'ok'
"""

HOOK_B2 = r"""
__HOOK.qr_decomp = null; __HOOK.cholesky_decomp = null;
['tril_solve', 'triu_solve', 'cholesky_solve'].forEach(function(name) {
  __HOOK[name] = function(t, y, orig) {
    return __hook_calln(name, [t, y], function(t, y) { return [orig(t, y)]; },
                        function(t, y) { var st = Array.from(t.shape), sy = Array.from(y.shape); return [__zeros(__bshape(st, sy, sy.slice(-2)))]; })[0];
  };
});
'ok'
"""

_WRAP = [  # (file below src/, exported function) -> the staged copy calls __HOOK.<name>(args..., original) when a hook is installed
    ("la/qr.js", "qr_decomp", "A"),
    ("la/cholesky.js", "cholesky_decomp", "S"),
    ("la/cholesky.js", "cholesky_solve", "L,y"),
    ("la/matmul.js", "matmul2", "a,b"),
    ("la/tri.js", "tril_solve", "L,Y"),
    ("la/tri.js", "triu_solve ", "U,Y"),
]

# group -> (extra hook script, test modules, spec filter)
GROUPS = {
    "a": (None, ["la/qr_test.js", "la/cholesky_test.js"],
          lambda nm: not nm.startswith("qr_decomp_full") and not nm.startswith("_qr_decomp_inplace") and "cholesky_solve" not in nm),
    "b1": (HOOK_B1, ["la/matmul_test.js"], lambda nm: "complex128" not in nm),
    "b2": (HOOK_B2, ["la/tri_test.js", "la/cholesky_test.js"], lambda nm: "tril_solve" in nm or "triu_solve" in nm or "cholesky_solve" in nm),
}


def _install_wrappers(eng):
    """Rewrite the staged copies (temp dir) of the hooked modules BEFORE they are imported."""
    for rel, name, arg in _WRAP:
        p = os.path.join(eng.src, rel)
        text = open(p, encoding="utf-8").read()
        head = "export function %s(%s)" % (name, arg)
        assert text.count(head) == 1, (rel, name)
        text = text.replace(head, "export function %s(%s) { const h = (typeof __HOOK !== 'undefined') ? __HOOK : null; return (h && h.%s) ? h.%s(%s, __orig_%s) : __orig_%s(%s); }\nfunction __orig_%s(%s)"
                            % (name, arg, name, name, arg, name, name, arg, name, arg))
        open(p, "w", encoding="utf-8").write(text)


def _paths(group):
    sfx = "" if group == "a" else "_" + group
    return (CALLS.replace(".npz", sfx + ".npz"), RESULTS.replace(".npz", sfx + ".npz"))


def _runner(mode, group):
    from . import jasmine, qjs
    eng = qjs.Engine()
    _install_wrappers(eng)
    run = jasmine.Runner(eng)
    eng.run(HOOK_JS % {"mode": mode, "nda": eng.module("nd_array.js"), "svd": eng.module("la/svd_jac_2sided.js"),
                       "max_elems": MAX_ELEMS, "max_calls": MAX_CALLS, "max_big": MAX_BIG})
    extra, modules, keep_fn = GROUPS[group]
    if extra:
        eng.run(extra)
    n0 = int(eng.run("__J.specs.length"))
    if group == "a":
        g = eng.module("la/_generic_test_svd_decomp.js")
        eng.run("%s.generic_test_svd_decomp(svd_jac_1sided); 'ok'" % g)
    for m in modules:
        eng.module(m)
    n1 = int(eng.run("__J.specs.length"))
    names = run.names(range(n0, n1))
    keep = [i for i, nm in zip(range(n0, n1), names) if keep_fn(nm)]
    return eng, run, keep


def record(group):
    calls_path, _ = _paths(group)
    eng, run, keep = _runner("record", group)
    for k, i in enumerate(keep):
        eng.run("__HOOK.reseed(%d); __HOOK.per_spec = 0; __HOOK.big_per_spec = 0; 'ok'" % (k + 1))
        r = run.run(i, budget_ms=600000)
        print("%-10s %5d calls so far  %s" % (r["status"], int(eng.run("__HOOK.n")), r["name"][:110]), file=sys.stderr)
    calls = json.loads(eng.run("JSON.stringify(__HOOK.calls)"))
    arrays = {"meta": np.frombuffer(json.dumps([{"fn": c["fn"], "spec": c["spec"], "args": [{"shape": a["shape"], "dt": a["dt"]} for a in c["args"]]}
                                                for c in calls]).encode(), np.uint8)}
    for i, c in enumerate(calls):
        for k, a in enumerate(c["args"]):
            arrays["a%d_%d" % (i, k)] = np.frombuffer(bytes.fromhex(a["h"]), "<f8").reshape(a["shape"])
    np.savez_compressed(calls_path, **arrays)
    print("recorded %d calls -> %s (%.1f MiB)" % (len(calls), calls_path, os.path.getsize(calls_path) / 2 ** 20), file=sys.stderr)


def _args(z, i, m):
    out = []
    for k, a in enumerate(m["args"]):
        x = z["a%d_%d" % (i, k)]
        out.append(x.astype(np.int32) if a["dt"] == "int32" else x)
    return out


def compute(group):
    """On the GPU box: the product path on every recorded argument."""
    import nd4js_b200 as nd
    nd.init([0])
    la = nd.la
    calls_path, results_path = _paths(group)
    z = np.load(calls_path)
    meta = json.loads(bytes(z["meta"]).decode())
    fns = {"svd": la.svd_jac_1sided, "qr": la.qr_decomp, "chol": la.cholesky_decomp, "matmul2": la.matmul2, "tril_solve": la.tril_solve,
           "triu_solve": la.triu_solve, "cholesky_solve": la.cholesky_solve}
    out = {"n": np.array(len(meta))}
    for i, m in enumerate(meta):
        try:
            res = fns[m["fn"]](*_args(z, i, m))
            res = [t.numpy() for t in res] if isinstance(res, (tuple, list)) else [res.numpy()]
            for k, r in enumerate(res):
                out["r%d_%d" % (i, k)] = np.ascontiguousarray(r, dtype=np.float64)
        except Exception as e:   # the suites also feed arguments that must be rejected
            out["e%d" % i] = np.frombuffer(str(e).encode(), np.uint8)
    os.makedirs(os.path.dirname(results_path), exist_ok=True)
    np.savez_compressed(results_path, **out)
    print("group %s: computed %d calls on the GPU (%d launches) -> %s" % (group, len(meta), nd.stats()["kernel_launches"], results_path))


def replay(group):
    calls_path, results_path = _paths(group)
    z, rz = np.load(calls_path), np.load(results_path)
    meta = json.loads(bytes(z["meta"]).decode())
    eng, run, keep = _runner("replay", group)
    results = []
    for i, m in enumerate(meta):
        entry = {"fn": m["fn"], "h": "|".join(np.ascontiguousarray(z["a%d_%d" % (i, k)], "<f8").tobytes().hex() for k in range(len(m["args"])))}
        if "e%d" % i in rz.files:
            entry["err"] = bytes(rz["e%d" % i]).decode()
        else:
            entry["out"] = []
            k = 0
            while "r%d_%d" % (i, k) in rz.files:
                r = rz["r%d_%d" % (i, k)]
                entry["out"].append({"shape": list(r.shape), "h": np.ascontiguousarray(r, "<f8").tobytes().hex()})
                k += 1
        results.append(entry)
    # hand the table over in slices (one giant string literal is slow to parse)
    eng.run("__HOOK.results = []; 'ok'")
    for lo in range(0, len(results), 50):
        eng.run("__HOOK.results = __HOOK.results.concat(%s); 'ok'" % json.dumps(results[lo:lo + 50]))
    lines, tot = [], {"passed": 0, "failed": 0, "stopped": 0, "truncated": 0}
    for k, i in enumerate(keep):
        eng.run("__HOOK.reseed(%d); __HOOK.per_spec = 0; __HOOK.big_per_spec = 0; 'ok'" % (k + 1))
        before = int(eng.run("__HOOK.n"))
        r = run.run(i, budget_ms=600000)
        calls = int(eng.run("__HOOK.n")) - before
        status = r["status"]
        if status == "failed" and len(r["failures"]) == 1 and "__STOP__" in r["failures"][0]:
            status = "stopped"           # MAX_CALLS items judged, none failed
        tot[status] = tot.get(status, 0) + 1
        lines.append("  %-9s %5d GPU results judged %7d expectations  %s" % (status, calls, r["n_expect"], r["name"]))
        if status == "failed":
            lines += ["      " + f[:500] for f in r["failures"]]
        print(lines[-1], file=sys.stderr)
    head = ["## group %s: %s" % (group, ", ".join((["la/_generic_test_svd_decomp.js (+ lstsq / rank / solve generics)"] if group == "a" else []) + GROUPS[group][1])),
            "# %d calls answered by the GPU path; argument mismatches between record and replay: %s" % (len(meta), eng.run("__HOOK.mismatches")),
            "# totals: " + ", ".join("%d %s" % (v, k) for k, v in tot.items() if v)]
    return head + lines + [""]


def report(groups):
    head = ["# The reference's own suites (nd4js v1.3.0) evaluated in QJSEngine on what the CUDA kernels returned for the suites' own items",
            "# (oracle/jsref/suite_replay.py: record here -> compute on the B200 -> replay here; hooked per group: a = qr_decomp, cholesky_decomp and",
            "# the SVD under test, b1 = matmul2, b2 = tril_solve / triu_solve / cholesky_solve).  'stopped' = the first %d items of the spec were" % MAX_CALLS,
            "# judged, none failed.", ""]
    body = []
    for g in groups:
        body += replay(g)
    open(REPORT, "w").write("\n".join(head + body) + "\n")
    print("\n".join(l for l in body if l.startswith("#")), file=sys.stderr)


if __name__ == "__main__":
    groups = sys.argv[2:] or list(GROUPS)
    if sys.argv[1] == "replay":
        report(groups)
    else:
        for g in groups:
            {"record": record, "compute": compute}[sys.argv[1]](g)
