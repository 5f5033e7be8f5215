"""Run the reference's own Jasmine suites (src/**/*_test.js) inside QJSEngine — TEST INFRASTRUCTURE.

A small synchronous stand-in for the Jasmine globals the suites use (describe / it / beforeEach / expect / jasmine.addMatchers
with the reference's CUSTOM_MATCHERS from src/jasmine_utils.js, setInterval-driven `forEachItemIn` loops).  Two uses:
  * the suites run against the reference's own functions: shows that this engine executes nd4js as its authors intended
    (every expectation of the path's suites holds), which is what makes the golden vectors trustworthy;
  * `record` / `replay` hooks (see suites_on_outputs) put another implementation's results under the SAME assertions.
"""
import json

from . import qjs

SHIM = r"""
var __J = {stack: [], before: [], specs: [], matchers: {}, cur: null, timers: [], tid: 0, n_expect: 0};
var performance = (typeof performance !== 'undefined') ? performance : {now: function(){ return Date.now(); }};
var console = (typeof console !== 'undefined' && console.warn) ? console : {log: function(){}, warn: function(){}, error: function(){}, info: function(){}};
function setInterval(fn, ms) { var id = ++__J.tid; __J.timers.push({id: id, fn: fn}); return id; }
function setTimeout(fn, ms) { var id = ++__J.tid; __J.timers.push({id: id, fn: fn, once: true}); return id; }
function clearInterval(id) { __J.timers = __J.timers.filter(function(t){ return t.id !== id; }); }
var clearTimeout = clearInterval;
// forEachItemIn wraps its loop in `new Promise`; nothing here awaits, so a synchronous stand-in records the outcome
function Promise(executor) {
  var self = this; self.state = 'pending';
  try { executor(function(v){ if (self.state === 'pending') { self.state = 'ok'; self.value = v; } },
                 function(e){ if (self.state === 'pending') { self.state = 'err'; self.value = e; } }); }
  catch (e) { self.state = 'err'; self.value = e; }
}
Promise.prototype.then = function(f, g) { if (this.state === 'ok' && f) f(this.value); if (this.state === 'err' && g) g(this.value); return this; };
Promise.prototype.catch = function(g) { return this.then(null, g); };
Promise.resolve = function(v) { return new Promise(function(res){ res(v); }); };

function __Spec(name, fn, before) { this.name = name; this.fn = fn; this.before = before; this.failures = []; this.n_expect = 0; this.status = 'pending'; }
__Spec.prototype.addExpectationResult = function(passed, data, isError) {
  this.n_expect++;
  if (!passed) this.failures.push('' + (data && (data.message !== undefined ? data.message : (data.error && data.error.message))));
};
function describe(name, fn) { __J.stack.push(name); var nb = __J.before.length; try { fn(); } finally { __J.stack.pop(); __J.before.length = nb; } }
var fdescribe = describe; function xdescribe() {} function xit() {}
function beforeEach(fn) { __J.before.push(fn); }
function it(name, fn, timeout) { var s = new __Spec(__J.stack.concat([name]).join(' > '), fn, __J.before.slice()); __J.specs.push(s); return s; }
var fit = it;
var jasmine = {addMatchers: function(m) { for (var k in m) __J.matchers[k] = m[k]; }, DEFAULT_TIMEOUT_INTERVAL: 1e9,
               any: function(T){ return {__any: T}; }};

function __eq(a, b) {
  if (b && b.__any) return (b.__any === Number) ? typeof a === 'number' : (b.__any === String) ? typeof a === 'string' : a instanceof b.__any;
  if (Object.is(a, b)) return true;
  if (typeof a === 'number' && typeof b === 'number') return a === b;          // jasmine: +0 and -0 differ? (treated equal here)
  if (a == null || b == null || typeof a !== 'object' && typeof a !== 'function' || typeof b !== 'object' && typeof b !== 'function') return false;
  if (Object.getPrototypeOf(a) !== Object.getPrototypeOf(b)) return false;
  if (ArrayBuffer.isView(a) || Array.isArray(a)) {
    if (a.length !== b.length) return false;
    for (var i = 0; i < a.length; i++) if (!__eq(a[i], b[i])) return false;
    return true;
  }
  var ka = Object.keys(a), kb = Object.keys(b);
  if (ka.length !== kb.length) return false;
  for (var j = 0; j < ka.length; j++) if (!__eq(a[ka[j]], b[ka[j]])) return false;
  return true;
}
var __builtin = {
  toBe: function(a, b) { return Object.is(a, b) || a === b; },
  toEqual: function(a, b) { return __eq(a, b); },
  toMatch: function(a, re) { return (re instanceof RegExp) ? re.test(a) : ('' + a).indexOf(re) >= 0; },
  toBeGreaterThan: function(a, b) { return a > b; }, toBeGreaterThanOrEqual: function(a, b) { return a >= b; },
  toBeLessThan: function(a, b) { return a < b; }, toBeLessThanOrEqual: function(a, b) { return a <= b; },
  toBeDefined: function(a) { return a !== undefined; }, toBeUndefined: function(a) { return a === undefined; },
  toBeNull: function(a) { return a === null; }, toBeNaN: function(a) { return a !== a; },
  toBeTruthy: function(a) { return !!a; }, toBeFalsy: function(a) { return !a; },
  toBeTrue: function(a) { return a === true; }, toBeFalse: function(a) { return a === false; },
  toContain: function(a, x) { return Array.prototype.some.call(a, function(y){ return __eq(y, x); }); },
  toBeInstanceOf: function(a, T) { return a instanceof T; },
  toThrow: function(f) { try { f(); } catch (e) { return true; } return false; },
  toThrowError: function(f) { try { f(); } catch (e) { return e instanceof Error; } return false; }
};
function __short(x) { var s; try { s = '' + x; } catch (e) { s = '?'; } return s.length > 300 ? s.substr(0, 300) + '...' : s; }
function expect(actual) {
  function make(negate) {
    return new Proxy({}, {get: function(_t, name) {
      if (name === 'not') return make(!negate);
      if (name === 'withContext') return function() { return make(negate); };
      return function() {
        var args = Array.prototype.slice.call(arguments), res;
        if (__J.matchers[name]) res = __J.matchers[name]({}, []).compare.apply(null, [actual].concat(args));
        else if (__builtin[name]) res = {pass: __builtin[name].apply(null, [actual].concat(args))};
        else throw new TypeError('jasmine stand-in: unknown matcher ' + String(name));
        var pass = negate ? !res.pass : !!res.pass;
        __J.n_expect++;
        __J.cur.addExpectationResult(pass, {message: pass ? '' : (res.message || ('Expected ' + __short(actual) + (negate ? ' not ' : ' ') + String(name) + ' ' + args.map(__short).join(', ')))}, false);
      };
    }});
  }
  return make(false);
}
function fail(msg) { __J.cur.addExpectationResult(false, {message: '' + msg}, false); }

function __run_spec(idx, budget_ms) {
  var s = __J.specs[idx], t0 = Date.now();
  __J.cur = s; __J.timers = []; __J.n_expect = 0;
  try {
    for (var i = 0; i < s.before.length; i++) s.before[i]();
    var r = s.fn();
    var rounds = 0;
    while (__J.timers.length > 0) {
      if (Date.now() - t0 > budget_ms) { s.status = 'truncated'; __J.timers = []; break; }
      var ts = __J.timers.slice();
      for (var k = 0; k < ts.length; k++) { if (ts[k].once) clearInterval(ts[k].id); ts[k].fn(); }
      rounds++;
    }
    if (r && r.state === 'err') throw r.value;
    if (s.status === 'pending') s.status = s.failures.length ? 'failed' : 'passed';
  } catch (e) {
    s.status = 'failed';
    s.failures.push('threw: ' + (e && e.message !== undefined ? e.message : e) + (e && e.stack ? ' | ' + ('' + e.stack).substr(0, 400) : ''));
  }
  __J.cur = null;
  return JSON.stringify({name: s.name, status: s.status, n_expect: __J.n_expect, ms: Date.now() - t0, failures: s.failures.slice(0, 3).map(function(m){ return __short(m); })});
}
"""


class Runner:
    def __init__(self, eng=None):
        self.eng = eng or qjs.engine()
        self.eng.run(SHIM + "\n'ok'")
        self.loaded = []

    def load(self, rel):
        """Import a *_test.js file: its describe/it calls register specs."""
        n0 = int(self.eng.run("__J.specs.length"))
        self.eng.module(rel)
        n1 = int(self.eng.run("__J.specs.length"))
        self.loaded.append((rel, n0, n1))
        return range(n0, n1)

    def names(self, idxs):
        return [self.eng.run("__J.specs[%d].name" % i) for i in idxs]

    def run(self, idx, budget_ms=20000):
        return json.loads(self.eng.run("__run_spec(%d, %d)" % (idx, budget_ms)))

    def run_file(self, rel, budget_ms=20000, verbose=False):
        out = []
        for i in self.load(rel):
            r = self.run(i, budget_ms)
            if verbose:
                print("  %-9s %6d ms %7d expects  %s" % (r["status"], r["ms"], r["n_expect"], r["name"]))
                for f in r["failures"]:
                    print("      " + f[:600])
            out.append(r)
        return out


if __name__ == "__main__":
    import os
    import sys
    budget = int(os.environ.get("JSREF_SPEC_BUDGET_MS", "20000"))
    files = sys.argv[1:] or ["kahan_sum_test.js", "la/matmul_test.js", "la/cholesky_test.js", "la/qr_test.js", "la/tri_test.js",
                             "la/_svd_jac_utils_test.js", "la/svd_jac_2sided_test.js", "rand/alea_rng_test.js", "_test_rng_test.js", "io/npy_test.js"]
    r = Runner()
    for f in files:
        print(f)
        try:
            r.run_file(f, budget_ms=budget, verbose=True)
        except qjs.JSError as e:      # e.g. rand/alea_rng_test.js needs the npm package 'seedrandom'
            print("  not loadable: %s" % e)
        sys.stdout.flush()
