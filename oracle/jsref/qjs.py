"""Run the reference's own JavaScript in this container — TEST INFRASTRUCTURE, never on the product path.

The image has no node / d8 / qjs binary, but Nsight Compute ships Qt 6.6.3 including libQt6Qml.so.6, whose
QJSEngine is a complete ECMAScript 2016 engine (classes, generators, destructuring, typed arrays, ES modules).
This module drives it through ctypes (Itanium C++ ABI by hand: `this` first, an sret pointer before it for
class-type returns) so that `/root/reference/src/**/*.js` — the reference's own sources, unmodified apart from the
`.js` suffix ES-module loaders outside bundlers need on relative import specifiers — execute here and produce the golden vectors under
`tests/golden/` (see `gen_golden.py`).  Nothing is copied into the repo: the suffix-patched tree lives in a temp
directory for the lifetime of the engine.

Qt's libQt6Core wants libglib-2.0 (absent from the image); a stub with the sixteen g_main_context_* / g_source_*
symbols is compiled into the same temp directory and `QT_NO_GLIB=1` keeps Qt from ever calling them.
"""
import ctypes
import glob
import os
import re
import shutil
import subprocess
import tempfile

REFERENCE = os.environ.get("ND4B_REFERENCE", "/root/reference")
_c = ctypes


def _qt_dir():
    hits = sorted(glob.glob("/opt/nvidia/nsight-compute/*/host/linux-desktop-glibc_*-x64/libQt6Qml.so.6"))
    return os.path.dirname(hits[-1]) if hits else None


def available():
    return _qt_dir() is not None and os.path.isdir(os.path.join(REFERENCE, "src")) and shutil.which("gcc") is not None


_GLIB_SYMS = """g_main_context_default g_main_context_iteration g_main_context_new g_main_context_pop_thread_default
g_main_context_push_thread_default g_main_context_ref g_main_context_unref g_main_context_wakeup g_source_add_poll
g_source_attach g_source_destroy g_source_new g_source_remove_poll g_source_set_can_recurse g_source_set_name
g_source_unref""".split()

_IMPORT = re.compile(r"""(\bfrom\s*|\bimport\s*)(['"])(\.{1,2}/[^'"]*|\.{1,2})\2""")


def stage_sources(dst):
    """Mirror reference/src into dst with resolvable relative specifiers ('./qr' -> './qr.js', '../dt' -> '../dt/index.js')."""
    src_root = os.path.join(REFERENCE, "src")
    for root, _dirs, files in os.walk(src_root):
        rel = os.path.relpath(root, src_root)
        os.makedirs(os.path.join(dst, rel), exist_ok=True)
        for f in files:
            if not f.endswith(".js"):
                continue
            text = open(os.path.join(root, f), encoding="utf-8").read()

            def fix(m, root=root):
                spec = m.group(3)
                target = os.path.normpath(os.path.join(root, spec))
                if os.path.isfile(target + ".js"):
                    spec += ".js"
                elif os.path.isdir(target):
                    spec = spec.rstrip("/") + "/index.js"
                return m.group(1) + m.group(2) + spec + m.group(2)

            # `import 'util';` (src/dt/complex_array.js:20) names Node's built-in module for a side effect nothing uses;
            # bundlers resolve it to a polyfill, a bare ES-module loader cannot
            text = re.sub(r"^import 'util';", "/* import 'util'; */", text, flags=re.M)
            # src/rand/alea_rng.js:24 imports a name `RNG` that src/_test_data_generators.js does not export and the file
            # never uses; webpack only warns about it, a strict ES-module linker refuses the file
            if f == "alea_rng.js":
                text = text.replace("import { RNG } from '../_test_data_generators';", "/* import { RNG } ... (dead import) */")
            # QJSEngine 6.6.3 crashes (SIGSEGV while defining the class) on `static get name()` next to a constructor in
            # a class that extends Function (src/nd_array.js:130), and on redefining that property afterwards; the accessor
            # is cosmetic (the class's display name), so the staged copy goes without it
            if rel == "." and f == "nd_array.js":
                text = text.replace("static get name() { return 'nd.Array'; }", "")
            open(os.path.join(dst, rel, f), "w", encoding="utf-8").write(_IMPORT.sub(fix, text))
    return dst


_PRELUDE = r"""
// QJSEngine 6.6.3: %TypedArray%.from(generator object) yields zeros (Sets and arrays are fine); route iterables
// through Array.from, which is what the specification's IterableToList step does (used by src/nd_array.js:63)
(function(){
  var TA = Object.getPrototypeOf(Int8Array), from0 = TA.from;
  Object.defineProperty(TA, 'from', {configurable: true, writable: true, value: function(src, fn, self) {
    if (src != null && typeof src[Symbol.iterator] === 'function' && !Array.isArray(src) && !ArrayBuffer.isView(src)) src = Array.from(src);
    return fn === undefined ? from0.call(this, src) : from0.call(this, src, fn, self);
  }});
})();
// QJSEngine 6.6.3 has no %TypedArray%.prototype.sort (src/la/_svd_jac_utils.js sorts an Int32Array of indices).
// ECMAScript 2019 requires a stable sort (V8: TimSort); a stable merge sort gives the same order for any consistent
// comparator, and for an inconsistent one (NaN singular values) no engine's order is specified.
(function(){
  var TA = Object.getPrototypeOf(Int8Array);
  if (typeof TA.prototype.sort === 'function') return;
  Object.defineProperty(TA.prototype, 'sort', {configurable: true, writable: true, value: function(cmp) {
    if (cmp === undefined) cmp = function(a, b) { return a < b ? -1 : a > b ? 1 : (a === 0 && b === 0) ? (Object.is(a, -0) ? (Object.is(b, -0) ? 0 : -1) : (Object.is(b, -0) ? 1 : 0)) : (a !== a ? (b !== b ? 0 : 1) : (b !== b ? -1 : 0)); };
    var n = this.length, a = Array.from(this), b = new Array(n);
    for (var w = 1; w < n; w *= 2) {
      for (var lo = 0; lo < n; lo += 2*w) {
        var mid = Math.min(lo + w, n), hi = Math.min(lo + 2*w, n), i = lo, j = mid, k = lo;
        while (i < mid && j < hi) b[k++] = (cmp(a[j], a[i]) < 0) ? a[j++] : a[i++];
        while (i < mid) b[k++] = a[i++];
        while (j < hi) b[k++] = a[j++];
      }
      var t = a; a = b; b = t;
    }
    for (var q = 0; q < n; q++) this[q] = a[q];
    return this;
  }});
})();
var __hex = '0123456789abcdef';
function __to_hex(arr) {            // raw little-endian bytes of a typed array as a hex string
  var b = new Uint8Array(arr.buffer, arr.byteOffset, arr.byteLength), s = new Array(b.length);
  for (var i = 0; i < b.length; i++) s[i] = __hex[b[i] >> 4] + __hex[b[i] & 15];
  return s.join('');
}
function __from_hex(s, Type) {
  var n = s.length >> 1, b = new Uint8Array(n);
  for (var i = 0; i < n; i++) b[i] = parseInt(s.substr(2*i, 2), 16);
  return new Type(b.buffer);
}
function __pack(x) {                // NDArray / typed array / array / scalar -> JSON-able
  if (x === undefined) return {u: 1};
  if (x === null || typeof x === 'number' || typeof x === 'string' || typeof x === 'boolean') return {v: x, f: (typeof x === 'number') ? __to_hex(Float64Array.of(x)) : undefined};
  if (x instanceof Float64Array) return {t: 'f8', h: __to_hex(x)};
  if (x instanceof Float32Array) return {t: 'f4', h: __to_hex(x)};
  if (x instanceof Int32Array)   return {t: 'i4', h: __to_hex(x)};
  if (x instanceof Uint32Array)  return {t: 'u4', h: __to_hex(x)};
  if (x instanceof Uint8Array)   return {t: 'u1', h: __to_hex(x)};
  if (x.shape !== undefined && x.data !== undefined) {
    var d = x.data;
    if (!(d instanceof Float64Array || d instanceof Int32Array || d instanceof Float32Array)) return {nd: Array.from(x.shape), dt: x.dtype, a: Array.from(d).map(__pack)};
    return {nd: Array.from(x.shape), dt: x.dtype, d: __pack(d)};
  }
  if (Array.isArray(x)) return {l: x.map(__pack)};
  var o = {}; for (var k in x) o[k] = __pack(x[k]); return {o: o};
}
function __call(f) {                // f: () => value; exceptions become {err: message}
  try { return JSON.stringify(__pack(f())); }
  catch (e) { return JSON.stringify({err: '' + (e && e.message !== undefined ? e.message : e), name: e && e.name, stack: '' + (e && e.stack)}); }
}
"""


class JSError(Exception):
    pass


class Engine:
    """One QJSEngine with the reference's modules importable by their path below src/ (e.g. 'la/qr.js')."""

    def __init__(self):
        if not available():
            raise RuntimeError("QJSEngine or the reference checkout is not present")
        self.tmp = tempfile.mkdtemp(prefix="nd4b_jsref_")
        stub = os.path.join(self.tmp, "stub.c")
        with open(stub, "w") as f:
            f.write("#include <stdlib.h>\n" + "".join("void *%s(void){abort();}\n" % s for s in _GLIB_SYMS))
        for so in ("libglib-2.0.so.0", "libgthread-2.0.so.0"):
            subprocess.check_call(["gcc", "-shared", "-fPIC", "-Wl,-soname," + so, "-o", os.path.join(self.tmp, so), stub])
        os.environ["QT_NO_GLIB"] = "1"
        os.environ.setdefault("QT_LOGGING_RULES", "qt.qml.compiler=false")
        G = _c.RTLD_GLOBAL
        for so in ("libglib-2.0.so.0", "libgthread-2.0.so.0"):
            _c.CDLL(os.path.join(self.tmp, so), mode=G)
        d = _qt_dir()
        # the ICU / zstd / libstdc++ copies next to the Qt libraries are found through Qt's $ORIGIN runpath
        self.core = _c.CDLL(os.path.join(d, "libQt6Core.so.6"), mode=G)
        _c.CDLL(os.path.join(d, "libQt6Network.so.6"), mode=G)
        self.qml = _c.CDLL(os.path.join(d, "libQt6Qml.so.6"), mode=G)
        P, core, qml = _c.c_void_p, self.core, self.qml

        def fn(lib, name, argtypes, restype=None):
            f = getattr(lib, name)
            f.argtypes, f.restype = argtypes, restype
            return f

        self._fromUtf8 = fn(core, "_ZN7QString8fromUtf8E14QByteArrayView", [P, _c.c_longlong, _c.c_char_p], P)
        self._evaluate = fn(qml, "_ZN9QJSEngine8evaluateERK7QStringS2_iP5QListIS0_E", [P, P, P, P, _c.c_int, P], P)
        self._import = fn(qml, "_ZN9QJSEngine12importModuleERK7QString", [P, P, P], P)
        self._global = fn(qml, "_ZNK9QJSEngine12globalObjectEv", [P, P], P)
        self._setprop = fn(qml, "_ZN8QJSValue11setPropertyERK7QStringRKS_", [P, P, P])
        self._tostr = fn(qml, "_ZNK8QJSValue8toStringEv", [P, P], P)
        self._iserr = fn(qml, "_ZNK8QJSValue7isErrorEv", [P], _c.c_bool)
        self._valdtor = fn(qml, "_ZN8QJSValueD1Ev", [P])
        self._gc = fn(qml, "_ZN9QJSEngine14collectGarbageEv", [P])
        self._keep = []
        argc = _c.c_int(1)
        argv = (_c.c_char_p * 2)(b"nd4b_jsref", None)
        app = _c.create_string_buffer(256)
        fn(core, "_ZN16QCoreApplicationC1ERiPPci", [P, P, P, _c.c_int])(app, _c.byref(argc), argv, 0x060603)
        self._keep += [argc, argv, app]
        self.eng = _c.create_string_buffer(512)
        fn(qml, "_ZN9QJSEngineC1Ev", [P])(self.eng)
        self.src = stage_sources(os.path.join(self.tmp, "src"))
        self._mods = {}
        self.run(_PRELUDE + "\n'ok'")

    # -- plumbing ---------------------------------------------------------------------------------------------
    def _qstr(self, s):
        b = s.encode("utf-8")
        buf = _c.create_string_buffer(24)            # QString = {d, ptr, size}; leaked on purpose (no inline dtor to call)
        self._fromUtf8(buf, len(b), b)
        self._keep_last = (b,)
        return buf

    def _str_of(self, val):
        s = _c.create_string_buffer(24)
        self._tostr(s, val)
        _d, p, n = _c.cast(s, _c.POINTER(_c.c_void_p * 3)).contents
        return _c.string_at(p, n * 2).decode("utf-16-le") if n else ""

    def run(self, code, name="nd4b.js"):
        """Evaluate a script; returns the completion value as a string; raises JSError on an uncaught exception."""
        v = _c.create_string_buffer(16)
        self._evaluate(v, self.eng, self._qstr(code), self._qstr(name), 1, None)
        try:
            s = self._str_of(v)
            if self._iserr(v):
                raise JSError(s)
            return s
        finally:
            self._valdtor(v)

    def module(self, rel, as_name=None):
        """Import src/<rel> as an ES module and bind its namespace object to a global variable (returned name)."""
        if rel in self._mods:
            return self._mods[rel]
        name = as_name or "M_" + re.sub(r"\W", "_", rel[:-3] if rel.endswith(".js") else rel)
        v = _c.create_string_buffer(16)
        self._import(v, self.eng, self._qstr(os.path.join(self.src, rel)))
        if self._iserr(v):
            msg = self._str_of(v)
            self._valdtor(v)
            raise JSError("import %s: %s" % (rel, msg))
        g = _c.create_string_buffer(16)
        self._global(g, self.eng)
        self._setprop(g, self._qstr(name), v)
        self._valdtor(g)
        self._valdtor(v)
        self._mods[rel] = name
        return name

    def gc(self):
        self._gc(self.eng)

    # -- value marshalling --------------------------------------------------------------------------------------
    def call(self, expr):
        """Evaluate `expr` (a JS expression) and return it as numpy arrays / Python values, bit-exact for typed arrays."""
        import json
        out = json.loads(self.run("__call(function(){ return (%s); })" % expr))
        return _unpack(out)

    def close(self):
        shutil.rmtree(self.tmp, ignore_errors=True)


def _unpack(x):
    import numpy as np
    if "err" in x:
        e = JSError(x["err"])
        e.js_name, e.js_stack = x.get("name"), x.get("stack")
        raise e
    if "u" in x:
        return None
    if "v" in x:
        if x.get("f"):
            return float(np.frombuffer(bytes.fromhex(x["f"]), "<f8")[0])
        return x["v"]
    if "t" in x:
        return np.frombuffer(bytes.fromhex(x["h"]), {"f8": "<f8", "f4": "<f4", "i4": "<i4", "u4": "<u4", "u1": "u1"}[x["t"]]).copy()
    if "nd" in x:
        if "a" in x:
            return np.array([_unpack(v) for v in x["a"]], dtype=object).reshape(x["nd"])
        return _unpack(x["d"]).reshape(x["nd"])
    if "l" in x:
        return [_unpack(v) for v in x["l"]]
    if "o" in x:
        return {k: _unpack(v) for k, v in x["o"].items()}
    raise ValueError(x)


def js_f64(a):
    """JS expression for a Float64Array holding exactly the bytes of `a`."""
    import numpy as np
    return "__from_hex('%s', Float64Array)" % np.ascontiguousarray(a, dtype="<f8").tobytes().hex()


def js_nd(mod_nd_array, a):
    """JS expression constructing the reference's NDArray (src/nd_array.js:135) from a numpy float64 array."""
    import numpy as np
    a = np.asarray(a, dtype=np.float64)
    return "new %s.NDArray(Int32Array.of(%s), %s)" % (mod_nd_array, ",".join(str(int(s)) for s in a.shape), js_f64(a))


_engine = None


def engine():
    global _engine
    if _engine is None:
        _engine = Engine()
    return _engine
