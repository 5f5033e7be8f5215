// nd4b_napi.cc — Node.js N-API addon: a thin, synchronous binding of the C ABI in include/nd4b.h.
//
// Every export takes typed arrays owned by the JS caller (Float64Array data, Int32Array shapes), reads
// the inputs only for the duration of the call, fills caller-allocated outputs and retains nothing —
// the ownership/threading contract of the reference's nd.la functions.  Errors are
// thrown as JS Error objects carrying the reference's own message text (nd4b_last_error()).
// There is no CPU fallback: without a usable B200 every call throws.
//
// The image has no Node.js; this file is compile-checked against node_api_min.h (tests/test_addon.py).
#ifdef ND4B_HAVE_NODE_API
#include <node_api.h>
#else
#include "node_api_min.h"
#endif
#include "../../include/nd4b.h"

#include <stdint.h>
#include <string.h>

namespace {

struct F64 { double* p; size_t n; };
struct I32 { int32_t* p; size_t n; };

bool get_f64(napi_env env, napi_value v, F64* out) {
  napi_typedarray_type t; void* data; size_t len;
  if (napi_get_typedarray_info(env, v, &t, &len, &data, nullptr, nullptr) != napi_ok || t != napi_float64_array) {
    napi_throw_error(env, nullptr, "nd4b: expected a Float64Array");
    return false;
  }
  out->p = static_cast<double*>(data); out->n = len;
  return true;
}
bool get_i32(napi_env env, napi_value v, I32* out) {
  napi_typedarray_type t; void* data; size_t len;
  if (napi_get_typedarray_info(env, v, &t, &len, &data, nullptr, nullptr) != napi_ok || t != napi_int32_array) {
    napi_throw_error(env, nullptr, "nd4b: expected an Int32Array");
    return false;
  }
  out->p = static_cast<int32_t*>(data); out->n = len;
  return true;
}
bool get_int(napi_env env, napi_value v, int64_t* out) {
  if (napi_get_value_int64(env, v, out) != napi_ok) { napi_throw_error(env, nullptr, "nd4b: expected an integer"); return false; }
  return true;
}
napi_value undefined(napi_env env) { napi_value u; napi_get_undefined(env, &u); return u; }
napi_value fail(napi_env env) { napi_throw_error(env, nullptr, nd4b_last_error()); return nullptr; }
int64_t prod(const I32& s) { int64_t p = 1; for (size_t i = 0; i < s.n; i++) p *= s.p[i]; return p; }

// matmulShape(aShape:Int32Array, bShape:Int32Array, cShape:Int32Array) -> ndim
napi_value MatmulShape(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value a[3];
  napi_get_cb_info(env, info, &argc, a, nullptr, nullptr);
  I32 as, bs, cs;
  if (argc < 3 || !get_i32(env, a[0], &as) || !get_i32(env, a[1], &bs) || !get_i32(env, a[2], &cs)) return nullptr;
  if (cs.n < (as.n > bs.n ? as.n : bs.n)) { napi_throw_error(env, nullptr, "nd4b: cShape too short"); return nullptr; }
  int nd = 0;
  if (nd4b_matmul_shape(as.p, (int)as.n, bs.p, (int)bs.n, cs.p, &nd)) return fail(env);
  napi_value r; napi_create_int32(env, nd, &r); return r;
}

// matmul(a:Float64Array, aShape, b:Float64Array, bShape, c:Float64Array, cShape)
napi_value Matmul(napi_env env, napi_callback_info info) {
  size_t argc = 6; napi_value v[6];
  napi_get_cb_info(env, info, &argc, v, nullptr, nullptr);
  F64 a, b, c; I32 as, bs, cs;
  if (argc < 6 || !get_f64(env, v[0], &a) || !get_i32(env, v[1], &as) || !get_f64(env, v[2], &b) ||
      !get_i32(env, v[3], &bs) || !get_f64(env, v[4], &c) || !get_i32(env, v[5], &cs)) return nullptr;
  if ((int64_t)a.n != prod(as) || (int64_t)b.n != prod(bs) || (int64_t)c.n != prod(cs)) {
    napi_throw_error(env, nullptr, "nd4b: data length does not match shape"); return nullptr;
  }
  if (nd4b_matmul_f64(a.p, as.p, (int)as.n, b.p, bs.p, (int)bs.n, c.p, cs.p, (int)cs.n)) return fail(env);
  return undefined(env);
}

// matmulPlan(mats:Array<Float64Array>, shapes:Array<Int32Array>, plan:Int32Array, c:Float64Array, cShape:Int32Array)
// — the chain product of nd.la.matmul with its intermediates kept in HBM (nd4b_matmul_plan_f64)
napi_value MatmulPlan(napi_env env, napi_callback_info info) {
  size_t argc = 5; napi_value v[5];
  napi_get_cb_info(env, info, &argc, v, nullptr, nullptr);
  uint32_t n = 0, ns = 0;
  I32 plan, cs; F64 c;
  if (argc < 5 || napi_get_array_length(env, v[0], &n) != napi_ok || napi_get_array_length(env, v[1], &ns) != napi_ok || n != ns ||
      n < 1 || n > 64 || !get_i32(env, v[2], &plan) || !get_f64(env, v[3], &c) || !get_i32(env, v[4], &cs)) {
    napi_throw_error(env, nullptr, "nd4b: matmulPlan(mats, shapes, plan, c, cShape)"); return nullptr;
  }
  const double* mats[64]; const int32_t* shapes[64]; int ndims[64];
  for (uint32_t i = 0; i < n; i++) {
    napi_value m, s; F64 mf; I32 sf;
    if (napi_get_element(env, v[0], i, &m) != napi_ok || napi_get_element(env, v[1], i, &s) != napi_ok ||
        !get_f64(env, m, &mf) || !get_i32(env, s, &sf)) return nullptr;
    if ((int64_t)mf.n != prod(sf)) { napi_throw_error(env, nullptr, "nd4b: data length does not match shape"); return nullptr; }
    mats[i] = mf.p; shapes[i] = sf.p; ndims[i] = (int)sf.n;
  }
  if ((int64_t)c.n != prod(cs)) { napi_throw_error(env, nullptr, "nd4b: data length does not match shape"); return nullptr; }
  if (nd4b_matmul_plan_f64((int)n, mats, shapes, ndims, plan.p, (int)plan.n, c.p, cs.p, (int)cs.n)) return fail(env);
  return undefined(env);
}

// cholesky(S:Float64Array, L:Float64Array, batch, n)
napi_value Cholesky(napi_env env, napi_callback_info info) {
  size_t argc = 4; napi_value v[4];
  napi_get_cb_info(env, info, &argc, v, nullptr, nullptr);
  F64 s, l; int64_t batch, n;
  if (argc < 4 || !get_f64(env, v[0], &s) || !get_f64(env, v[1], &l) || !get_int(env, v[2], &batch) || !get_int(env, v[3], &n)) return nullptr;
  if ((int64_t)s.n != batch * n * n || l.n != s.n) { napi_throw_error(env, nullptr, "nd4b: data length does not match shape"); return nullptr; }
  int64_t bad = -1;
  if (nd4b_cholesky_f64(s.p, l.p, batch, (int)n, &bad)) return fail(env);
  return undefined(env);
}

// qr(A, Q, R, batch, rows, cols)
napi_value Qr(napi_env env, napi_callback_info info) {
  size_t argc = 6; napi_value v[6];
  napi_get_cb_info(env, info, &argc, v, nullptr, nullptr);
  F64 a, q, r; int64_t batch, rows, cols;
  if (argc < 6 || !get_f64(env, v[0], &a) || !get_f64(env, v[1], &q) || !get_f64(env, v[2], &r) ||
      !get_int(env, v[3], &batch) || !get_int(env, v[4], &rows) || !get_int(env, v[5], &cols)) return nullptr;
  const int64_t l = rows < cols ? rows : cols;
  if ((int64_t)a.n != batch * rows * cols || (int64_t)q.n != batch * rows * l || (int64_t)r.n != batch * l * cols) {
    napi_throw_error(env, nullptr, "nd4b: data length does not match shape"); return nullptr;
  }
  if (nd4b_qr_f64(a.p, q.p, r.p, batch, (int)rows, (int)cols)) return fail(env);
  return undefined(env);
}

// qrInplace(A, Y, R, QtY, batch, M, N, L)
napi_value QrInplace(napi_env env, napi_callback_info info) {
  size_t argc = 8; napi_value v[8];
  napi_get_cb_info(env, info, &argc, v, nullptr, nullptr);
  F64 a, y, r, q; int64_t batch, m, n, l;
  if (argc < 8 || !get_f64(env, v[0], &a) || !get_f64(env, v[1], &y) || !get_f64(env, v[2], &r) || !get_f64(env, v[3], &q) ||
      !get_int(env, v[4], &batch) || !get_int(env, v[5], &m) || !get_int(env, v[6], &n) || !get_int(env, v[7], &l)) return nullptr;
  if ((int64_t)a.n != batch * m * n || r.n != a.n || (int64_t)y.n != batch * m * l || q.n != y.n) {
    napi_throw_error(env, nullptr, "nd4b: data length does not match shape"); return nullptr;
  }
  if (nd4b_qr_inplace_f64(a.p, y.p, r.p, q.p, batch, (int)m, (int)n, (int)l)) return fail(env);
  return undefined(env);
}

// qrLstsq(Q, R, Y, X, batch, N, M, I, J) — the fused form for thin factors with one common batch
napi_value QrLstsq(napi_env env, napi_callback_info info) {
  size_t argc = 9; napi_value v[9];
  napi_get_cb_info(env, info, &argc, v, nullptr, nullptr);
  F64 q, r, y, x; int64_t batch, n, m, i, j;
  if (argc < 9 || !get_f64(env, v[0], &q) || !get_f64(env, v[1], &r) || !get_f64(env, v[2], &y) || !get_f64(env, v[3], &x) ||
      !get_int(env, v[4], &batch) || !get_int(env, v[5], &n) || !get_int(env, v[6], &m) || !get_int(env, v[7], &i) || !get_int(env, v[8], &j)) return nullptr;
  if ((int64_t)q.n != batch * n * m || (int64_t)r.n != batch * m * i || (int64_t)y.n != batch * n * j || (int64_t)x.n != batch * i * j) {
    napi_throw_error(env, nullptr, "nd4b: data length does not match shape"); return nullptr;
  }
  if (nd4b_qr_lstsq_f64(q.p, r.p, y.p, x.p, batch, (int)n, (int)m, (int)i, (int)j)) return fail(env);
  return undefined(env);
}

// svdJac1(A, U, sv, V, batch, rows, cols) -> sweeps
napi_value SvdJac1(napi_env env, napi_callback_info info) {
  size_t argc = 7; napi_value v[7];
  napi_get_cb_info(env, info, &argc, v, nullptr, nullptr);
  F64 a, u, s, vt; int64_t batch, rows, cols;
  if (argc < 7 || !get_f64(env, v[0], &a) || !get_f64(env, v[1], &u) || !get_f64(env, v[2], &s) || !get_f64(env, v[3], &vt) ||
      !get_int(env, v[4], &batch) || !get_int(env, v[5], &rows) || !get_int(env, v[6], &cols)) return nullptr;
  const int64_t l = rows < cols ? rows : cols;
  if ((int64_t)a.n != batch * rows * cols || (int64_t)u.n != batch * rows * l || (int64_t)s.n != batch * l ||
      (int64_t)vt.n != batch * l * cols) {
    napi_throw_error(env, nullptr, "nd4b: data length does not match shape"); return nullptr;
  }
  int sweeps = 0;
  if (nd4b_svd_jac1_f64(a.p, u.p, s.p, vt.p, batch, (int)rows, (int)cols, &sweeps)) return fail(env);
  napi_value r; napi_create_int32(env, sweeps, &r); return r;
}

// svdRank(sv:Float64Array, rank:Int32Array, n) — nd4b_svd_rank_f64 (svd.js:31-58)
napi_value SvdRank(napi_env env, napi_callback_info info) {
  size_t argc = 3; napi_value v[3];
  napi_get_cb_info(env, info, &argc, v, nullptr, nullptr);
  F64 sv; I32 rank; int64_t n;
  if (argc < 3 || !get_f64(env, v[0], &sv) || !get_i32(env, v[1], &rank) || !get_int(env, v[2], &n)) return nullptr;
  if (n < 1 || (int64_t)sv.n != (int64_t)rank.n * n) { napi_throw_error(env, nullptr, "nd4b: data length does not match shape"); return nullptr; }
  if (nd4b_svd_rank_f64(sv.p, rank.p, (int64_t)rank.n, (int)n)) return fail(env);
  return undefined(env);
}

// svdLstsqShape(uShape, svShape, vShape, yShape, xShape) -> ndim: the reference's checks and texts (svd.js:112-147)
napi_value SvdLstsqShape(napi_env env, napi_callback_info info) {
  size_t argc = 5; napi_value v[5];
  napi_get_cb_info(env, info, &argc, v, nullptr, nullptr);
  I32 us, ss, vs, ys, xs;
  if (argc < 5 || !get_i32(env, v[0], &us) || !get_i32(env, v[1], &ss) || !get_i32(env, v[2], &vs) || !get_i32(env, v[3], &ys) ||
      !get_i32(env, v[4], &xs)) return nullptr;
  size_t need = us.n > ss.n + 1 ? us.n : ss.n + 1;
  if (vs.n > need) need = vs.n;
  if (ys.n > need) need = ys.n;
  if (xs.n < need) { napi_throw_error(env, nullptr, "nd4b: xShape too short"); return nullptr; }
  int nd = 0;
  if (nd4b_svd_lstsq_shape(us.p, (int)us.n, ss.p, (int)ss.n, vs.p, (int)vs.n, ys.p, (int)ys.n, xs.p, &nd)) return fail(env);
  napi_value r; napi_create_int32(env, nd, &r); return r;
}

// svdLstsq(U, uShape, sv, svShape, V, vShape, Y, yShape, X, xShape) — nd4b_svd_lstsq_f64 (svd.js:103-226)
napi_value SvdLstsq(napi_env env, napi_callback_info info) {
  size_t argc = 10; napi_value v[10];
  napi_get_cb_info(env, info, &argc, v, nullptr, nullptr);
  F64 u, s, vt, y, x; I32 us, ss, vs, ys, xs;
  if (argc < 10 || !get_f64(env, v[0], &u) || !get_i32(env, v[1], &us) || !get_f64(env, v[2], &s) || !get_i32(env, v[3], &ss) ||
      !get_f64(env, v[4], &vt) || !get_i32(env, v[5], &vs) || !get_f64(env, v[6], &y) || !get_i32(env, v[7], &ys) ||
      !get_f64(env, v[8], &x) || !get_i32(env, v[9], &xs)) return nullptr;
  if ((int64_t)u.n != prod(us) || (int64_t)s.n != prod(ss) || (int64_t)vt.n != prod(vs) || (int64_t)y.n != prod(ys) || (int64_t)x.n != prod(xs)) {
    napi_throw_error(env, nullptr, "nd4b: data length does not match shape"); return nullptr;
  }
  if (nd4b_svd_lstsq_f64(u.p, us.p, (int)us.n, s.p, ss.p, (int)ss.n, vt.p, vs.p, (int)vs.n, y.p, ys.p, (int)ys.n, x.p, xs.p, (int)xs.n))
    return fail(env);
  return undefined(env);
}

// triSolve(op, T, tShape, Y, yShape, X, xShape)   op 0 tril_solve, 1 triu_solve, 2 cholesky_solve
napi_value TriSolve(napi_env env, napi_callback_info info) {
  size_t argc = 7; napi_value v[7];
  napi_get_cb_info(env, info, &argc, v, nullptr, nullptr);
  int64_t op; F64 t, y, x; I32 ts, ys, xs;
  if (argc < 7 || !get_int(env, v[0], &op) || !get_f64(env, v[1], &t) || !get_i32(env, v[2], &ts) || !get_f64(env, v[3], &y) ||
      !get_i32(env, v[4], &ys) || !get_f64(env, v[5], &x) || !get_i32(env, v[6], &xs)) return nullptr;
  if ((int64_t)t.n != prod(ts) || (int64_t)y.n != prod(ys) || (int64_t)x.n != prod(xs)) {
    napi_throw_error(env, nullptr, "nd4b: data length does not match shape"); return nullptr;
  }
  if (nd4b_tri_solve_f64((int)op, t.p, ts.p, (int)ts.n, y.p, ys.p, (int)ys.n, x.p, xs.p, (int)xs.n)) return fail(env);
  return undefined(env);
}

void free_pinned(napi_env, void* data, void*) { nd4b_host_free(data); }

// pinnedFloat64Array(length) -> Float64Array backed by page-locked memory (skips the staging copy)
napi_value PinnedFloat64Array(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value v[1];
  napi_get_cb_info(env, info, &argc, v, nullptr, nullptr);
  int64_t n;
  if (argc < 1 || !get_int(env, v[0], &n) || n < 0) return nullptr;
  void* p = nd4b_host_alloc((size_t)n * sizeof(double));
  if (!p) return fail(env);
  memset(p, 0, (size_t)n * sizeof(double));
  napi_value ab, ta;
  if (napi_create_external_arraybuffer(env, p, (size_t)n * sizeof(double), free_pinned, nullptr, &ab) != napi_ok) {
    nd4b_host_free(p); napi_throw_error(env, nullptr, "nd4b: external ArrayBuffers are not allowed in this runtime"); return nullptr;
  }
  napi_create_typedarray(env, napi_float64_array, (size_t)n, ab, 0, &ta);
  return ta;
}

// init(devices:Int32Array) ; deviceCount() ; stats()
napi_value Init(napi_env env, napi_callback_info info) {
  size_t argc = 1; napi_value v[1];
  napi_get_cb_info(env, info, &argc, v, nullptr, nullptr);
  I32 d{nullptr, 0};
  if (argc >= 1 && !get_i32(env, v[0], &d)) return nullptr;
  if (nd4b_init(d.p, (int)d.n)) return fail(env);
  return undefined(env);
}
napi_value DeviceCount(napi_env env, napi_callback_info) { napi_value r; napi_create_int32(env, nd4b_device_count(), &r); return r; }
napi_value Stats(napi_env env, napi_callback_info) {
  nd4b_stats s;
  if (nd4b_get_stats(&s)) return fail(env);
  napi_value o, x; napi_create_object(env, &o);
  napi_create_double(env, (double)s.calls, &x); napi_set_named_property(env, o, "calls", x);
  napi_create_double(env, (double)s.kernel_launches, &x); napi_set_named_property(env, o, "kernelLaunches", x);
  napi_create_double(env, (double)s.h2d_bytes, &x); napi_set_named_property(env, o, "h2dBytes", x);
  napi_create_double(env, (double)s.d2h_bytes, &x); napi_set_named_property(env, o, "d2hBytes", x);
  napi_create_int32(env, s.last_sweeps, &x); napi_set_named_property(env, o, "lastSweeps", x);
  napi_create_int32(env, s.n_devices, &x); napi_set_named_property(env, o, "devices", x);
  return o;
}

napi_value RegisterAll(napi_env env, napi_value exports) {
  const napi_property_descriptor props[] = {
      {"matmulShape", nullptr, MatmulShape, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"matmul", nullptr, Matmul, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"matmulPlan", nullptr, MatmulPlan, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"cholesky", nullptr, Cholesky, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"qr", nullptr, Qr, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"qrLstsq", nullptr, QrLstsq, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"qrInplace", nullptr, QrInplace, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"triSolve", nullptr, TriSolve, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"svdJac1", nullptr, SvdJac1, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"svdRank", nullptr, SvdRank, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"svdLstsqShape", nullptr, SvdLstsqShape, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"svdLstsq", nullptr, SvdLstsq, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"pinnedFloat64Array", nullptr, PinnedFloat64Array, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"init", nullptr, Init, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"deviceCount", nullptr, DeviceCount, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"stats", nullptr, Stats, nullptr, nullptr, nullptr, napi_default, nullptr},
  };
  napi_define_properties(env, exports, sizeof props / sizeof props[0], props);
  return exports;
}

napi_module g_module = {1, 0, __FILE__, RegisterAll, "nd4b", nullptr, {nullptr, nullptr, nullptr, nullptr}};
__attribute__((constructor)) void register_nd4b() { napi_module_register(&g_module); }

}  // namespace
