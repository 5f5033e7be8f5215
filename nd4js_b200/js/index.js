'use strict';
// nd4b JS shim — drop-in replacements for the nd4js functions on the batched dense-LA hot path.
// Same names, signatures, result shapes and error texts as the reference:
//   matmul2 / matmul       nd4js src/la/matmul.js:91-147, :150-236
//   cholesky_decomp        nd4js src/la/cholesky.js:50-72
//   qr_decomp              nd4js src/la/qr.js:80-145
//   svd_jac_1sided         contract of nd4js src/la/svd_jac_2sided.js:30-144 (new export)
// asarray(), dtype upcasts and the matrix-chain ordering stay in JS exactly as in the reference;
// the arithmetic is done by the N-API addon (addon/nd4b_napi.cc -> libnd4b.so -> CUDA on B200).
// Float64 only; there is no CPU fallback: other dtypes and a missing GPU throw.
const nd = require('nd4js');
const addon = require('../addon/build/Release/nd4b.node');
const {NDArray, asarray} = nd;

// Result storage: large results are page-locked (addon.pinnedFloat64Array -> nd4b_host_alloc, a cached pool), so the D2H
// copies of the call are DMA'd straight into the array the caller receives and a later nd.la call on it is DMA'd
// straight out of it; small ones are ordinary typed arrays.
const PINNED_MIN_LENGTH = (1 << 20) / 8;
const alloc = n => n >= PINNED_MIN_LENGTH ? addon.pinnedFloat64Array(n) : new Float64Array(n);

function f64(a, who) {
  if (a.dtype === 'float64') return a.data;
  if (a.dtype === 'int32') return Float64Array.from(a.data); // upcast as qr.js:93 / cholesky.js:58 do
  throw new Error(`${who}: dtype '${a.dtype}' is not supported by the GPU path (float64 only).`);
}

function matmul2(a, b) {
  a = asarray(a); b = asarray(b);
  if (a.ndim < 2) throw new Error('A must be at least 2D.');
  if (b.ndim < 2) throw new Error('B must be at least 2D.');
  const shape = new Int32Array(Math.max(a.ndim, b.ndim));
  addon.matmulShape(a.shape, b.shape, shape); // throws the reference's texts on mismatch
  const c = alloc(shape.reduce((m, n) => m * n, 1));
  addon.matmul(f64(a, 'matmul2'), a.shape, f64(b, 'matmul2'), b.shape, c, shape);
  return new NDArray(shape, c);
}

function matmul(...matrices) {
  matrices = matrices.map(m => asarray(m));
  if (matrices.length === 1) return matrices[0];
  if (matrices.length === 2) return matmul2(...matrices);
  // chain ordering by broadcast-aware flop counts, as matmul.js:159-235
  const nOps = (sa, sb) => {
    const ndim = Math.max(sa.length, sb.length), shape = new Int32Array(ndim).fill(1);
    if (sb[sb.length - 2] !== sa[sa.length - 1]) throw new Error('Shape mismatch.');
    shape[ndim - 2] = sa[sa.length - 2]; shape[ndim - 1] = sb[sb.length - 1];
    for (const shp of [sa, sb])
      for (let i = ndim - 2, j = shp.length - 2; i-- > 0 && j-- > 0;)
        if (shape[i] === 1) shape[i] = shp[j];
        else if (shape[i] !== shp[j] && shp[j] !== 1) throw new Error('Shapes are not broadcast-compatible.');
    return [shape.reduce((x, y) => x * y, 1) * sa[sa.length - 1], shape];
  };
  const n = matrices.length, op = Array.from({length: n}, () => []);
  for (let i = 0; i < n; i++) op[i][i] = [0, matrices[i].shape];
  for (let len = 2; len <= n; len++)
    for (let i = 0; i <= n - len; i++) {
      let best = Infinity, bestShape;
      for (let j = 1; j < len; j++) {
        const [lf, ls] = op[i][i + j - 1], [rf, rs] = op[i + j][i + len - 1];
        let [f, s] = nOps(ls, rs); f += lf + rf;
        if (f < best) { best = f; bestShape = s; }
      }
      if (bestShape === undefined) throw new Error('Integer overflow (too many FLOPs).');
      op[i][i + len - 1] = [best, bestShape];
    }
  // postfix plan of the optimal parenthesisation: i pushes operand i, -1 multiplies the two topmost items; the products
  // then run with every intermediate kept in HBM (nd4b_matmul_plan_f64) instead of one host round trip per matmul2
  const plan = [];
  const product = (from, to) => {
    if (from === to) { plan.push(from); return; }
    let best = Infinity, idx;
    for (let i = from; i < to; i++) {
      const [lf, ls] = op[from][i], [rf, rs] = op[i + 1][to];
      let [f] = nOps(ls, rs); f += lf + rf;
      if (f < best) { best = f; idx = i; }
    }
    product(from, idx); product(idx + 1, to); plan.push(-1);
  };
  product(0, n - 1);
  for (const [i, m] of matrices.entries())
    if (m.ndim < 2) throw new Error(i === 0 ? 'A must be at least 2D.' : 'B must be at least 2D.');
  const shape = Int32Array.from(op[0][n - 1][1]), c = alloc(shape.reduce((x, y) => x * y, 1));
  addon.matmulPlan(matrices.map(m => f64(m, 'matmul')), matrices.map(m => Int32Array.from(m.shape)), Int32Array.from(plan), c, shape);
  return new NDArray(shape, c);
}

function cholesky_decomp(S) {
  S = asarray(S);
  const [N, M] = S.shape.slice(-2);
  if (N !== M) throw new Error('Last two dimensions must be quadratic.');
  const s = f64(S, 'cholesky_decomp'), L = alloc(s.length);
  addon.cholesky(s, L, s.length / (N * N), N); // throws 'Matrix contains NaNs or is (near) singular.'
  return new NDArray(S.shape, L);
}

function qr_decomp(A) {
  A = asarray(A);
  if (A.ndim < 2) throw new Error('qr_decomp(A): A.ndim must be at least 2.');
  const [N, M] = A.shape.slice(-2), L = Math.min(N, M), a = f64(A, 'qr_decomp'), batch = a.length / (N * M);
  const qShape = Int32Array.from(A.shape), rShape = Int32Array.from(A.shape);
  qShape[qShape.length - 1] = L; rShape[rShape.length - 2] = L;
  const Q = alloc(batch * N * L), R = alloc(batch * L * M);
  addon.qr(a, Q, R, batch, N, M);
  return [new NDArray(qShape, Q), new NDArray(rShape, R)];
}

// _qr_decomp_inplace(M,N,L, A,A_off, Y,Y_off) (qr.js:147-183): same in-place contract on flat arrays; A becomes R, Y becomes Q^T Y.
function _qr_decomp_inplace(M, N, L, A, A_off, Y, Y_off) {
  for (const x of [M, N, L, A_off, Y_off]) if (x % 1 !== 0 || !(0 <= x)) throw new Error('Assertion failed.');
  if (!(M * N <= A.length - A_off) || !(M * L <= Y.length - Y_off)) throw new Error('Assertion failed.');
  if (M === 0 || N === 0 || L === 0) return;
  const a = Float64Array.from(A.subarray(A_off, A_off + M * N)), y = Float64Array.from(Y.subarray(Y_off, Y_off + M * L));
  const R = new Float64Array(M * N), QtY = new Float64Array(M * L);
  addon.qrInplace(a, y, R, QtY, 1, M, N, L);
  A.set(R, A_off); Y.set(QtY, Y_off);
}

function svd_jac_1sided(A) {
  A = asarray(A);
  if (A.dtype.startsWith('complex')) throw new Error('svd_jac_1sided(A): A.dtype must be float.');
  if (A.ndim < 2) throw new Error('svd_jac_1sided(A): A.ndim must be at least 2.');
  const [N, M] = A.shape.slice(-2), L = Math.min(N, M), a = f64(A, 'svd_jac_1sided'), batch = a.length / (N * M);
  const uShape = Int32Array.from(A.shape), vShape = Int32Array.from(A.shape), sShape = A.shape.slice(0, -1);
  uShape[uShape.length - 1] = L; vShape[vShape.length - 2] = L; sShape[sShape.length - 1] = L;
  const U = alloc(batch * N * L), sv = alloc(batch * L), V = alloc(batch * L * M);
  addon.svdJac1(a, U, sv, V, batch, N, M);
  return [new NDArray(uShape, U), new NDArray(sShape, sv), new NDArray(vShape, V)];
}

// tril_solve / triu_solve (tri.js:156-293) and cholesky_solve (cholesky.js:75-144): op 0 / 1 / 2
function triSolve(op, T, Y, errT, errY) {
  T = asarray(T); if (T.ndim < 2) throw new Error(errT);
  Y = asarray(Y); if (Y.ndim < 2) throw new Error(errY);
  const ndim = Math.max(T.ndim, Y.ndim), xShape = new Int32Array(ndim).fill(1);
  xShape[ndim - 2] = Y.shape[Y.ndim - 2]; xShape[ndim - 1] = Y.shape[Y.ndim - 1];
  for (const shp of [T.shape, Y.shape])
    for (let i = ndim - 2, j = shp.length - 2; i-- > 0 && j-- > 0;)
      if (xShape[i] === 1) xShape[i] = shp[j];   // a mismatch is reported by the library with the reference's text
  const X = alloc(xShape.reduce((a, b) => a * b, 1));
  addon.triSolve(op, f64(T, 'tri_solve'), T.shape, f64(Y, 'tri_solve'), Y.shape, X, xShape);
  return new NDArray(xShape, X);
}
const tril_solve = (L, Y) => triSolve(0, L, Y, 'tril_solve(L,Y): L.ndim must be at least 2.', 'tril_solve(L,Y): Y.ndim must be at least 2.');
const triu_solve = (U, Y) => triSolve(1, U, Y, 'triu_solve(U,Y): U.ndim must be at least 2.', 'triu_solve(U,Y): Y.ndim must be at least 2.');
const cholesky_solve = (L, y) => triSolve(2, L, y, 'L must be at least 2D.', 'y must be at least 2D.');

// qr_lstsq(Q,R,y) (qr.js:186-273): fused bit-exact kernel for thin factors with one common batch, otherwise Q^T y on the GEMM
// path followed by the bit-exact back substitution
function qr_lstsq(Q, R, y) {
  if (y === undefined) { y = R; [Q, R] = Q; }
  Q = asarray(Q); if (Q.ndim < 2) throw new Error('qr_lstsq(Q,R,y): Q.ndim must be at least 2.');
  R = asarray(R); if (R.ndim < 2) throw new Error('qr_lstsq(Q,R,y): R.ndim must be at least 2.');
  y = asarray(y); if (y.ndim < 2) throw new Error('qr_lstsq(Q,R,y): y.ndim must be at least 2.');
  const [N, M] = Q.shape.slice(-2), I = R.shape[R.ndim - 1], J = y.shape[y.ndim - 1], L = Math.min(M, I);
  if (N !== y.shape[y.ndim - 2]) throw new Error("qr_lstsq(Q,R,y): Q and y don't match.");
  if (M !== R.shape[R.ndim - 2]) throw new Error("qr_lstsq(Q,R,y): Q and R don't match.");
  if (I > N) throw new Error('qr_lstsq(Q,R,y): Under-determined systems not supported. Use rrqr instead.');
  const lead = a => Array.from(a.shape.slice(0, -2)).join(), sameBatch = lead(Q) === lead(R) && lead(Q) === lead(y);
  if (M <= 32 && I <= 32 && sameBatch) {
    const xShape = Int32Array.from([...Q.shape.slice(0, -2), I, J]), q = f64(Q, 'qr_lstsq'), batch = q.length / (N * M);
    const X = alloc(batch * I * J);
    addon.qrLstsq(q, f64(R, 'qr_lstsq'), f64(y, 'qr_lstsq'), X, batch, N, M, I, J);
    return new NDArray(xShape, X);
  }
  // composed path: Q^T y by matmul2 (broadcasting as in matmul2), its first L rows solved against R[:L,:L]; rows L..I-1 stay zero
  const QTy = matmul2(Q.T, y);                                    // [..., M, J]
  const nq = QTy.data.length / (M * J), top = new Float64Array(nq * L * J);
  for (let b = 0; b < nq; b++) top.set(QTy.data.subarray(b * M * J, b * M * J + L * J), b * L * J);
  const topShape = Int32Array.from(QTy.shape); topShape[topShape.length - 2] = L;
  const rd = f64(R, 'qr_lstsq'), nr = rd.length / (M * I), Rl = new Float64Array(nr * L * L);
  for (let b = 0; b < nr; b++)
    for (let i = 0; i < L; i++) Rl.set(rd.subarray(b * M * I + i * I, b * M * I + i * I + L), b * L * L + i * L);
  const rlShape = Int32Array.from(R.shape); rlShape[rlShape.length - 2] = L; rlShape[rlShape.length - 1] = L;
  const x = triu_solve(new NDArray(rlShape, Rl), new NDArray(topShape, top));   // [..., L, J]
  if (L === I) return x;
  const xShape = Int32Array.from(x.shape); xShape[xShape.length - 2] = I;
  const X = new Float64Array(xShape.reduce((p, q) => p * q, 1));
  for (let b = 0; b * I * J < X.length; b++) X.set(x.data.subarray(b * L * J, (b + 1) * L * J), b * I * J);
  return new NDArray(xShape, X);
}

// svd_rank / svd_lstsq / svd_solve (svd.js:31-226): the rank rule and the fused solve run on the device
function svd_rank(sv) {
  sv = asarray(sv);
  const N = sv.shape[sv.ndim - 1], rShape = sv.shape.slice(0, -1), d = f64(sv, 'svd_rank'), r = new Int32Array(d.length / N);
  addon.svdRank(d, r, N);   // throws 'svd_rank(): NaN or Infinity encountered.'
  return new NDArray(rShape, r);
}

function svd_lstsq(U, sv, V, y) {
  if (y == undefined) {
    if (V != undefined) throw new Error('svd_lstsq(Q,R,P, y): Either 2 ([Q,R,P], y) or 4 arguments (Q,R,P, y) expected.');
    y = sv; [U, sv, V] = U;
  }
  U = asarray(U); sv = asarray(sv); V = asarray(V); y = asarray(y);
  const xShapeMax = new Int32Array(Math.max(U.ndim, sv.ndim + 1, V.ndim, y.ndim, 2));
  const ndim = addon.svdLstsqShape(U.shape, sv.shape, V.shape, y.shape, xShapeMax);   // throws the reference's texts
  const xShape = xShapeMax.slice(0, ndim), X = alloc(xShape.reduce((a, b) => a * b, 1));
  addon.svdLstsq(f64(U, 'svd_lstsq'), U.shape, f64(sv, 'svd_lstsq'), sv.shape, f64(V, 'svd_lstsq'), V.shape, f64(y, 'svd_lstsq'), y.shape, X, xShape);
  return new NDArray(xShape, X);
}

// The reference's singularity scan (svd.js:87-95) starts from an undefined loop variable and never runs: svd_solve returns
// the least-squares solution for every square system.
function svd_solve(U, sv, V, y) {
  if (y == undefined) {
    if (V != undefined) throw new Error('svd_lstsq(Q,R,P, y): Either 2 ([Q,R,P], y) or 4 arguments (Q,R,P, y) expected.');
    y = sv; [U, sv, V] = U;
  }
  U = asarray(U); sv = asarray(sv); V = asarray(V);
  if (U.shape[U.ndim - 2] !== V.shape[V.ndim - 1]) throw new Error('rrqr_solve(Q,R,P, y): System not square.');
  return svd_lstsq(U, sv, V, y);
}

module.exports = {svd_rank, svd_lstsq, svd_solve, qr_lstsq, matmul2, matmul, cholesky_decomp, qr_decomp, _qr_decomp_inplace, svd_jac_1sided, tril_solve, triu_solve, cholesky_solve,
                  init: d => addon.init(Int32Array.from(d || [])), stats: addon.stats,
                  pinnedFloat64Array: addon.pinnedFloat64Array};
