/* nd4ref.c — CPU ORACLE (test infrastructure only; see nd4ref.h for the rules and pinning status).
 *
 * Plain C restatement of the nd4js v1.3.0 routines on the batched dense-LA hot path.
 * Build:  gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared   (see oracle/Makefile)
 * JS numbers are IEEE-754 doubles and V8 never fuses a*b+c, hence -ffp-contract=off.
 */
#include "nd4ref.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define EPS64 2.220446049250313e-16 /* Number.EPSILON, src/dt/index.js:33-43 */

/* JS Math.max propagates NaN (C fmax does not). */
static double js_max(double a, double b) {
  if (isnan(a) || isnan(b)) return NAN;
  return a > b ? a : b;
}

/* ------------------------------------------------------------------ matmul ---- */

/* src/la/matmul.js:91-116 */
int nd4ref_matmul_shape(const int32_t* a_shape, int a_ndim, const int32_t* b_shape, int b_ndim,
                        int32_t* c_shape, int* c_ndim) {
  if (a_ndim < 2) return ND4REF_E_A_NDIM;
  if (b_ndim < 2) return ND4REF_E_B_NDIM;
  const int32_t I = a_shape[a_ndim - 2], K = a_shape[a_ndim - 1], J = b_shape[b_ndim - 1];
  if (b_shape[b_ndim - 2] != K) return ND4REF_E_INNER;
  const int ndim = a_ndim > b_ndim ? a_ndim : b_ndim;
  for (int d = 0; d < ndim; d++) c_shape[d] = 1;
  c_shape[ndim - 2] = I;
  c_shape[ndim - 1] = J;
  const int32_t* shp[2] = {a_shape, b_shape};
  const int nd[2] = {a_ndim, b_ndim};
  for (int w = 0; w < 2; w++)
    for (int i = ndim - 2, j = nd[w] - 2; i-- > 0 && j-- > 0;) {
      if (c_shape[i] == 1) c_shape[i] = shp[w][j];
      else if (c_shape[i] != shp[w][j] && shp[w][j] != 1) return ND4REF_E_BROADCAST;
    }
  *c_ndim = ndim;
  return ND4REF_OK;
}

/* src/la/matmul.js:31-74: the generated odometer walks C's leading dims in row-major
 * order and rewinds an operand whose dim is 1/absent; here that is expressed as a
 * stride that is 0 for such dims.  Inner loops keep the reference's i-k-j order. */
int nd4ref_matmul_f64(const double* A, const int32_t* a_shape, int a_ndim,
                      const double* B, const int32_t* b_shape, int b_ndim,
                      double* C, const int32_t* c_shape, int c_ndim) {
  int32_t want[64];
  int want_ndim = 0;
  if (a_ndim > 64 || b_ndim > 64) return ND4REF_E_SHAPE;
  int rc = nd4ref_matmul_shape(a_shape, a_ndim, b_shape, b_ndim, want, &want_ndim);
  if (rc) return rc;
  if (want_ndim != c_ndim) return ND4REF_E_SHAPE;
  for (int d = 0; d < c_ndim; d++)
    if (want[d] != c_shape[d]) return ND4REF_E_SHAPE;

  const int64_t I = a_shape[a_ndim - 2], K = a_shape[a_ndim - 1], J = b_shape[b_ndim - 1];
  const int nb = c_ndim - 2;
  int64_t a_str[64], b_str[64], idx[64];
  {
    int64_t sa = I * K, sb = K * J;
    for (int d = nb - 1; d >= 0; d--) {
      const int da = d - c_ndim + a_ndim, db = d - c_ndim + b_ndim;
      const int64_t na = da >= 0 ? a_shape[da] : 1, nbb = db >= 0 ? b_shape[db] : 1;
      a_str[d] = na > 1 ? sa : 0;
      b_str[d] = nbb > 1 ? sb : 0;
      sa *= na;
      sb *= nbb;
      idx[d] = 0;
    }
  }
  int64_t total = 1;
  for (int d = 0; d < nb; d++) total *= c_shape[d];

  for (int64_t m = 0; m < total; m++) {
    int64_t ao = 0, bo = 0;
    for (int d = 0; d < nb; d++) { ao += idx[d] * a_str[d]; bo += idx[d] * b_str[d]; }
    const double* a = A + ao;
    const double* b = B + bo;
    double* c = C + m * I * J;
    for (int64_t x = 0; x < I * J; x++) c[x] = 0.0;
    for (int64_t i = 0; i < I; i++)
      for (int64_t k = 0; k < K; k++) {
        const double aik = a[i * K + k];
        for (int64_t j = 0; j < J; j++) c[i * J + j] += aik * b[k * J + j];
      }
    for (int d = nb - 1; d >= 0; d--) {
      if (++idx[d] < c_shape[d]) break;
      idx[d] = 0;
    }
  }
  return ND4REF_OK;
}

/* ---------------------------------------------------------------- cholesky ---- */

/* src/la/cholesky.js:27-47 with KahanSum (src/kahan_sum.js:19-42) inlined. */
static int ref_cholesky_one(int n, double* L) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      double sum = L[n * i + j], rst = 0.0; /* kahan.set */
      if (isnan(sum)) return ND4REF_E_NAN_INPUT;
      for (int k = 0; k < j; k++) { /* kahan.add( -L_ik*L_jk ) */
        const double val = -L[n * i + k] * L[n * j + k];
        const double cor = val - rst;
        const double s2 = sum + cor;
        rst = (s2 - sum) - cor;
        sum = s2;
      }
      if (i > j) L[n * i + j] = sum / L[n * j + j];
      else {
        L[n * i + i] = sqrt(sum);
        if (isnan(L[n * i + i])) return ND4REF_E_SINGULAR;
      }
    }
  return ND4REF_OK;
}

/* src/la/cholesky.js:50-72 */
int nd4ref_cholesky_f64(const double* S, double* L, int64_t batch, int n, int64_t* first_bad) {
  if (first_bad) *first_bad = -1;
  const int64_t nn = (int64_t)n * n;
  memset(L, 0, sizeof(double) * (size_t)(batch * nn));
  for (int64_t b = 0; b < batch; b++) {
    double* l = L + b * nn;
    const double* s = S + b * nn;
    for (int i = 0; i < n; i++)
      for (int j = 0; j <= i; j++) l[n * i + j] = s[n * i + j];
    int rc = ref_cholesky_one(n, l);
    if (rc) {
      if (first_bad) *first_bad = b;
      return rc;
    }
  }
  return ND4REF_OK;
}

/* ---------------------------------------------------------------- givens ------ */

/* src/la/_giv_rot.js:22-37 */
void nd4ref_giv_rot_qr(double a, double b, double out[3]) {
  const double mx = js_max(fabs(a), fabs(b));
  if (mx == 0.0) { out[0] = 1; out[1] = 0; out[2] = 0; return; }
  a /= mx;
  b /= mx;
  double norm = sqrt(a * a + b * b);
  a /= norm;
  b /= norm;
  norm *= mx;
  out[0] = a; out[1] = b; out[2] = norm;
}

/* src/la/_giv_rot.js:42-67 */
static void giv_rot_rows(double* W, int64_t N, int64_t i, int64_t j, double c, double s) {
  for (int64_t k = N; k-- > 0;) {
    const double wi = W[i], wj = W[j];
    W[i] = c * wi + s * wj;
    W[j] = c * wj - s * wi;
    i++; j++;
  }
}

/* src/la/_giv_rot.js:72-87 */
static void giv_rot_cols(double* W, int64_t N, int64_t i, int64_t j, double c, double s) {
  for (int64_t k = N; k-- > 0;) {
    const double wi = W[i], wj = W[j];
    W[i] = c * wi - s * wj;
    W[j] = c * wj + s * wi;
    i += N; j += N;
  }
}

/* src/la/transpose_inplace.js:22-31 */
static void transpose_inplace(int64_t N, double* A) {
  for (int64_t i = N; --i > 0;)
    for (int64_t j = i; j-- > 0;) {
      const double t = A[N * i + j];
      A[N * i + j] = A[N * j + i];
      A[N * j + i] = t;
    }
}

/* -------------------------------------------------------------------- qr ------ */

/* src/la/qr.js:27-77; here rows=M, cols=N in the reference's naming. Q is MxM, R is MxN. */
static int ref_qr_full_one(int M, int N, const double* A, double* Q, double* R) {
  const int B = 8;
  memcpy(R, A, sizeof(double) * (size_t)M * N);
  memset(Q, 0, sizeof(double) * (size_t)M * M);
  for (int i = 0; i < M; i++) Q[M * i + i] = 1.0;
  for (int J = 0; J < N; J += B)
    for (int I = J; I < M; I += B)
      for (int i = I; i < I + B && i < M; i++)
        for (int j = J; j < J + B && j < N && j < i; j++) {
          const int64_t ij = (int64_t)N * i + j;
          const double R_ij = R[ij];
          if (R_ij == 0.0) continue;
          const int64_t jj = (int64_t)N * j + j;
          double g[3];
          nd4ref_giv_rot_qr(R[jj], R_ij, g);
          if (!(0 <= g[2])) return ND4REF_E_NAN_INPUT; /* DEBUG assertion _giv_rot.js:35 */
          R[ij] = 0.0;
          if (g[1] == 0.0) continue;
          R[jj] = g[2];
          giv_rot_rows(R, N - 1 - j, jj + 1, ij + 1, g[0], g[1]);
          giv_rot_rows(Q, 1 + i, (int64_t)M * j, (int64_t)M * i, g[0], g[1]);
        }
  transpose_inplace(M, Q);
  return ND4REF_OK;
}

/* src/la/qr.js:93-139; rows=N > cols=M in the reference's naming. Q is NxM, R is MxM. */
static int ref_qr_tall_one(int N, int M, const double* A, double* Q, double* R) {
  memcpy(Q, A, sizeof(double) * (size_t)N * M);
  memset(R, 0, sizeof(double) * (size_t)M * M);
  for (int i = 1; i < N; i++) {
    const int I = i < M ? i : M;
    for (int j = 0; j < I; j++) {
      const int64_t ij = (int64_t)M * i + j;
      const double R_ij = Q[ij];
      if (R_ij == 0.0) continue;
      const int64_t jj = (int64_t)M * j + j;
      double g[3];
      nd4ref_giv_rot_qr(Q[jj], R_ij, g);
      if (!(0 <= g[2])) return ND4REF_E_NAN_INPUT;
      double c = g[0], s = g[1], norm = g[2];
      if (s != 0.0) {
        if (c < 0) { c *= -1; s *= -1; norm *= -1; }
        giv_rot_rows(Q, M - 1 - j, jj + 1, ij + 1, c, s);
        Q[jj] = norm;
      }
      Q[ij] = s;
    }
  }
  for (int i = 0; i < M; i++)
    for (int j = i; j < M; j++) {
      R[M * i + j] = Q[M * i + j];
      Q[M * i + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int i = N; --i > 0;) {
    const int I = i < M ? i : M;
    for (int j = I; j-- > 0;) {
      const int64_t ij = (int64_t)M * i + j;
      const double s = Q[ij];
      if (s == 0.0) continue;
      Q[ij] = 0.0;
      const double c = sqrt((1 - s) * (1 + s));
      giv_rot_rows(Q, M - j, ij, (int64_t)M * j + j, c, s);
    }
  }
  return ND4REF_OK;
}

/* src/la/qr.js:80-145 */
/* _qr_decomp_inplace(M,N,L, A,A_off, Y,Y_off): src/la/qr.js:147-183.  Givens elimination of A_ij (i > j) in the
 * reference's order, the same rotations applied to the rows of Y; A ends as R (zeros below the diagonal), Y as Q^T Y. */
static void ref_qr_inplace_one(int M, int N, int L, double* A, double* Y) {
  for (int i = 1; i < M; i++)
    for (int j = 0; j < N && j < i; j++) {
      const int64_t ij = (int64_t)N * i + j, jj = (int64_t)N * j + j;
      const double A_ij = A[ij];
      if (0 == A_ij) continue;
      double csn[3];
      nd4ref_giv_rot_qr(A[jj], A_ij, csn);
      A[ij] = 0;
      if (0 == csn[1]) continue;
      A[jj] = csn[2];
      giv_rot_rows(A, N - 1 - j, jj + 1, ij + 1, csn[0], csn[1]);
      giv_rot_rows(Y, L, (int64_t)L * j, (int64_t)L * i, csn[0], csn[1]);
    }
}

int nd4ref_qr_inplace_f64(const double* A, const double* Y, double* R, double* QtY, int64_t batch, int M, int N, int L) {
  if (!A || !Y || !R || !QtY || batch < 1 || M < 1 || N < 1 || L < 1) return ND4REF_E_SHAPE;
  memcpy(R, A, sizeof(double) * (size_t)batch * M * N);
  memcpy(QtY, Y, sizeof(double) * (size_t)batch * M * L);
  for (int64_t b = 0; b < batch; b++) ref_qr_inplace_one(M, N, L, R + b * (int64_t)M * N, QtY + b * (int64_t)M * L);
  return 0;
}

int nd4ref_qr_f64(const double* A, double* Q, double* R, int64_t batch, int rows, int cols) {
  const int L = rows < cols ? rows : cols;
  for (int64_t b = 0; b < batch; b++) {
    const double* a = A + b * (int64_t)rows * cols;
    double* q = Q + b * (int64_t)rows * L;
    double* r = R + b * (int64_t)L * cols;
    int rc = rows <= cols ? ref_qr_full_one(rows, cols, a, q, r) : ref_qr_tall_one(rows, cols, a, q, r);
    if (rc) return rc;
  }
  return ND4REF_OK;
}

/* qr_decomp_full for any shape (src/la/qr.js:27-77): Q [batch,rows,rows], R [batch,rows,cols]. */
int nd4ref_qr_full_f64(const double* A, double* Q, double* R, int64_t batch, int rows, int cols) {
  if (!A || !Q || !R || batch < 1 || rows < 1 || cols < 1) return ND4REF_E_SHAPE;
  for (int64_t b = 0; b < batch; b++) {
    int rc = ref_qr_full_one(rows, cols, A + b * (int64_t)rows * cols, Q + b * (int64_t)rows * rows, R + b * (int64_t)rows * cols);
    if (rc) return rc;
  }
  return ND4REF_OK;
}

/* -------------------------------------------------------------------- svd ----- */

/* src/la/_svd_jac_utils.js:72-114 */
void nd4ref_svd_jac_angles(double S_pp, double S_pq, double S_qp, double S_qq, double out[4]) {
  double x = atan2(S_qp - S_pq, S_qq + S_pp), y = atan2(S_qp + S_pq, S_qq - S_pp);
  const double a = (x - y) / 2, b = (x + y) / 2;
  double ca = cos(a), sa = sin(a), cb = cos(b), sb = sin(b);
  x = cb * (sa * S_qp + ca * S_pp) - sb * (sa * S_qq + ca * S_pq);
  y = sb * (ca * S_qp - sa * S_pp) + cb * (ca * S_qq - sa * S_pq);
  if (fabs(x) < fabs(y)) {
    double t = sa; sa = ca; ca = -t; /* [sa,ca] = [ca,-sa] */
    t = cb; cb = sb; sb = -t;        /* [cb,sb] = [sb,-cb] */
    x = y;
  }
  if (x < 0) { cb = -cb; sb = -sb; }
  out[0] = ca; out[1] = sa; out[2] = cb; out[3] = sb;
}

static const double* g_sort_key;
static int cmp_desc_stable(const void* pa, const void* pb) {
  const int i = *(const int*)pa, j = *(const int*)pb;
  const double d = g_sort_key[j] - g_sort_key[i]; /* comparator (i,j) => sv[j]-sv[i] */
  if (d < 0) return -1;
  if (d > 0) return 1;
  return i < j ? -1 : (i > j ? 1 : 0); /* Array.prototype.sort is stable; ord starts sorted */
}

/* src/la/_svd_jac_utils.js:139-188 (UT holds U transposed on entry, U on exit). */
static void svd_jac_post_skip1(int N, double* UT, double* V, double* sv, int* ord) {
  for (int i = N; i-- > 0;) {
    const double s = sv[i];
    if (s < 0 || (s == 0 && signbit(s))) {
      sv[i] = -s;
      for (int j = 0; j < N; j++) UT[N * i + j] *= -1;
    }
  }
  /* ord is the identity here (the swap cycles below restore it), so tie-break on index = stable */
  g_sort_key = sv;
  qsort(ord, (size_t)N, sizeof(int), cmp_desc_stable);
  for (int i = 0; i < N; ++i)
    for (int j = i;;) {
      int tmp = ord[j];
      ord[j] = j;
      j = tmp;
      if (j <= i) break;
      double* rI = UT + (int64_t)ord[j] * N; /* NB: ord[j] is read AFTER the assignment above, as in the JS */
      double* rJ = UT + (int64_t)j * N;
      for (int k = 0; k < N; k++) { const double t = rI[k]; rI[k] = rJ[k]; rJ[k] = t; }
      rI = V + (int64_t)ord[j] * N;
      rJ = V + (int64_t)j * N;
      for (int k = 0; k < N; k++) { const double t = rI[k]; rI[k] = rJ[k]; rJ[k] = t; }
      const double t = sv[ord[j]];
      sv[ord[j]] = sv[j];
      sv[j] = t;
    }
  for (int i = 0; i < N - 1; i++)
    for (int j = i; ++j < N;) {
      const double t = UT[N * i + j];
      UT[N * i + j] = UT[N * j + i];
      UT[N * j + i] = t;
    }
}

/* src/la/svd_jac_2sided.js:54-143, square matrices only. */
static int ref_svd_jac2_square(const double* A, double* U, double* sv, double* V,
                               int64_t batch, int N, int* sweeps_out) {
  const double TOL = (N * EPS64) * (N * EPS64);
  const int B = 8;
  int max_sweeps = 0;
  if (N == 1) { /* :66-78 — returns [sign, |a|, 1] */
    for (int64_t i = 0; i < batch; i++) {
      double a = A[i];
      if (a < 0.0) { sv[i] = -a; U[i] = -1.0; } else { sv[i] = a; U[i] = 1.0; }
      V[i] = 1.0;
    }
    if (sweeps_out) *sweeps_out = 0;
    return ND4REF_OK;
  }
  double* S = (double*)malloc(sizeof(double) * (size_t)N * N);
  int* ord = (int*)malloc(sizeof(int) * (size_t)N);
  for (int i = 0; i < N; i++) ord[i] = i;
  for (int64_t m = 0; m < batch; m++) {
    double* u = U + m * (int64_t)N * N;
    double* v = V + m * (int64_t)N * N;
    const double* a = A + m * (int64_t)N * N;
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) {
        S[N * i + j] = a[N * i + j];
        u[N * i + j] = (i == j);
        v[N * i + j] = (i == j);
      }
    int sweeps = 0;
    for (int finished = 0; !finished;) {
      finished = 1;
      sweeps++;
      for (int Q = 0; Q < N; Q += B)
        for (int P = 0; P <= Q; P += B)
          for (int q = Q; q < Q + B && q < N; q++)
            for (int p = P; p < P + B && p < q; p++) {
              const double S_pp = S[N * p + p], S_pq = S[N * p + q], S_qp = S[N * q + p], S_qq = S[N * q + q];
              if (!(S_pq * S_pq + S_qp * S_qp > fabs(S_pp * S_qq) * TOL)) continue;
              finished = 0;
              double g[4];
              nd4ref_svd_jac_angles(S_pp, S_pq, S_qp, S_qq, g);
              giv_rot_rows(S, N, (int64_t)N * p, (int64_t)N * q, g[0], g[1]);
              giv_rot_cols(S, N, p, q, g[2], g[3]);
              S[N * p + q] = S[N * q + p] = 0.0;
              giv_rot_rows(u, N, (int64_t)N * p, (int64_t)N * q, g[0], g[1]);
              giv_rot_rows(v, N, (int64_t)N * p, (int64_t)N * q, g[2], -g[3]);
            }
      if (sweeps > 10000) { free(S); free(ord); return ND4REF_E_NAN_INPUT; } /* NaN input never converges */
    }
    if (sweeps > max_sweeps) max_sweeps = sweeps;
    double* s = sv + m * N;
    for (int i = 0; i < N; i++) s[i] = S[N * i + i]; /* _svd_jac_post :123-134 */
    svd_jac_post_skip1(N, u, v, s, ord);
  }
  free(S);
  free(ord);
  if (sweeps_out) *sweeps_out = max_sweeps;
  return ND4REF_OK;
}

/* plain [rows,inner]x[inner,cols] per matrix, reference loop order */
static void mm(const double* a, const double* b, double* c, int I, int K, int J) {
  for (int x = 0; x < I * J; x++) c[x] = 0.0;
  for (int i = 0; i < I; i++)
    for (int k = 0; k < K; k++)
      for (int j = 0; j < J; j++) c[i * J + j] += a[i * K + k] * b[k * J + j];
}

/* src/la/svd_jac_2sided.js:30-53 dispatch on shape. */
int nd4ref_svd_jac2_f64(const double* A, double* U, double* sv, double* V,
                        int64_t batch, int rows, int cols, int* sweeps_out) {
  if (rows == cols) return ref_svd_jac2_square(A, U, sv, V, batch, rows, sweeps_out);
  const int L = rows < cols ? rows : cols, T = rows > cols ? rows : cols;
  /* tall: A = Q R, R = U' S V  ->  U = Q U'.   wide: A^T = Q R, R = U' S V' -> A = V'^T S (Q U')^T */
  double* At = (double*)malloc(sizeof(double) * (size_t)T * L);
  double* Q = (double*)malloc(sizeof(double) * (size_t)T * L);
  double* R = (double*)malloc(sizeof(double) * (size_t)L * L);
  double* u2 = (double*)malloc(sizeof(double) * (size_t)L * L);
  double* v2 = (double*)malloc(sizeof(double) * (size_t)L * L);
  double* qu = (double*)malloc(sizeof(double) * (size_t)T * L);
  int rc = ND4REF_OK, max_sweeps = 0;
  for (int64_t m = 0; m < batch && !rc; m++) {
    const double* a = A + m * (int64_t)rows * cols;
    double* u = U + m * (int64_t)rows * L;
    double* s = sv + m * (int64_t)L;
    double* v = V + m * (int64_t)L * cols;
    int sw = 0;
    if (rows > cols) {
      rc = nd4ref_qr_f64(a, Q, R, 1, rows, cols);
      if (!rc) rc = ref_svd_jac2_square(R, u2, s, v, 1, L, &sw);
      if (!rc) mm(Q, u2, u, rows, L, L);
    } else {
      for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++) At[j * rows + i] = a[i * cols + j];
      rc = nd4ref_qr_f64(At, Q, R, 1, cols, rows);
      if (!rc) rc = ref_svd_jac2_square(R, u2, s, v2, 1, L, &sw);
      if (!rc) {
        transpose_inplace(L, v2);
        memcpy(u, v2, sizeof(double) * (size_t)L * L);
        mm(Q, u2, qu, cols, L, L);
        for (int i = 0; i < cols; i++)
          for (int j = 0; j < L; j++) v[j * cols + i] = qu[i * L + j];
      }
    }
    if (sw > max_sweeps) max_sweeps = sw;
  }
  free(At); free(Q); free(R); free(u2); free(v2); free(qu);
  if (sweeps_out) *sweeps_out = max_sweeps;
  return rc;
}

/* ---------------------------------------------------------------- solves ----- */

/* src/la/tri.js:45-71 (_tril_solve), :73-98 (_triu_solve), :100-125 (_tril_t_solve); M rows, N = row stride of T, O rhs columns */
static void tril_solve1(int M, int N, int O, const double* L, double* X) {
  for (int i = 0; i < M; i++) {
    for (int k = 0; k < i; k++)
      for (int j = 0; j < O; j++) X[O * i + j] -= L[N * i + k] * X[O * k + j];
    for (int j = 0; j < O; j++) X[O * i + j] /= L[N * i + i];
  }
}
static void triu_solve1(int M, int N, int O, const double* U, double* X) {
  for (int i = M; i-- > 0;)
    for (int j = O; j-- > 0;) {
      for (int k = M; --k > i;) X[O * i + j] -= U[N * i + k] * X[O * k + j];
      X[O * i + j] /= U[N * i + i];
    }
}
static void tril_t_solve1(int M, int N, int O, const double* L, double* X) {
  for (int k = M; k-- > 0;) {
    for (int j = O; j-- > 0;) X[O * k + j] /= L[N * k + k];
    for (int i = k; i-- > 0;)
      for (int j = O; j-- > 0;) X[O * i + j] -= L[N * k + i] * X[O * k + j];
  }
}

/* op 0: tril_solve (tri.js:156-230), 1: triu_solve (:233-293), 2: cholesky_solve (cholesky.js:75-144).
 * T[...,M,M], Y[...,M,J] -> X[...,M,J], leading dims broadcast as in the reference's solv() recursion. */
int nd4ref_tri_solve_f64(int op, const double* T, const int32_t* t_shape, int t_ndim,
                         const double* Y, const int32_t* y_shape, int y_ndim,
                         double* X, const int32_t* x_shape, int x_ndim) {
  if (t_ndim < 2 || y_ndim < 2 || t_ndim > 64 || y_ndim > 64) return ND4REF_E_A_NDIM;
  const int M = y_shape[y_ndim - 2], J = y_shape[y_ndim - 1];
  if (t_shape[t_ndim - 2] != M) return ND4REF_E_INNER;
  if (t_shape[t_ndim - 1] != M) return ND4REF_E_NOT_SQUARE;
  const int ndim = t_ndim > y_ndim ? t_ndim : y_ndim;
  if (x_ndim != ndim) return ND4REF_E_SHAPE;
  int32_t want[64];
  for (int d = 0; d < ndim; d++) want[d] = 1;
  want[ndim - 2] = M;
  want[ndim - 1] = J;
  const int32_t* shp[2] = {t_shape, y_shape};
  const int nd[2] = {t_ndim, y_ndim};
  for (int w = 0; w < 2; w++)
    for (int i = ndim - 2, j = nd[w] - 2; i-- > 0 && j-- > 0;) {
      if (want[i] == 1) want[i] = shp[w][j];
      else if (want[i] != shp[w][j] && shp[w][j] != 1) return ND4REF_E_BROADCAST;
    }
  for (int d = 0; d < ndim; d++)
    if (want[d] != x_shape[d]) return ND4REF_E_SHAPE;
  const int nb = ndim - 2;
  int64_t t_str[64], y_str[64], idx[64], total = 1;
  {
    int64_t st = (int64_t)M * M, sy = (int64_t)M * J;
    for (int d = nb - 1; d >= 0; d--) {
      const int dt = d - ndim + t_ndim, dy = d - ndim + y_ndim;
      const int64_t nt = dt >= 0 ? t_shape[dt] : 1, ny = dy >= 0 ? y_shape[dy] : 1;
      t_str[d] = nt > 1 ? st : 0;
      y_str[d] = ny > 1 ? sy : 0;
      st *= nt;
      sy *= ny;
      idx[d] = 0;
      total *= x_shape[d];
    }
  }
  for (int64_t m = 0; m < total; m++) {
    int64_t to = 0, yo = 0;
    for (int d = 0; d < nb; d++) { to += idx[d] * t_str[d]; yo += idx[d] * y_str[d]; }
    double* x = X + m * (int64_t)M * J;
    memcpy(x, Y + yo, sizeof(double) * (size_t)M * J);
    if (op == 0) tril_solve1(M, M, J, T + to, x);
    else if (op == 1) triu_solve1(M, M, J, T + to, x);
    else { tril_solve1(M, M, J, T + to, x); tril_t_solve1(M, M, J, T + to, x); }
    for (int d = nb - 1; d >= 0; d--) {
      if (++idx[d] < x_shape[d]) break;
      idx[d] = 0;
    }
  }
  return ND4REF_OK;
}

/* ------------------------------------------------------------------- norm ----- */

/* src/la/norm.js:22-68 (FrobeniusNorm.include / result) */
double nd4ref_frobenius(const double* x, int64_t n) {
  double sum = 0.0, max = 0.0;
  for (int64_t i = 0; i < n; i++) {
    double v = fabs(x[i]);
    if (v != 0) {
      if (max < v) {
        const double s = max / v;
        sum *= s * s;
        max = v;
      }
      v /= max;
      sum += v * v;
    }
  }
  return isfinite(max) ? sqrt(sum) * max : max;
}

/* ------------------------------------------------------- svd_rank / svd_lstsq ----- */

#define ND4REF_SQRT_EPS 1.4901161193847656e-08 /* Math.sqrt(Number.EPSILON) = 2^-26 */

/* src/la/svd.js:31-58 — rank[b] = index of the first |sv_r| <= sqrt(eps)*|sv_0|; a non-finite entry met before that throws */
int nd4ref_svd_rank_f64(const double* sv, int32_t* rank, int64_t batch, int n) {
  for (int64_t off = 0; off < batch; off++) {
    const double T = ND4REF_SQRT_EPS * fabs(sv[n * off]);
    int r = 0;
    for (; r < n; r++) {
      const double sv_r = fabs(sv[n * off + r]);
      if (!isfinite(sv_r)) return ND4REF_E_NAN_INPUT;
      if (sv_r <= T) break;
    }
    rank[off] = r;
  }
  return ND4REF_OK;
}

/* src/la/svd.js:103-226 — U[...,N,M], sv[...,M], V[...,M,I], y[...,N,J] -> x[...,I,J] (shape checked by the caller's wrapper
 * against nd4ref_svd_lstsq_shape).  The recursion `solv(d)` with its rewinding offsets (:154-222) is restated as an
 * odometer over the result's leading dims with stride 0 for broadcast dims, which visits the same operand slices. */
int nd4ref_svd_lstsq_shape(const int32_t* u_shape, int u_ndim, const int32_t* sv_shape, int sv_ndim,
                           const int32_t* v_shape, int v_ndim, const int32_t* y_shape, int y_ndim,
                           int32_t* x_shape, int* x_ndim) {
  if (u_ndim < 2) return ND4REF_E_A_NDIM;
  if (sv_ndim < 1) return ND4REF_E_A_NDIM;
  if (v_ndim < 2) return ND4REF_E_A_NDIM;
  if (y_ndim < 2) return ND4REF_E_B_NDIM;
  const int N = u_shape[u_ndim - 2], M = u_shape[u_ndim - 1], I = v_shape[v_ndim - 1], J = y_shape[y_ndim - 1];
  if (N != y_shape[y_ndim - 2]) return ND4REF_E_INNER;
  if (M != sv_shape[sv_ndim - 1]) return ND4REF_E_INNER;
  if (M != v_shape[v_ndim - 2]) return ND4REF_E_INNER;
  int ndim = u_ndim;
  if (sv_ndim + 1 > ndim) ndim = sv_ndim + 1;
  if (v_ndim > ndim) ndim = v_ndim;
  if (y_ndim > ndim) ndim = y_ndim;
  for (int d = 0; d < ndim; d++) x_shape[d] = 1;
  x_shape[ndim - 2] = I;
  x_shape[ndim - 1] = J;
  const int32_t* shp[4] = {u_shape, v_shape, y_shape, sv_shape};
  const int lead[4] = {u_ndim - 2, v_ndim - 2, y_ndim - 2, sv_ndim - 1};
  for (int w = 0; w < 4; w++)
    for (int i = ndim - 2, j = lead[w]; i-- > 0 && j-- > 0;) {
      if (x_shape[i] == 1) x_shape[i] = shp[w][j];
      else if (x_shape[i] != shp[w][j] && shp[w][j] != 1) return ND4REF_E_BROADCAST;
    }
  *x_ndim = ndim;
  return ND4REF_OK;
}

int nd4ref_svd_lstsq_f64(const double* U, const int32_t* u_shape, int u_ndim, const double* sv, const int32_t* sv_shape, int sv_ndim,
                         const double* V, const int32_t* v_shape, int v_ndim, const double* Y, const int32_t* y_shape, int y_ndim,
                         double* X, const int32_t* x_shape, int x_ndim) {
  int32_t want[64];
  int ndim = 0;
  const int rc = nd4ref_svd_lstsq_shape(u_shape, u_ndim, sv_shape, sv_ndim, v_shape, v_ndim, y_shape, y_ndim, want, &ndim);
  if (rc) return rc;
  if (ndim != x_ndim) return ND4REF_E_SHAPE;
  for (int d = 0; d < ndim; d++)
    if (want[d] != x_shape[d]) return ND4REF_E_SHAPE;
  const int N = u_shape[u_ndim - 2], M = u_shape[u_ndim - 1], I = v_shape[v_ndim - 1], J = y_shape[y_ndim - 1];
  const int nb = ndim - 2;
  const int32_t* shp[4] = {u_shape, sv_shape, v_shape, y_shape};
  const int lead[4] = {u_ndim - 2, sv_ndim - 1, v_ndim - 2, y_ndim - 2};
  const int64_t elems[4] = {(int64_t)N * M, M, (int64_t)M * I, (int64_t)N * J};
  int64_t str[4][64], idx[64], total = 1;
  for (int o = 0; o < 4; o++) {
    int64_t s = elems[o];
    for (int d = nb - 1; d >= 0; d--) {
      const int k = d - nb + lead[o];
      const int64_t n = k >= 0 ? shp[o][k] : 1;
      str[o][d] = n > 1 ? s : 0;
      s *= n;
    }
  }
  for (int d = 0; d < nb; d++) { idx[d] = 0; total *= x_shape[d]; }
  double* tmp = (double*)malloc(sizeof(double) * (size_t)M * J);
  if (!tmp) return ND4REF_E_SHAPE;
  memset(X, 0, sizeof(double) * (size_t)total * I * J);   /* new DTypeArray(...) is zero-filled, svd.js:151 */
  for (int64_t m = 0; m < total; m++) {
    int64_t off[4] = {0, 0, 0, 0};
    for (int o = 0; o < 4; o++)
      for (int d = 0; d < nb; d++) off[o] += idx[d] * str[o][d];
    const double *u = U + off[0], *s = sv + off[1], *v = V + off[2], *y = Y + off[3];
    double* x = X + m * (int64_t)I * J;
    /* rank cut, svd.js:163-174 */
    int R = M;
    {
      const double T = ND4REF_SQRT_EPS * fabs(s[0]);
      for (int r = 0; r < M; r++) {
        const double sv_r = fabs(s[r]);
        if (!isfinite(sv_r)) { free(tmp); return ND4REF_E_NAN_INPUT; }
        if (sv_r <= T) { R = r; break; }
      }
    }
    /* tmp = U.T @ y, svd.js:180-184 */
    for (int e = 0; e < M * J; e++) tmp[e] = 0.0;
    for (int k = 0; k < N; k++)
      for (int i = 0; i < R; i++)
        for (int j = 0; j < J; j++) tmp[J * i + j] += u[(int64_t)M * k + i] * y[(int64_t)J * k + j];
    /* tmp \= diag(sv), svd.js:186-189 */
    for (int i = 0; i < R; i++)
      for (int j = 0; j < J; j++) tmp[J * i + j] /= s[i];
    /* x = V.T @ tmp, svd.js:191-195 */
    for (int k = 0; k < R; k++)
      for (int i = 0; i < I; i++)
        for (int j = 0; j < J; j++) x[J * i + j] += v[(int64_t)I * k + i] * tmp[J * k + j];
    for (int d = nb - 1; d >= 0; d--) {
      if (++idx[d] < x_shape[d]) break;
      idx[d] = 0;
    }
  }
  free(tmp);
  return ND4REF_OK;
}
