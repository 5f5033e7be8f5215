/* nd4ref.h — CPU ORACLE for the nd4js batched dense-LA hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (libnd4b.so) never links or calls anything in oracle/.
 *
 * Each function is a plain-C restatement (single thread, IEEE-754 binary64, no FMA
 * contraction: build with -ffp-contract=off) of one nd4js v1.3.0 routine; the cited
 * file:line ranges are relative to the reference checkout.
 *
 * Pinning status: PINNED BY THE REFERENCE ITSELF.  The reference's own JavaScript (the src/la sources of
 * nd4js v1.3.0) runs in the build container inside QJSEngine (Qt 6.6.3, shipped with Nsight Compute; oracle/jsref/qjs.py);
 * oracle/jsref/gen_golden.py commits its outputs on 81 seeded cases as tests/golden/jsref_golden.npz, and
 * tests/test_jsref_golden.py checks that every function below reproduces them BIT FOR BIT (matmul2, matmul chains,
 * cholesky_decomp incl. the failure texts, tril/triu/cholesky_solve, qr_decomp, qr_decomp_full, _qr_decomp_inplace,
 * qr_lstsq, svd_jac_2sided, svd_rank, svd_lstsq, svd_solve); tests/jsref_live_cases.py repeats it on fresh seeds against
 * the live engine.  The reference's own known answers are checked as well (matmul_test.js:32-78, help.js:1876-1885,
 * _generic_test_svd_decomp.js:180-216) and its property suites with its tolerances.
 * nd.la.svd_jac_1sided does not exist in the reference snapshot: for it PARITY IS UNPINNED by the reference; the
 * reference-pinned svd_jac_2sided supplies the singular values it is compared with.
 */
#ifndef ND4REF_H
#define ND4REF_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* status codes shared with include/nd4b.h */
#define ND4REF_OK                 0
#define ND4REF_E_A_NDIM         (-1)  /* 'A must be at least 2D.'                        matmul.js:95  */
#define ND4REF_E_B_NDIM         (-2)  /* 'B must be at least 2D.'                        matmul.js:96  */
#define ND4REF_E_INNER          (-3)  /* 'The last dimension of A and the 2nd to last…'  matmul.js:101 */
#define ND4REF_E_BROADCAST      (-4)  /* 'Shapes are not broadcast-compatible.'          matmul.js:116 */
#define ND4REF_E_SHAPE          (-5)  /* caller-supplied result shape does not match                   */
#define ND4REF_E_NOT_SQUARE     (-6)  /* 'Last two dimensions must be quadratic.'        cholesky.js:61 */
#define ND4REF_E_NAN_INPUT      (-7)  /* KahanSum.set 'Assertion failed.'                kahan_sum.js:29 */
#define ND4REF_E_SINGULAR         1   /* 'Matrix contains NaNs or is (near) singular.'   cholesky.js:44 */

/* matmul2: shape inference, src/la/matmul.js:91-116.  c_shape must hold max(a_ndim,b_ndim) ints. */
int nd4ref_matmul_shape(const int32_t* a_shape, int a_ndim,
                        const int32_t* b_shape, int b_ndim,
                        int32_t* c_shape, int* c_ndim);

/* matmul2_RR: src/la/matmul.js:31-74 (i-k-j order, separate mul and add, C zeroed first). */
int nd4ref_matmul_f64(const double* A, const int32_t* a_shape, int a_ndim,
                      const double* B, const int32_t* b_shape, int b_ndim,
                      double* C, const int32_t* c_shape, int c_ndim);

/* cholesky_decomp: src/la/cholesky.js:27-72 + src/kahan_sum.js:19-42.
 * returns 0, ND4REF_E_SINGULAR (first_bad = batch index of first failing matrix) or <0. */
int nd4ref_cholesky_f64(const double* S, double* L, int64_t batch, int n, int64_t* first_bad);

/* qr_decomp: src/la/qr.js:80-145 (rows>cols: Givens, thin) and :27-77 (rows<=cols: qr_decomp_full).
 * Q is [batch,rows,min(rows,cols)], R is [batch,min(rows,cols),cols]. */
int nd4ref_qr_f64(const double* A, double* Q, double* R, int64_t batch, int rows, int cols);
/* qr_decomp_full for any shape: src/la/qr.js:27-77.  Q [batch,rows,rows], R [batch,rows,cols]. */
int nd4ref_qr_full_f64(const double* A, double* Q, double* R, int64_t batch, int rows, int cols);

/* _qr_decomp_inplace: src/la/qr.js:147-183, batched and out of place: R[batch,M,N] <- A rotated to upper trapezoidal form,
 * QtY[batch,M,L] <- the same Givens rotations applied to Y. */
int nd4ref_qr_inplace_f64(const double* A, const double* Y, double* R, double* QtY, int64_t batch, int M, int N, int L);

/* svd_jac_2sided: src/la/svd_jac_2sided.js:30-144 + src/la/_svd_jac_utils.js:72-188.
 * U [batch,rows,L], sv [batch,L], V [batch,L,cols], L=min(rows,cols); A = U diag(sv) V.
 * sweeps_out (may be NULL) receives the maximum number of sweeps over the batch. */
int nd4ref_svd_jac2_f64(const double* A, double* U, double* sv, double* V,
                        int64_t batch, int rows, int cols, int* sweeps_out);

/* tril_solve / triu_solve (src/la/tri.js:156-293) and cholesky_solve (src/la/cholesky.js:75-144):
 * op 0, 1, 2.  T[...,M,M], Y[...,M,J] -> X[...,M,J] with broadcast leading dims. */
int nd4ref_tri_solve_f64(int op, const double* T, const int32_t* t_shape, int t_ndim,
                         const double* Y, const int32_t* y_shape, int y_ndim,
                         double* X, const int32_t* x_shape, int x_ndim);

/* svd_rank (src/la/svd.js:31-58): rank[batch] int32.  svd_lstsq (src/la/svd.js:103-226): x = V^T diag(1/sv[:R]) U^T y with
 * the rank cut R, four independently broadcast operands; returns ND4REF_E_NAN_INPUT where the reference throws
 * 'svd_solve(): NaN or Infinity encountered.' / 'svd_rank(): NaN or Infinity encountered.' */
int nd4ref_svd_rank_f64(const double* sv, int32_t* rank, int64_t batch, int n);
int nd4ref_svd_lstsq_shape(const int32_t* u_shape, int u_ndim, const int32_t* sv_shape, int sv_ndim,
                           const int32_t* v_shape, int v_ndim, const int32_t* y_shape, int y_ndim,
                           int32_t* x_shape, int* x_ndim);
int nd4ref_svd_lstsq_f64(const double* U, const int32_t* u_shape, int u_ndim, const double* sv, const int32_t* sv_shape, int sv_ndim,
                         const double* V, const int32_t* v_shape, int v_ndim, const double* Y, const int32_t* y_shape, int y_ndim,
                         double* X, const int32_t* x_shape, int x_ndim);

/* scalar helpers exposed for unit tests */
void nd4ref_giv_rot_qr(double a, double b, double out_c_s_norm[3]);      /* _giv_rot.js:22-37   */
void nd4ref_svd_jac_angles(double Spp, double Spq, double Sqp, double Sqq,
                           double out_ca_sa_cb_sb[4]);                   /* _svd_jac_utils.js:72-114 */
double nd4ref_frobenius(const double* x, int64_t n);                     /* norm.js:22-68       */

#ifdef __cplusplus
}
#endif
#endif
