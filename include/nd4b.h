/* nd4b.h — C ABI of libnd4b.so: the B200 (sm_100a) implementation of nd4js's broadcast-batched
 * Float64 dense linear-algebra hot path.
 *
 * The reference (nd4js v1.3.0) has no FFI layer: call sites import plain JS functions.  The
 * drop-in boundary is therefore "the JS function body": a JS shim with the reference's names and
 * signatures keeps asarray()/shape inference/error texts in JS and hands flat Float64Array /
 * Int32Array buffers to an N-API addon that binds exactly the entry points below
 * (see INTEGRATION.md for the shim and the addon).  Each entry point cites the reference
 * function it replaces (paths relative to the nd4js checkout).
 *
 * Conventions
 *   - plain pointers and sizes only; all matrices dense, row-major, batch-major (src/nd_array.js:128-147)
 *   - return 0 on success, <0 argument errors (same codes as the texts the reference throws),
 *     >0 numerical failure; nd4b_last_error() returns the reference's message text (thread-local)
 *   - "host" entry points take HOST pointers (pinned or pageable), block until the results are in
 *     the caller's buffers, never write to inputs, retain nothing; work is sharded over the
 *     context's devices by contiguous ranges of the flattened leading batch index
 *   - "dev" entry points take DEVICE pointers valid on `device`, enqueue on `stream`
 *     (a cudaStream_t passed as void*) and return without synchronising
 *   - there is NO CPU fallback: every entry point fails with ND4B_E_CUDA if no device is usable
 */
#ifndef ND4B_H
#define ND4B_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define ND4B_OK             0
#define ND4B_E_A_NDIM     (-1)  /* 'A must be at least 2D.'                         src/la/matmul.js:95   */
#define ND4B_E_B_NDIM     (-2)  /* 'B must be at least 2D.'                         src/la/matmul.js:96   */
#define ND4B_E_INNER      (-3)  /* 'The last dimension of A and the 2nd to last dimension of B do not match.' :101-102 */
#define ND4B_E_BROADCAST  (-4)  /* 'Shapes are not broadcast-compatible.'           src/la/matmul.js:116  */
#define ND4B_E_SHAPE      (-5)  /* result shape passed by the caller is not the broadcast shape           */
#define ND4B_E_NOT_SQUARE (-6)  /* 'Last two dimensions must be quadratic.'         src/la/cholesky.js:61 */
#define ND4B_E_NAN_INPUT  (-7)  /* 'Assertion failed.' (KahanSum.set on NaN)        src/kahan_sum.js:29   */
#define ND4B_E_ARG        (-8)  /* null pointer / non-positive dimension (NDArray rejects dims < 1, src/nd_array.js:138) */
#define ND4B_E_CUDA       (-9)  /* CUDA runtime failure or no usable device; text in nd4b_last_error()     */
#define ND4B_E_NO_CONVERGENCE (-10) /* Jacobi SVD hit the sweep limit (60 sweeps; not seen on any input so far)  */
#define ND4B_E_SINGULAR     1   /* 'Matrix contains NaNs or is (near) singular.'    src/la/cholesky.js:44 */

#define ND4B_MAX_NDIM 32

/* ---- context ------------------------------------------------------------------------------- */

/* Creates the process-wide context on the given CUDA device ordinals (NULL/0: the ND4B_DEVICES
 * environment variable, e.g. "0,1,2,3", else the current device).  Idempotent for an equal list. */
int nd4b_init(const int* devices, int n_devices);
int nd4b_shutdown(void);
int nd4b_device_count(void);            /* devices in the context (0 before init)            */
const char* nd4b_last_error(void);      /* message of the last failing call on this thread   */
const char* nd4b_version(void);

/* Page-locked host memory, so a caller (the N-API addon creates external ArrayBuffers from it)
 * can keep Float64Array storage DMA-able and skip the staging copy. */
void* nd4b_host_alloc(size_t bytes);
void nd4b_host_free(void* p);
/* Freed blocks are cached for reuse (page-locking costs more than the copy it saves; at most ND4B_PINNED_CACHE_MB,
 * default 8192, of idle blocks); this releases the idle ones. */
void nd4b_host_trim(void);

/* Bytes of pipeline chunk per device stream (default 32 MiB); also ND4B_CHUNK_MB in the environment. */
int nd4b_set_chunk_bytes(size_t bytes);

typedef struct nd4b_stats {
  uint64_t calls;            /* host entry-point calls                      */
  uint64_t kernel_launches;  /* kernels launched by this library            */
  uint64_t h2d_bytes;        /* bytes copied host -> device                 */
  uint64_t d2h_bytes;        /* bytes copied device -> host                 */
  uint64_t staged_bytes;     /* bytes that went through the pinned ring     */
  int32_t last_sweeps;       /* Jacobi sweeps of the last svd call (max over batch) */
  int32_t n_devices;
} nd4b_stats;
int nd4b_get_stats(nd4b_stats* out);
int nd4b_reset_stats(void);

/* ---- nd.la.matmul2 — src/la/matmul.js:91-147 (shape rules) + :31-74 (matmul2_RR) ------------- */

/* Broadcast result shape of matmul2(a,b); c_shape must hold max(a_ndim,b_ndim) entries. */
int nd4b_matmul_shape(const int32_t* a_shape, int a_ndim, const int32_t* b_shape, int b_ndim,
                      int32_t* c_shape, int* c_ndim);

/* C[...,I,J] = A[...,I,K] . B[...,K,J] with NumPy-style broadcasting of the leading dims. */
int nd4b_matmul_f64(const double* A, const int32_t* a_shape, int a_ndim,
                    const double* B, const int32_t* b_shape, int b_ndim,
                    double* C, const int32_t* c_shape, int c_ndim);

/* ---- nd.la.matmul(...matrices) — src/la/matmul.js:150-236 --------------------------------------- */

/* Executes a matrix-chain product with all intermediates resident in HBM.  The caller (the JS shim / la.py) keeps the
 * reference's ordering DP (:159-235) and passes the parenthesisation as a postfix plan: an entry >= 0 pushes operand
 * mats[entry] (shape shapes[entry], ndims[entry] dims), -1 replaces the two topmost items a, b (b on top) by matmul2(a, b)
 * with the broadcasting and error texts of nd4b_matmul_f64.  E.g. a.(b.c) = {0, 1, 2, -1, -1}.  c_shape must be the shape
 * of the final product. */
int nd4b_matmul_plan_f64(int n, const double* const* mats, const int32_t* const* shapes, const int* ndims,
                         const int32_t* plan, int plan_len, double* C, const int32_t* c_shape, int c_ndim);

/* ---- nd.la.cholesky_decomp — src/la/cholesky.js:50-72, _cholesky_decomp :27-47 --------------- */

/* Reads only the lower triangle of every S; L's strict upper triangle is +0.  On ND4B_E_SINGULAR
 * *first_bad is the flattened batch index of the first failing matrix (reference: the exception
 * aborts the whole call), L is then unspecified. */
int nd4b_cholesky_f64(const double* S, double* L, int64_t batch, int n, int64_t* first_bad);

/* ---- nd.la.qr_decomp — src/la/qr.js:80-145 (and :27-77 for rows <= cols) --------------------- */

/* Q[batch,rows,min(rows,cols)], R[batch,min(rows,cols),cols]; R exactly upper triangular with
 * diag(R) >= 0 (the reference's Givens QR leaves arbitrary signs on diag(R): compare sign-normalised). */
int nd4b_qr_f64(const double* A, double* Q, double* R, int64_t batch, int rows, int cols);

/* ---- nd.la._qr_decomp_inplace — src/la/qr.js:147-183 (the QR of the trust-region least-squares solvers,
 *      src/opt/_trust_region_solver_tls.js:1126): factorise A and apply Q^T to right-hand sides without forming Q -------- */

/* A[batch,M,N], Y[batch,M,L] -> R[batch,M,N] (upper trapezoidal, exact zeros below the diagonal, diag >= 0) and
 * QtY[batch,M,L] = Q^T Y with the full M x M orthogonal Q of A = Q R.  The reference works in place on A and Y; here the
 * inputs are read-only and the results go to R and QtY (which may not alias the inputs).  Rows min(M,N).. of QtY are
 * coordinates in an orthonormal basis of range(A)'s complement: only their norms are basis independent. */
int nd4b_qr_inplace_f64(const double* A, const double* Y, double* R, double* QtY, int64_t batch, int M, int N, int L);

/* ---- nd.la.qr_lstsq — src/la/qr.js:186-273, fused for thin factors ----------------------------------------------- */

/* x = argmin |Q R x - y| from the factors of qr_decomp: X[batch,I,J] = R[:L,:L]^-1 (Q[:, :L]^T Y), L = min(M,I), rows L..I-1
 * zero — Q[batch,N,M], R[batch,M,I], Y[batch,N,J], all with the same leading batch (the shim composes matmul2 and
 * triu_solve for broadcast operands and for factors with more than 32 columns).  Bit-identical with the reference: the
 * products of Q^T y are summed in its order (k ascending, unfused), the back substitution is _triu_solve's (tri.js:73-98). */
int nd4b_qr_lstsq_f64(const double* Q, const double* R, const double* Y, double* X, int64_t batch, int N, int M, int I, int J);

/* ---- nd.la.svd_jac_1sided — contract of the svd_jac_* family, src/la/svd_jac_2sided.js:30-144,
 *      ordering/sign rules src/la/_svd_jac_utils.js:123-188, shapes src/help.js:2321-2337 ------- */

/* U[batch,rows,L], sv[batch,L] (descending, >= +0), V[batch,L,cols], L=min(rows,cols),
 * A = U diag(sv) V.  *sweeps_out (may be NULL): max Jacobi sweeps over the batch.  NaN / Infinity in a matrix gives NaN
 * results for that matrix and ND4B_OK, as the reference's Jacobi loops do (every threshold test is false for NaN, so the
 * iteration stops at once: svd_jac_2sided.js:112).  Singular values below 2^-200 of the largest are reported as 0. */
int nd4b_svd_jac1_f64(const double* A, double* U, double* sv, double* V,
                      int64_t batch, int rows, int cols, int* sweeps_out);

/* ---- nd.la.svd_rank / svd_lstsq / svd_solve — src/la/svd.js:31-58, 103-226, 61-100 ------------------------------- */

/* rank[batch] (int32): number of leading singular values with |sv_r| > sqrt(eps) * |sv_0| (the scan stops at the first one
 * at or below the cut, svd.js:44-52).  ND4B_E_NAN_INPUT ('svd_rank(): NaN or Infinity encountered.') when a non-finite entry
 * is met before the cut, as the reference throws. */
int nd4b_svd_rank_f64(const double* sv, int32_t* rank, int64_t batch, int n);

/* Validation (the reference's checks and texts, svd.js:112-147) and broadcast result shape [..., I, J] of svd_lstsq for
 * U[...,N,M], sv[...,M], V[...,M,I], y[...,N,J]; x_shape must hold max(u_ndim, sv_ndim + 1, v_ndim, y_ndim) entries. */
int nd4b_svd_lstsq_shape(const int32_t* u_shape, int u_ndim, const int32_t* sv_shape, int sv_ndim,
                         const int32_t* v_shape, int v_ndim, const int32_t* y_shape, int y_ndim,
                         int32_t* x_shape, int* x_ndim);

/* X = V^T diag(1/sv[:R]) U^T Y with R the rank cut above, all four operands broadcast independently over the leading dims.
 * One fused kernel; bit-identical with the reference (the sums of svd.js:177-193 in their order, products and sums rounded
 * separately, IEEE division).  svd_solve is this call plus the squareness check, which stays in the shim (svd.js:74-75; the
 * reference's singularity scan :87-95 never runs because its loop variable starts undefined). */
int nd4b_svd_lstsq_f64(const double* U, const int32_t* u_shape, int u_ndim, const double* sv, const int32_t* sv_shape, int sv_ndim,
                       const double* V, const int32_t* v_shape, int v_ndim, const double* Y, const int32_t* y_shape, int y_ndim,
                       double* X, const int32_t* x_shape, int x_ndim);

/* ---- nd.la.tril_solve / triu_solve (src/la/tri.js:156-293) and nd.la.cholesky_solve (src/la/cholesky.js:75-144) ---- */

#define ND4B_TRIL_SOLVE     0   /* X = L^-1 Y, forward substitution  (_tril_solve, tri.js:45-71)          */
#define ND4B_TRIU_SOLVE     1   /* X = U^-1 Y, backward substitution (_triu_solve, tri.js:73-98)          */
#define ND4B_CHOLESKY_SOLVE 2   /* X = L^-T L^-1 Y                   (_tril_solve + _tril_t_solve)        */
/* T[...,M,M], Y[...,M,J] -> X[...,M,J], leading dims broadcast like matmul2.  Bit-exact with the reference
 * (same per-entry operation sequence).  Only the referenced triangle of T is read. */
int nd4b_tri_solve_f64(int op, const double* T, const int32_t* t_shape, int t_ndim,
                       const double* Y, const int32_t* y_shape, int y_ndim,
                       double* X, const int32_t* x_shape, int x_ndim);

/* ---- device-resident forms (inputs already in HBM; used for composition and kernel timing) --- */

/* Batched C[m] = A[m*a_stride] . B[m*b_stride]; strides in ELEMENTS between consecutive matrices,
 * 0 broadcasts the operand (the whole-operand case of matmul.js:59-67). */
int nd4b_dev_matmul_f64(int device, void* stream, const double* A, int64_t a_stride,
                        const double* B, int64_t b_stride, double* C,
                        int64_t batch, int I, int K, int J);
/* info (device int64, may be NULL): min batch index whose factorisation failed, else INT64_MAX;
 * the caller initialises it. */
int nd4b_dev_cholesky_f64(int device, void* stream, const double* S, double* L,
                          int64_t batch, int n, long long* info);
int nd4b_dev_qr_f64(int device, void* stream, const double* A, double* Q, double* R,
                    int64_t batch, int rows, int cols, double* workspace, size_t workspace_bytes);
/* op as in nd4b_tri_solve_f64; T[m*t_stride] is M x M, Y[m*y_stride] is M x J (strides in elements, 0 broadcasts), X[batch,M,J]. */
int nd4b_dev_tri_solve_f64(int device, void* stream, int op, const double* T, int64_t t_stride, const double* Y, int64_t y_stride,
                           double* X, int64_t batch, int M, int J);
int nd4b_dev_qr_lstsq_f64(int device, void* stream, const double* Q, const double* R, const double* Y, double* X,
                          int64_t batch, int N, int M, int I, int J);
int nd4b_dev_qr_inplace_f64(int device, void* stream, const double* A, const double* Y, double* R, double* QtY,
                            int64_t batch, int M, int N, int L);
/* All operands with the same batch; fail_flag (device int32, may be NULL) is set when a non-finite sv precedes the cut. */
int nd4b_dev_svd_lstsq_f64(int device, void* stream, const double* U, const double* sv, const double* V, const double* Y, double* X,
                           int64_t batch, int N, int M, int I, int J, int* fail_flag);
/* sweeps (device int32, may be NULL): atomicMax of sweeps used.  workspace as reported below. */
int nd4b_dev_svd_jac1_f64(int device, void* stream, const double* A, double* U, double* sv, double* V,
                          int64_t batch, int rows, int cols, int* sweeps,
                          double* workspace, size_t workspace_bytes);
/* Scratch bytes the dev_qr / dev_svd forms need for this problem (0 for the specialised shapes). */
size_t nd4b_dev_qr_workspace(int64_t batch, int rows, int cols);
size_t nd4b_dev_svd_workspace(int64_t batch, int rows, int cols);

/* NCCL over NVLink, used only to gather results: shard d — counts[d] doubles on the d-th device of the context, e.g.
 * the outputs of nd4b_dev_*_f64 on that device's contiguous range of the batch — is copied to offset counts[0] + .. + counts[d-1]
 * of full[e] on EVERY device e (full[e] must hold the sum of the counts; full[d] may contain shards[d] in place).  streams[d]
 * (cudaStream_t as void*, may be NULL: the library's stream) is the stream of device d the copies are ordered on.  One grouped
 * ncclBroadcast per shard; NCCL is dlopen()ed on first use (ND4B_NCCL_LIB overrides the library name) and never needed by a
 * one-device context.  The host-buffer entry points do not use it: their results land in host memory slice by slice. */
int nd4b_dev_all_gather_f64(const double* const* shards, const int64_t* counts, double* const* full, void* const* streams);

/* Diagnostic for the flop accounting of the benchmark: from now on every SVD launch on `device` adds the sweep count of
 * each matrix to *counter (device uint64, caller-initialised; NULL switches it off).  The Jacobi sweep count is data
 * dependent (C5: 10.06 on average, 12 at most), and work per matrix is proportional to it. */
int nd4b_dev_svd_sweep_counter(int device, unsigned long long* counter);
/* The same for the single-precision sweeps of the 64 x 64 preconditioner (csrc/svd_pre.cu), which runs when the dev / host
 * call has its workspace (nd4b_dev_svd_workspace); ND4B_SVD_PRE=0 in the environment switches the preconditioner off. */
int nd4b_dev_svd_pre_sweep_counter(int device, unsigned long long* counter);

/* ---- diagnostics ------------------------------------------------------------------------------ */

/* Times one launch of an FP64 pipe probe on `device` (which: 0 = DFMA vector pipe, 16 chains x 2 flop x iters
 * per thread; 1 = DMMA.8x8x4 tensor pipe, 8 tiles x 512 flop x iters per warp).  Used to measure the FP64
 * roofline denominators that MEASURED_PEAKS.json lacks. */
int nd4b_probe_fp64(int device, int which, int iters, int blocks, int threads, float* ms_out);

/* Self-check of the kernels' branch-free IEEE helpers on `device`: draws `samples` pseudo-random doubles over all
 * exponents (plus values at the limits of the fast-path ranges) and compares the inlined sqrt / division fast paths
 * (csrc/common.cuh: what the bit-exact Cholesky and triangular-solve kernels execute) with the compiler's sqrt() and `/`.
 * counts[0] = square roots that differ, counts[1] = quotients that differ (both must be 0), counts[2], counts[3] = how
 * many samples took the fast paths. */
int nd4b_selfcheck_ieee(int device, long long samples, unsigned long long seed, unsigned long long counts[4]);

#ifdef __cplusplus
}
#endif
#endif
