// kernels.h — host-side launch interface of the nd4b CUDA kernels (internal; the public ABI is include/nd4b.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace nd4b {

// Broadcast odometer of matmul2 (nd4js src/la/matmul.js:44-70) in closed form: the flattened batch
// index m of C is decomposed over size[0..nd) (row-major) and mapped to operand element offsets by
// per-dim strides; a stride of 0 repeats the operand along that dim (dim is 1 or absent).
struct BatchMap {
  int nd;             // number of (collapsed) leading dims used by the odometer, 0..8
  int64_t base;       // global batch index of the first matrix of this launch (sharding / chunking)
  int64_t a_lin;      // >= 0: operand A is addressed linearly, offset = local_index * a_lin (0 = one matrix for all)
  int64_t b_lin;      //  < 0: operand is resident as a whole and addressed through the odometer with the global index
  int64_t size[8];
  int64_t a_str[8];
  int64_t b_str[8];
};

// The same odometer for the four independently broadcast operands of svd_lstsq (U, sv, V, y; src/la/svd.js:201-218).
struct BatchMap4 {
  int nd;
  int64_t base;
  int64_t lin[4];     // >= 0: offset = local_index * lin; < 0: resident operand addressed through the odometer
  int64_t size[8];
  int64_t str[4][8];
};

cudaError_t launch_matmul(cudaStream_t s, const double* A, const double* B, double* C,
                          int64_t batch, int I, int K, int J, const BatchMap& map, int sm_count);

// info: device int64, atomicMin of (base_index + m) over matrices whose factorisation produced a NaN pivot.
cudaError_t launch_cholesky(cudaStream_t s, const double* S, double* L, int64_t batch, int n,
                            long long* info, long long base_index);

// op 0 tril_solve, 1 triu_solve, 2 cholesky_solve; map.a_* addresses T (M*M per matrix), map.b_* addresses Y (M*J).
cudaError_t launch_tri_solve(cudaStream_t s, int op, const double* T, const double* Y, double* X,
                             int64_t batch, int M, int J, const BatchMap& map);

// qr_lstsq for thin factors (M, I <= 32), same batch for Q[N,M], R[M,I], Y[N,J]; X[I,J]; bit-exact with the reference
cudaError_t launch_qr_lstsq(cudaStream_t s, const double* Q, const double* R, const double* Y, double* X,
                            int64_t batch, int N, int M, int I, int J);

// svd_lstsq (src/la/svd.js:103-226): X[batch,I,J] from U[N,M], sv[M], V[M,I], Y[N,J] addressed through `map`; *fail is set
// when a non-finite singular value is met before the rank cut.  svd_rank (src/la/svd.js:31-58): rank[batch] (int32).
cudaError_t launch_svd_lstsq(cudaStream_t s, const double* U, const double* SV, const double* V, const double* Y, double* X,
                             int64_t batch, int N, int M, int I, int J, const BatchMap4& map, int* fail);
cudaError_t launch_svd_rank(cudaStream_t s, const double* SV, int* rank, int64_t batch, int M, int* fail);

size_t qr_workspace_bytes(int64_t batch, int rows, int cols);
cudaError_t launch_qr(cudaStream_t s, const double* A, double* Q, double* R, int64_t batch, int rows, int cols,
                      double* work, size_t work_bytes);

// rows=64, cols=32 on the FP64 tensor pipe (qr_blocked.cu); no alignment requirement beyond 8 bytes
cudaError_t launch_qr64x32_blocked(cudaStream_t s, const double* A, double* Q, double* R, int64_t batch, int variant);
cudaError_t launch_qr_inplace_blocked(cudaStream_t s, const double* A, const double* Y, double* R, double* QtY,
                                      int64_t batch, int rows, int cols, int nrhs);
cudaError_t launch_qr_padded_blocked(cudaStream_t s, const double* A, double* Q, double* R, int64_t batch, int rows, int cols);

// _qr_decomp_inplace: R[M,N] (zero below the diagonal) and Q^T Y [M,L] without forming Q
cudaError_t launch_qr_inplace(cudaStream_t s, const double* A, const double* Y, double* R, double* QtY,
                              int64_t batch, int M, int N, int L);

size_t svd_workspace_bytes(int64_t batch, int rows, int cols);
// sweeps: device int32 (atomicMax).  fail: device int32 set to 1 if some matrix hit the sweep limit.
cudaError_t launch_svd_jac1(cudaStream_t s, const double* A, double* U, double* sv, double* V,
                            int64_t batch, int rows, int cols, int* sweeps, int* fail,
                            double* work, size_t work_bytes);

// 64x64 preconditioner (svd_pre.cu): FP32 Jacobi, then V1 = orthogonalised rotation and G1 = A V1 into `work`
// (svd64_pre_workspace_bytes(batch) bytes); *sweep_sum (may be null) accumulates the FP32 sweeps.
size_t svd64_pre_workspace_bytes(int64_t batch);
cudaError_t launch_svd64_pre(cudaStream_t s, const double* A, int64_t batch, double* work, unsigned long long* sweep_sum,
                             const double** G1, const double** V1);

// diagnostic: per-matrix sweep counts of the following SVD launches on `device` are added to *counter (device memory)
void set_svd_sweep_counter(int device, unsigned long long* counter);
void set_svd_pre_sweep_counter(int device, unsigned long long* counter);   // the same for the FP32 sweeps of the preconditioner

// fp64 peak probes (tools/ and bench use them to measure the FP64 roofline denominators)
cudaError_t launch_probe_dfma(cudaStream_t s, double* out, int iters, int blocks, int threads);
cudaError_t launch_probe_dmma(cudaStream_t s, double* out, int iters, int blocks, int threads);
cudaError_t launch_selfcheck(cudaStream_t s, long long per_thread, unsigned long long seed, unsigned long long* out,
                             int blocks, int threads);

}  // namespace nd4b
