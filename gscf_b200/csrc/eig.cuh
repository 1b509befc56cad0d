// Symmetric eigensolvers (all eigenpairs, ascending) - the FLOP sink of ScfBase::update_eigenpairs
// (ScfBase.hh:291-339).  Two hand-written paths, no cuSOLVER:
//   * jacobi.cu   one CTA per matrix, one-sided cyclic Jacobi in shared memory   (n <= 160; AUTO uses it for n <= 64)
//   * tridiag.cu (+ trd_fused*.cuh) + stedc.cu   fused one-launch Householder tridiagonalisation, divide & conquer on
//                 the tridiagonal matrix (DMMA GEMM merges), blocked back-transformation
#pragma once
#include "common.cuh"

namespace gscfb {

constexpr int kJacobiMaxN = 160;

size_t jacobi_smem_bytes(int n);
int jacobi_syev_batched(int n, const double* A, int lda, long long strideA, double* w,
                        long long stride_w, double* Z, int ldz, long long strideZ, int batch,
                        const int* batch_map, int* info_dev, cudaStream_t st);

// Workspace plan of the large path for `batch` matrices of order n.
struct DcPlan;
size_t tridiag_dc_worksize(int n, int batch);
// A (n x n, ld lda, FULL symmetric storage, destroyed) -> w ascending, Z (n x n) eigenvectors.
// batch_map (optional): entry z works on matrix batch_map[z] of every strided array.
// nvec (n_eigenpairs < n, ScfBase.hh:87): a partial-spectrum solve - the top-level merge of the divide & conquer forms
// eigenvectors only for the nvec lowest eigenvalues (three quarters of its flops otherwise) and only those nvec vectors are
// back-transformed; all n eigenvalues are always returned, columns >= nvec of Z are unspecified.
int tridiag_dc_syev_batched(int n, double* A, int lda, long long strideA, double* w,
                            long long stride_w, double* Z, int ldz, long long strideZ, int batch,
                            const int* batch_map, void* work, size_t work_bytes, int* info_dev,
                            cudaStream_t st, int nvec = -1);

// Mirror the lower triangle into the upper one (batched).
int symmetrise_from_lower(double* A, int lda, long long strideA, int n, int batch,
                          const int* batch_map, cudaStream_t st);

// Method dispatch used by the engine and by gscf_b200_syev.
int resolve_eig_method(int method, int n);
size_t syev_worksize(int n, int batch, int method);
int syev_batched(int method, int n, double* A, int lda, long long strideA, double* w,
                 long long stride_w, double* Z, int ldz, long long strideZ, int batch,
                 const int* batch_map, void* work, size_t work_bytes, int* info_dev,
                 cudaStream_t st, int nvec = -1);

}  // namespace gscfb
