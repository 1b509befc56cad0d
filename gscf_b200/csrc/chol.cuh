// Blocked Cholesky factorisation + triangular inverse (the orthogonaliser X = L^-1 of S = L L^T), batched.
#pragma once
#include "common.cuh"

namespace gscfb {

constexpr int kCholBlock = 64;

// scratch bytes for `batch` matrices of order n
size_t cholesky_inverse_worksize(int n, int batch);

// A (n x n, lower triangle referenced, destroyed) -> X = L^-1 (lower triangular, zeros above the diagonal).
// Lw: n x n scratch per matrix (receives the strictly-lower blocks of L).  info_dev[z] = 1 if S is not positive definite.
int cholesky_inverse_batched(int n, double* A, int lda, long long sA, double* Lw, int ldl, long long sL, double* X,
                             int ldx, long long sX, void* work, int batch, int* info_dev, cudaStream_t st);

}  // namespace gscfb
