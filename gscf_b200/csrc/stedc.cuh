// Divide & conquer eigensolver for symmetric tridiagonal matrices, batched (see stedc.cu).
#pragma once
#include "common.cuh"
#include "gemm.cuh"

namespace gscfb {

struct StedcWs {
  int n = 0, ld = 0, cap = 0, L = 0, nleaf = 1, lmax = 0;
  // per matrix (indexed by the matrix index q = batch_map[z])
  double* lam[2] = {nullptr, nullptr};  // [cap][n] eigenvalues, ping-pong over the levels
  double* dmod = nullptr;               // [cap][n] diagonal after the rank-one splits
  double* dwork = nullptr;              // [cap][n] eigenvalues of the children (modified by deflation)
  double* z = nullptr;                  // [cap][n]
  double* dlamda = nullptr;             // [cap][n] poles
  double* w = nullptr;                  // [cap][n] z components of the poles
  double* zhat = nullptr;               // [cap][n]
  double* rot_c = nullptr;              // [cap][n]
  double* rot_s = nullptr;              // [cap][n]
  double* rho = nullptr;                // [cap][nleaf]
  double* scale = nullptr;              // [cap] factor of the eigenvalues: max-norm of T (the D&C runs on T / norm) x outer
  const double* outer = nullptr;        // [cap] optional: what the caller divided the dense matrix by (1 if nothing)
  int* idx = nullptr;                   // [cap][n]
  int* ctype = nullptr;                 // [cap][n]
  int* ndcol = nullptr;                 // [cap][n]
  int* dcol = nullptr;                  // [cap][n]
  int* gpos = nullptr;                  // [cap][n]
  int* rot_a = nullptr;                 // [cap][n]
  int* rot_b = nullptr;                 // [cap][n]
  int* rank = nullptr;                  // [cap][n]
  int* meta = nullptr;                  // [cap][nleaf][8]: K, nrot, c1, c2, c3
  double* Zb = nullptr;                 // [cap][ld*n] second eigenvector buffer
  double* S = nullptr;                  // [cap][ld*n] delta(j,i) = d_j - lam_i, stored [j*ld + i]
  double* U = nullptr;                  // [cap][ld*n] eigenvectors of the rank-one problems (transposed)
  double* Gq = nullptr;                 // [cap][ld*n] gathered non-deflated columns
  // scratch indexed by launch slot z
  GemmDesc* descs = nullptr;            // [cap][nleaf]  (2 per merge, <= nleaf/2 merges)
};

// Carve the D&C workspace out of `base` starting at `offset`; returns the end offset.  With
// base == nullptr only the size is computed.
size_t stedc_carve(char* base, size_t offset, int n, int cap, StedcWs& ws);

// T = tridiag(d, e) (n entries each with stride dstride per matrix; e[0..n-2]; scaled in place to unit max-norm) ->
// w ascending (stride_w per matrix) and eigenvectors Z (n x n, ldz, strideZ).
int stedc_batched(int n, double* d, double* e, long long dstride, double* w,
                  long long stride_w, double* Z, int ldz, long long strideZ, const StedcWs& ws,
                  int batch, const int* batch_map, int* info_dev, cudaStream_t st, int nvec = -1);

}  // namespace gscfb
