// Eigensolver dispatch (see eig.cuh).
#include "eig.cuh"

#include "gemm.cuh"

namespace gscfb {

__global__ void symmetrise_kernel(double* __restrict__ A, int lda, long long strideA, int n,
                                  const int* __restrict__ batch_map) {
  const long long p = batch_map ? batch_map[blockIdx.z] : blockIdx.z;
  double* a = A + p * strideA;
  // 32x32 tiles of the strict lower triangle, transposed through shared memory
  __shared__ double tile[32][33];
  const int bi = blockIdx.x, bj = blockIdx.y;  // tile (bi, bj) with bi >= bj holds lower part
  if (bi < bj) return;
  const int r0 = bi * 32, c0 = bj * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + threadIdx.x, c = c0 + j;
    tile[j][threadIdx.x] = (r < n && c < n) ? a[(size_t)c * lda + r] : 0.0;
  }
  __syncthreads();
  // write transposed: element (c0 + x, r0 + j) = lower(r0 + j, c0 + x)
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int rr = c0 + threadIdx.x;  // row in upper part
    const int cc = r0 + j;            // column in upper part
    if (rr < n && cc < n && cc > rr) a[(size_t)cc * lda + rr] = tile[threadIdx.x][j];
  }
}

int symmetrise_from_lower(double* A, int lda, long long strideA, int n, int batch,
                          const int* batch_map, cudaStream_t st) {
  if (batch <= 0) return GSCF_B200_OK;
  const int t = ceil_div(n, 32);
  symmetrise_kernel<<<dim3(t, t, batch), dim3(32, 8), 0, st>>>(A, lda, strideA, n, batch_map);
  count_launch();
  GSCF_CHECK_LAUNCH();
  return GSCF_B200_OK;
}

int resolve_eig_method(int method, int n) {
  // measured on B200 (4096 problems of order 128): Jacobi 25.6 us per eigensolve amortised, tridiagonalisation in one
  // CTA's shared memory + divide & conquer 11.9 us; Jacobi keeps the tiny problems (and the D&C leaves)
  if (method == GSCF_B200_EIG_AUTO) return n <= 64 ? GSCF_B200_EIG_JACOBI : GSCF_B200_EIG_TRIDIAG_DC;
  if (method == GSCF_B200_EIG_JACOBI) {
    if (n > kJacobiMaxN) {
      set_last_error("Jacobi eigensolver supports n <= 160");
      return -1;
    }
    return method;
  }
  if (method == GSCF_B200_EIG_TRIDIAG_DC) return method;
  set_last_error("unknown eigensolver method");
  return -1;
}

size_t syev_worksize(int n, int batch, int method) {
  const int m = resolve_eig_method(method, n);
  if (m == GSCF_B200_EIG_TRIDIAG_DC) return tridiag_dc_worksize(n, batch);
  return 0;
}

int syev_batched(int method, int n, double* A, int lda, long long strideA, double* w,
                 long long stride_w, double* Z, int ldz, long long strideZ, int batch,
                 const int* batch_map, void* work, size_t work_bytes, int* info_dev,
                 cudaStream_t st, int nvec) {
  const int m = resolve_eig_method(method, n);
  if (m < 0) return GSCF_B200_ERR_INVALID_ARG;
  if (m == GSCF_B200_EIG_JACOBI)
    return jacobi_syev_batched(n, A, lda, strideA, w, stride_w, Z, ldz, strideZ, batch, batch_map,
                               info_dev, st);
  return tridiag_dc_syev_batched(n, A, lda, strideA, w, stride_w, Z, ldz, strideZ, batch,
                                 batch_map, work, work_bytes, info_dev, st, nvec);
}

}  // namespace gscfb
