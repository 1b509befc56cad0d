// Cholesky orthogonaliser of the generalised eigenproblem  F C = S C eps  (ScfBase.hh:291-339; the route LAPACK's
// dsygv takes, which is what the reference reaches through lazyten/armadillo):
//     S = L L^T,   X = L^-1 (lower triangular),   F' = X F X^T,   C = X^T Z.
// Replaces the Loewdin set-up (a full eigensolve of S) by a blocked right-looking factorisation whose flops are
// DMMA GEMMs, and makes every product with X triangular (half the flops, GemmParams::ktri).
//
//   per 64-column block:  chol_diag_kernel   L11 = chol(A11) and L11^-1, both in shared memory (one CTA per matrix)
//                         GEMM               L21 = A21 L11^-T
//                         GEMM (lower tiles) A22 -= L21 L21^T
//   inverse, block row i: GEMM               P = L(i, 0:i) X(0:i, 0:i)
//                         GEMM               X(i, 0:i) = -L_ii^-1 P
#include "chol.cuh"

#include "gemm.cuh"

namespace gscfb {

namespace {

constexpr int CB = kCholBlock;
constexpr int CLD = CB + 1;

// A11 (nb x nb at (j0, j0), lower triangle referenced) -> L11^-1 into X(j0.., j0..) (lower, zeros above) and into
// Linv (column-major, ld CB).  info[q] = 1 when a pivot is not positive.
__global__ void __launch_bounds__(256)
chol_diag_kernel(const double* __restrict__ Ab, int lda, long long sA, double* __restrict__ Xb, int ldx, long long sX,
                 double* __restrict__ Linvb, int j0, int nb, int* __restrict__ info) {
  extern __shared__ __align__(16) double chol_sm[];
  double* shL = chol_sm;             // shL[c*CLD + r]
  double* shX = chol_sm + CB * CLD;  // shX[c*CLD + r] = (L^-1)(r, c)
  __shared__ int s_bad;
  const long long q = blockIdx.x;
  const double* A = Ab + q * sA + (size_t)j0 * lda + j0;
  const int tid = threadIdx.x;
  if (tid == 0) s_bad = 0;
  for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
    const int c = idx / nb, r = idx - c * nb;
    shL[c * CLD + r] = r >= c ? A[(size_t)c * lda + r] : 0.0;
    shX[c * CLD + r] = 0.0;
  }
  __syncthreads();
  // Right-looking unblocked Cholesky with ONE barrier per column: thread (r = tid mod 64, cp = tid / 64) owns row r and
  // every 4th column of the trailing block; the scaled column k is formed on the fly (1 / L_kk from one reciprocal square
  // root) and only stored after everybody has used the unscaled one.  (The first version took 108 us per 64 x 64 block -
  // three barriers, a square root + divide and an integer division per updated element - 16 blocks in sequence were
  // 1.7 of the 2.9 ms of the n = 1024 set-up.)
  static_assert(CB == 64, "thread mapping below: 256 threads = 64 rows x 4 column phases");
  __shared__ double s_rd[CB];  // 1 / L(k, k)
  {
    const int r = tid & (CB - 1), cp = tid >> 6;
    for (int k = 0; k < nb; ++k) {
      const double d = shL[k * CLD + k];
      const bool ok = d > 0.0;
      const double inv = ok ? rsqrt(d) : 0.0;
      if (!ok && tid == 0) s_bad = 1;
      const bool mine = r > k && r < nb;
      const double lrk = mine ? shL[k * CLD + r] * inv : 0.0;
      if (mine) {
        for (int c = k + 1 + cp; c <= r; c += 4) {
          const double lck = shL[k * CLD + c] * inv;
          shL[c * CLD + r] = fma(-lrk, lck, shL[c * CLD + r]);
        }
      }
      __syncthreads();  // every read of the unscaled column k (and of its pivot) is done; column k + 1 is final
      if (cp == 0 && r >= k && r < nb) shL[k * CLD + r] = (r == k) ? d * inv : lrk;  // L_kk = d / sqrt(d)
      if (tid == 0) s_rd[k] = inv;
    }
  }
  __syncthreads();
  // inverse: column c by forward substitution, one quad of lanes per column
  {
    const int c = tid >> 2, part = tid & 3;
    const unsigned qmask = 0xFu << (tid & 28);
    if (c < nb) {
      if (part == 0) shX[c * CLD + c] = s_rd[c];
      __syncwarp(qmask);
      for (int r = c + 1; r < nb; ++r) {
        double s = 0.0;
        for (int k = c + part; k < r; k += 4) s = fma(shL[k * CLD + r], shX[c * CLD + k], s);
        s += __shfl_xor_sync(qmask, s, 1);
        s += __shfl_xor_sync(qmask, s, 2);
        if (part == 0) shX[c * CLD + r] = -s * s_rd[r];
        __syncwarp(qmask);
      }
    }
  }
  __syncthreads();
  double* X = Xb + q * sX + (size_t)j0 * ldx + j0;
  double* Linv = Linvb + q * (long long)CB * CB;
  for (int idx = tid; idx < nb * nb; idx += blockDim.x) {
    const int c = idx / nb, r = idx - c * nb;
    const double v = r >= c ? shX[c * CLD + r] : 0.0;
    X[(size_t)c * ldx + r] = v;
    Linv[c * CB + r] = v;
  }
  if (tid == 0 && s_bad && info) info[q] = 1;
}

__global__ void chol_zero_kernel(double* __restrict__ X, int ldx, long long sX, int n, int* __restrict__ info) {
  double* x = X + blockIdx.y * sX;
  const long long total = (long long)ldx * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    x[i] = 0.0;
  if (blockIdx.x == 0 && threadIdx.x == 0 && info) info[blockIdx.y] = 0;
}

}  // namespace

size_t cholesky_inverse_worksize(int n, int batch) {
  return ((size_t)batch * ((size_t)CB * padded_ld(n) + (size_t)CB * CB)) * sizeof(double);
}

int cholesky_inverse_batched(int n, double* A, int lda, long long sA, double* Lw, int ldl, long long sL, double* X,
                             int ldx, long long sX, void* work, int batch, int* info_dev, cudaStream_t st) {
  if (batch <= 0 || n <= 0) return GSCF_B200_OK;
  const int ldp = padded_ld(n);
  double* P = static_cast<double*>(work);                                   // [batch][CB * ldp]  (CB x n, ld CB)
  double* Linv = P + (size_t)batch * CB * ldp;                              // [batch][CB * CB]
  chol_zero_kernel<<<dim3(std::max(1, std::min(64, ceil_div(ldx * n, 4096))), batch), 256, 0, st>>>(X, ldx, sX, n, info_dev);
  count_launch();
  GSCF_CHECK_LAUNCH();
  const int nblk = ceil_div(n, CB);
  constexpr size_t kDiagSmem = 2 * (size_t)CB * CLD * sizeof(double);
  static PerDeviceOnce once;
  int dev;
  if (once.need(dev)) {
    GSCF_CUDA_TRY(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDiagSmem));
    once.done(dev);
  }
  // ---- factorisation: L (strictly lower blocks) into Lw, L_ii^-1 into the diagonal blocks of X ---------------------
  for (int b = 0; b < nblk; ++b) {
    const int j0 = b * CB, nb = std::min(CB, n - j0), m = n - j0 - nb;
    chol_diag_kernel<<<batch, 256, kDiagSmem, st>>>(A, lda, sA, X, ldx, sX, Linv, j0, nb, info_dev);
    count_launch();
    GSCF_CHECK_LAUNCH();
    if (m <= 0) break;
    // L21 = A21 L11^-T  -> Lw(j0+nb.., j0..)
    GemmParams g = gemm_params(m, nb, nb, 1.0, A + (size_t)j0 * lda + j0 + nb, lda, Linv, CB, 0.0,
                               Lw + (size_t)j0 * ldl + j0 + nb, ldl);
    g.sA = sA; g.sB = (long long)CB * CB; g.sC = sL;
    GSCF_TRY(launch_gemm(false, true, g, batch, m, nb, st));
    // A22 -= L21 L21^T (lower tiles)
    GemmParams u = gemm_params(m, m, nb, -1.0, Lw + (size_t)j0 * ldl + j0 + nb, ldl, Lw + (size_t)j0 * ldl + j0 + nb, ldl,
                               1.0, A + (size_t)(j0 + nb) * lda + j0 + nb, lda);
    u.sA = u.sB = sL; u.sC = sA; u.lower_only = 1;
    GSCF_TRY(launch_gemm(false, true, u, batch, m, m, st));
  }
  // ---- inverse: X(i, 0:i) = -L_ii^-1 (L(i, 0:i) X(0:i, 0:i)) --------------------------------------------------------
  for (int b = 1; b < nblk; ++b) {
    const int i0 = b * CB, nb = std::min(CB, n - i0);
    GemmParams g = gemm_params(nb, i0, i0, 1.0, Lw + i0, ldl, X, ldx, 0.0, P, CB);
    g.sA = sL; g.sB = sX; g.sC = (long long)CB * ldp;
    GSCF_TRY(launch_gemm(false, false, g, batch, nb, i0, st));
    GemmParams h = gemm_params(nb, i0, nb, -1.0, X + (size_t)i0 * ldx + i0, ldx, P, CB, 0.0, X + i0, ldx);
    h.sA = sX; h.sB = (long long)CB * ldp; h.sC = sX;
    GSCF_TRY(launch_gemm(false, false, h, batch, nb, i0, st));
  }
  return GSCF_B200_OK;
}

}  // namespace gscfb
