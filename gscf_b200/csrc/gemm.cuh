// FP64 tensor-core GEMM for sm_100a (DMMA.8x8x4 tiles fed by tensor-map TMA; 1-D bulk copies / cp.async where a launch
// does not qualify).
//
// Replaces every dense contraction the reference delegates to lazyten/BLAS: the
// orthogonalisation X^T F X and back-transformation X C' of the generalised eigensolve
// (ScfBase.hh:291-339), the density build (FockMatrix.cc:202), S C_o / F C_o and the outer
// products of the Pulay error (pulay_error.cc:36-43), F_prev C_o of the TODA trace
// (TruncatedOptDampScf.hh:332), and the Q x U merges of the divide & conquer eigensolver.
#pragma once
#include "common.cuh"

namespace gscfb {

// One GEMM of a group whose sizes are only known on the device (D&C merges after deflation).
struct GemmDesc {
  const double* A;
  const double* B;
  double* C;
  int m, n, k;
  int lda, ldb, ldc;
  double alpha, beta;
};

struct GemmParams {
  const double* A;
  const double* B;
  double* C;
  int m, n, k;
  int lda, ldb, ldc;
  long long sA, sB, sC;   // strides between batch entries (blockIdx.z)
  int kbatch;             // C = alpha * sum_{q<kbatch} op(A_q) op(B_q) + beta*C
  long long skA, skB;     // strides between the q slices
  double alpha, beta;
  const int* batch_map;   // optional: batch entry z works on problem batch_map[z]
  int entries;            // with batch_map: number of problems the operands hold (1 + the largest index the map may contain);
                          // 0 = unknown, which keeps the launch off the tensor-map kernels (see gemm.cu: the batch extent of
                          // a tensor map must not reach beyond the allocation)
  const GemmDesc* descs;  // optional: grouped mode, one descriptor per blockIdx.z
  int ktri;               // triangular operand: restrict the k range of every tile (kbatch == 1, no split-k)
                          //   1: k < n0 + BN   (op(B)(k, j) = 0 for k > j)      2: k < m0 + BM   (op(A)(i, k) = 0 for k > i)
                          //   3: k >= m0       (op(A)(i, k) = 0 for k < i)
  int lower_only;         // 1: skip tiles strictly above the diagonal (symmetric result, mirrored by the caller)
  int mirror;             // with lower_only (square C, beta == 0): the epilogue also stores the transposed entry,
                          //   +1: C(c, r) =  C(r, c)  (symmetric result, replaces a separate symmetrise pass)
                          //   -1: C(c, r) = -C(r, c), zero diagonal  (antisymmetric result: the Pulay error)
  double* C2;             // optional second output C2 = alpha2 * op(A) op(B) of the same product (no beta), strided like C
  int ldc2;
  long long sC2;
  double alpha2;
  int use_tma;            // operand staging level, set by launch_gemm from GSCF_B200_GEMM_TMA (0 cp.async ... 3 default:
                          // tensor maps for both operand layouts; see gemm.cu)
  int ksplit;             // > 1: split-k, blockIdx.z = entry * ksplit + split, partial tiles go to Cpart
  double* Cpart;          // [batch * ksplit][m * n] partial products (compact, ld = m)
};

inline GemmParams gemm_params(int m, int n, int k, double alpha, const double* A, int lda,
                              const double* B, int ldb, double beta, double* C, int ldc) {
  GemmParams p{};
  p.A = A; p.B = B; p.C = C; p.m = m; p.n = n; p.k = k;
  p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.sA = p.sB = p.sC = 0; p.kbatch = 1; p.skA = p.skB = 0;
  p.alpha = alpha; p.beta = beta; p.batch_map = nullptr; p.entries = 0; p.descs = nullptr;
  p.ksplit = 1; p.Cpart = nullptr; p.ktri = 0; p.lower_only = 0; p.use_tma = 0;
  p.mirror = 0; p.C2 = nullptr; p.ldc2 = 0; p.sC2 = 0; p.alpha2 = 0.0;
  return p;
}

// Launch C = alpha op(A) op(B) + beta C for `batch` entries (grid.z).  max_m/max_n bound the
// grid in grouped mode (device-side sizes); pass p.m/p.n otherwise.
int launch_gemm(bool ta, bool tb, const GemmParams& p, int batch, int max_m, int max_n,
                cudaStream_t stream);

// Convenience wrappers.
inline int gemm(bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int lda,
                const double* B, int ldb, double beta, double* C, int ldc, cudaStream_t st) {
  GemmParams p = gemm_params(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  return launch_gemm(ta, tb, p, 1, m, n, st);
}
// Split-k variant for small outputs with a long k (TN products of the back-transformation): `splits`
// CTAs share one output tile, partials are reduced by a second kernel.  work: batch*splits*m*n doubles.
int launch_gemm_splitk(bool ta, bool tb, GemmParams p, int batch, int splits, double* work, cudaStream_t stream);

inline int gemm_batched(bool ta, bool tb, int m, int n, int k, double alpha, const double* A,
                        int lda, long long sA, const double* B, int ldb, long long sB,
                        double beta, double* C, int ldc, long long sC, int batch,
                        const int* batch_map, cudaStream_t st) {
  GemmParams p = gemm_params(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
  p.sA = sA; p.sB = sB; p.sC = sC; p.batch_map = batch_map;
  return launch_gemm(ta, tb, p, batch, m, n, st);
}

// Pulay error E = fac [(S C_o)(F C_o)^T - (F C_o)(S C_o)^T]  (src/gscfmock/pulay_error.cc:25-44; fac = 2 closed shell),
// batched, in three launches and 6 n^2 o flops (the reference's form is four products, 8 n^2 o):
//   W = [S C_o | F C_o] and W2 = [F C_o | -S C_o] come out of the same two GEMMs (second output of the epilogue),
//   E = fac W W2^T is ONE k = 2 o product over the tiles on and below the diagonal; the epilogue stores every strictly
//   lower entry a second time, transposed with the opposite sign, and an exact zero on the diagonal (E is antisymmetric).
// work: per batch entry ldw x 4 o doubles (ldw = padded_ld(n)), entries stride_work apart.
struct PulayErrorArgs {
  int n, o;
  double fac;
  const double* S; int lds; long long sS;
  const double* F; int ldf; long long sF;
  const double* C; int ldc; long long sC;
  double* E; int lde; long long sE;
  double* work; long long swork;
  int batch;
  const int* batch_map;
  int entries;            // as GemmParams::entries
};
inline size_t pulay_error_work_doubles(int n, int o) { return (size_t)padded_ld(n) * 4 * (o > 0 ? o : 1); }
int pulay_error_fused(const PulayErrorArgs& a, cudaStream_t st);

}  // namespace gscfb
