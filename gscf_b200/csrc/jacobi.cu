// Batched symmetric eigensolver for small matrices (n <= 160): one CTA per matrix, the whole
// solve in shared memory, one warp per column pair.
//
// Algorithm: one-sided (Hestenes) cyclic Jacobi on G = A + sigma*I.  sigma = 2 * Gershgorin
// radius makes G positive definite with spectrum in [R, 3R], so G = U diag(lambda + sigma) U^T
// and orthogonalising the columns of G by plane rotations (G <- G J) converges to
// G J = U diag(lambda+sigma): column norms give the eigenvalues, normalised columns the
// eigenvectors, both to absolute accuracy O(eps * ||A||) - what the 1e-9 / 1e-8 parity
// targets of the SCF need.  One-sided Jacobi needs a single n x n array (128 KiB at n = 128);
// the two-sided form needs A and V (256 KiB > 227 KiB of shared memory per CTA).
//
// Used for: small problems (n <= 64 under EIG_AUTO: the reference example, n = 14) and, on request (EIG_JACOBI), batched
// ensembles up to n = 160; the n = 128 ensemble of BASELINE config 4 runs faster through tridiagonalisation + D&C.
#include "eig.cuh"

namespace gscfb {

constexpr int kJacobiMaxSweeps = 40;

__device__ __forceinline__ long long sort_key(double x) {
  const long long b = __double_as_longlong(x);
  return b < 0 ? (b ^ 0x7fffffffffffffffLL) : b;
}

template <int NPL, int PPW>
__global__ void __launch_bounds__(1024)
jacobi_onesided_kernel(const double* __restrict__ Ain, int lda, long long strideA,
                       double* __restrict__ w, long long stride_w, double* __restrict__ Z,
                       int ldz, long long strideZ, int n, const int* __restrict__ batch_map,
                       int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  const int ne = (n + 1) & ~1;         // even number of (possibly one dummy) columns
  const int ldg = n;                   // column stride in shared memory
  double* G = sm;                      // ne * ldg
  double* lam = G + (size_t)ne * ldg;  // ne
  __shared__ double s_red[33];
  __shared__ double s_sigma, s_scale;

  const long long prob = batch_map ? batch_map[blockIdx.x] : blockIdx.x;
  const double* A = Ain + prob * strideA;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;

  // ---- load the full symmetric matrix from the lower triangle; Gershgorin radius ------------
  double rmax = 0.0;
  for (int idx = tid; idx < ne * ldg; idx += nthr) {
    const int c = idx / ldg, r = idx - c * ldg;
    double v = 0.0;
    if (c < n) v = (r >= c) ? A[(size_t)c * lda + r] : A[(size_t)r * lda + c];
    G[idx] = v;
  }
  __syncthreads();
  for (int c = tid; c < n; c += nthr) {
    double s = 0.0;
    for (int r = 0; r < n; ++r) s += fabs(G[(size_t)c * ldg + r]);
    rmax = fmax(rmax, s);
  }
  rmax = warp_max(rmax);
  if (lane == 0) s_red[warp] = rmax;
  __syncthreads();
  if (tid == 0) {
    double m = 0.0;
    for (int i = 0; i < nwarp; ++i) m = fmax(m, s_red[i]);
    // dsyev's scaling: the squared column norms of a matrix far from unit norm leave the double range
    const bool ext = m > 0.0 && isfinite(m) && (m < 1e-100 || m > 1e100);
    s_scale = ext ? m : 1.0;
    if (ext) m = 1.0;
    s_sigma = (m > 0.0 && isfinite(m)) ? 2.0 * m : 1.0;
  }
  __syncthreads();
  const double sigma = s_sigma, ascale = s_scale;
  if (ascale != 1.0) {
    for (int idx = tid; idx < ne * ldg; idx += nthr) G[idx] /= ascale;
    __syncthreads();
  }
  for (int c = tid; c < n; c += nthr) G[(size_t)c * ldg + c] += sigma;
  __syncthreads();

  const double tol = sqrt((double)n) * 2.220446049250313e-16;
  const int npairs = ne >> 1;
  // Every step has three phases separated by block barriers:
  //   1. one warp per column pair: a = |g_p|^2, b = |g_q|^2, c = g_p.g_q (columns stay in registers)
  //   2. one THREAD per pair: rotation angle (sqrt / divisions).  Done per warp this scalar chain is ~100 FP64
  //      instructions executed 32-fold redundantly and it saturated the FP64 pipe (64 lanes / clock / SM on B200)
  //   3. the warps apply the rotations from registers
  // PPW = pairs per warp (template): 1 for n <= 64, 2 for n <= 128, 3 for n <= 160 (32 warps)
  __shared__ double s_a[kJacobiMaxN / 2 + 1], s_b[kJacobiMaxN / 2 + 1], s_c[kJacobiMaxN / 2 + 1];
  __shared__ double s_cs[kJacobiMaxN / 2 + 1], s_sn[kJacobiMaxN / 2 + 1];
  int sweeps = 0;
  bool converged = false;
  for (; sweeps < kJacobiMaxSweeps; ++sweeps) {
    int rotated = 0;
    for (int step = 0; step < ne - 1; ++step) {
      double xp[PPW][NPL], xq[PPW][NPL];
      int pp[PPW], qq[PPW];
#pragma unroll
      for (int u = 0; u < PPW; ++u) {
        const int pr = warp + u * nwarp;
        pp[u] = -1; qq[u] = -1;
        if (pr < npairs) {
          // round-robin tournament: index ne-1 is fixed, the others rotate
          int p, q;
          if (pr == 0) {
            p = ne - 1;
            q = step;
          } else {
            p = (step + pr) % (ne - 1);
            q = (step - pr + (ne - 1)) % (ne - 1);
          }
          if (p > q) { const int t = p; p = q; q = t; }
          pp[u] = p; qq[u] = q;
          const double* gp = G + (size_t)p * ldg;
          const double* gq = G + (size_t)q * ldg;
          double a = 0.0, b = 0.0, c = 0.0;
#pragma unroll
          for (int i = 0; i < NPL; ++i) {
            const int r = lane + 32 * i;
            xp[u][i] = r < n ? gp[r] : 0.0;
            xq[u][i] = r < n ? gq[r] : 0.0;
            a = fma(xp[u][i], xp[u][i], a);
            b = fma(xq[u][i], xq[u][i], b);
            c = fma(xp[u][i], xq[u][i], c);
          }
          // three sums with 7 exchanges: lanes split the values after the first butterfly step
          {
            const bool h16 = lane & 16, h8 = lane & 8;
            // step 1 (xor 16): low half keeps (a, b), high half keeps (c, 0)
            const double s0 = h16 ? a : c, s1 = h16 ? b : 0.0;
            const double k0 = h16 ? c : a, k1 = h16 ? 0.0 : b;
            const double r0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
            const double r1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
            // step 2 (xor 8): keep one value
            double v = (h8 ? r1 : r0) + __shfl_xor_sync(0xffffffffu, h8 ? r0 : r1, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            // lanes 0-7: a, 8-15: b, 16-23: c
            if (lane == 0) s_a[pr] = v;
            if (lane == 8) s_b[pr] = v;
            if (lane == 16) s_c[pr] = v;
          }
        }
      }
      __syncthreads();
      if (tid < npairs) {
        const double a = s_a[tid], b = s_b[tid], c = s_c[tid];
        double cs = 1.0, sn = 0.0;
        if (fabs(c) > tol * sqrt(a * b) && a > 0.0 && b > 0.0) {
          rotated = 1;
          const double zeta = (b - a) / (2.0 * c);
          const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
          cs = 1.0 / sqrt(1.0 + t * t);
          sn = cs * t;
        }
        s_cs[tid] = cs;
        s_sn[tid] = sn;
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < PPW; ++u) {
        const int pr = warp + u * nwarp;
        if (pr < npairs) {
          const double cs = s_cs[pr], sn = s_sn[pr];
          if (sn != 0.0) {
            double* gp = G + (size_t)pp[u] * ldg;
            double* gq = G + (size_t)qq[u] * ldg;
#pragma unroll
            for (int i = 0; i < NPL; ++i) {
              const int r = lane + 32 * i;
              if (r < n) {
                gp[r] = cs * xp[u][i] - sn * xq[u][i];
                gq[r] = sn * xp[u][i] + cs * xq[u][i];
              }
            }
          }
        }
      }
      __syncthreads();
    }
    if (!__syncthreads_or(rotated)) {
      converged = true;
      ++sweeps;
      break;
    }
  }

  // ---- eigenvalues = column norms - sigma; normalise columns --------------------------------
  for (int c = warp; c < n; c += nwarp) {
    double* g = G + (size_t)c * ldg;
    double s = 0.0;
    for (int r = lane; r < n; r += 32) s = fma(g[r], g[r], s);
    s = warp_sum(s);
    const double nrm = sqrt(s);
    const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    for (int r = lane; r < n; r += 32) g[r] *= inv;
    if (lane == 0) lam[c] = (nrm - sigma) * ascale;
  }
  __syncthreads();
  // ---- eigenvalues refined as Rayleigh quotients z^T A z with the ORIGINAL matrix -----------------------------
  // A column norm minus sigma carries the rounding of numbers of size sigma ~ 2 x (Gershgorin radius), an order of
  // magnitude above eps ||A||_2: harmless for the 1e-9 gate of the DIIS path, but TruncatedOptDampScf forms its line
  // search from sums of occupied orbital energies (TruncatedOptDampScf.hh:309-319) against total energies, where 1e-14
  // decides the branch near convergence.  The quotient is accurate to a few eps ||A||_2 (its error is quadratic in
  // the eigenvector error); the input matrix is still intact in global memory (L2) - n^3 fused multiply-adds.
  {
    const double inva = 1.0 / ascale;  // exactly 1 unless the matrix had an extreme norm
    for (int c = warp; c < n; c += nwarp) {
      const double* z = G + (size_t)c * ldg;
      double acc = 0.0;
      for (int r = lane; r < n; r += 32) {
        double y = 0.0;
        for (int k = 0; k < n; ++k) {
          const double a = (r >= k) ? A[(size_t)k * lda + r] : A[(size_t)r * lda + k];
          y = fma(a * inva, z[k], y);
        }
        acc = fma(z[r], y, acc);
      }
      acc = warp_sum(acc);
      if (lane == 0 && isfinite(acc)) lam[c] = acc * ascale;
    }
    __syncthreads();
  }
  // garbage in (NaN / Inf) must be reported, not sorted into a plausible-looking spectrum
  int bad = 0;
  for (int c = tid; c < n; c += nthr) bad |= isfinite(lam[c]) ? 0 : 1;
  bad = __syncthreads_or(bad);

  // ---- rank sort ascending (ties broken by index) and scatter to global ------------------------
  double* wout = w + prob * stride_w;
  double* Zout = Z + prob * strideZ;
  for (int c = warp; c < n; c += nwarp) {
    const double lc = lam[c];
    const long long kc = sort_key(lc);
    int rank = 0;
    for (int j = lane; j < n; j += 32) {
      const long long kj = sort_key(lam[j]);  // total order on the bit patterns: a permutation even with NaNs
      rank += (kj < kc || (kj == kc && j < c)) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    if (lane == 0) wout[rank] = lc;
    const double* g = G + (size_t)c * ldg;
    for (int r = lane; r < n; r += 32) Zout[(size_t)rank * ldz + r] = g[r];
  }
  if (tid == 0 && info) info[prob] = (converged && !bad) ? 0 : 1;
}

size_t jacobi_smem_bytes(int n) {
  const int ne = (n + 1) & ~1;
  return ((size_t)ne * n + ne) * sizeof(double);
}

int jacobi_syev_batched(int n, const double* A, int lda, long long strideA, double* w,
                        long long stride_w, double* Z, int ldz, long long strideZ, int batch,
                        const int* batch_map, int* info_dev, cudaStream_t st) {
  if (batch <= 0 || n <= 0) return GSCF_B200_OK;
  if (n > kJacobiMaxN) {
    set_last_error("jacobi_syev_batched: n > 160 does not fit in shared memory");
    return GSCF_B200_ERR_INVALID_ARG;
  }
  const size_t smem = jacobi_smem_bytes(n);
  const int npairs = (n + 1) / 2;
  int warps = npairs < 32 ? npairs : 32;
  if (warps < 1) warps = 1;
  const int threads = warps * 32;
#define GSCF_JACOBI_LAUNCH(NPL, PPW)                                                                \
  do {                                                                                           \
    static PerDeviceOnce once;                                                                   \
    int dev_;                                                                                    \
    if (once.need(dev_)) {                                                                       \
      GSCF_CUDA_TRY(cudaFuncSetAttribute(jacobi_onesided_kernel<NPL, PPW>,                        \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                         (int)jacobi_smem_bytes(32 * NPL > kJacobiMaxN          \
                                                                    ? kJacobiMaxN               \
                                                                    : 32 * NPL)));              \
      once.done(dev_);                                                                           \
    }                                                                                            \
    jacobi_onesided_kernel<NPL, PPW><<<batch, threads, smem, st>>>(A, lda, strideA, w, stride_w, Z, \
                                                              ldz, strideZ, n, batch_map,       \
                                                              info_dev);                        \
  } while (0)
  if (n <= 32) GSCF_JACOBI_LAUNCH(1, 1);
  else if (n <= 64) GSCF_JACOBI_LAUNCH(2, 1);
  else if (n <= 128) GSCF_JACOBI_LAUNCH(4, 2);
  else GSCF_JACOBI_LAUNCH(5, 3);
#undef GSCF_JACOBI_LAUNCH
  count_launch();
  GSCF_CHECK_LAUNCH();
  return GSCF_B200_OK;
}

}  // namespace gscfb
