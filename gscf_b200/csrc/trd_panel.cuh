// Persistent panel kernel of the Householder tridiagonalisation (included by tridiag.cu).
//
// One cooperative launch reduces a whole panel of NB columns.  A group of G CTAs works on one
// matrix; the three global dependencies of every column (norm of the updated column -> v,
// trailing product y = A22 v, dot w^T v) are group-wide barriers on a monotonic counter in global
// memory instead of kernel boundaries, which removes ~3 launch latencies (and their cold starts)
// per column.  Cooperative launch guarantees that all CTAs of the grid are co-resident, so the
// spin barrier cannot dead-lock; distinct matrices of a batch use distinct counters.
//
// Work split inside a group:
//   rows   : CTA c owns rows [c*RPB, (c+1)*RPB) for the whole panel (P1 / P3 are row-parallel)
//   symv   : 128 x 32 tiles of the trailing matrix, round-robin over the CTAs; every thread issues
//            its 16 independent loads up front (memory-level parallelism, L2-resident for n <= ~2k)
#pragma once

namespace gscfb {
namespace {

constexpr int PT = 256;   // threads per CTA
constexpr int SR = 128;   // symv tile rows
constexpr int SC = 32;    // symv tile columns
constexpr int GMAX = 1024;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void group_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    while (ld_acquire_u32(ctr) < target) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
}

// sum of `count` doubles written by other CTAs (L2 loads), result broadcast through shared memory
__device__ __forceinline__ double group_sum(const double* p, int count, double* s_slot) {
  if (threadIdx.x < 32) {
    double s = 0.0;
    for (int b = threadIdx.x; b < count; b += 32) s += __ldcg(p + b);
    s = warp_sum(s);
    if (threadIdx.x == 0) *s_slot = s;
  }
  __syncthreads();
  return *s_slot;
}

__global__ void __launch_bounds__(PT)
trd_panel_kernel(double* __restrict__ Ab, int lda, long long sA, TrdWs ws, int j0, int pw, int last_panel,
                 unsigned bar_base, const int* __restrict__ map, int mat0) {
  __shared__ double sVi[NB], sWi[NB], sp1[NB], sp2[NB];
  __shared__ double sv[SC];
  __shared__ double sy[SR];
  __shared__ double red[33];
  __shared__ double s_tmp, s_beta, s_tau, s_scale;
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const long long q = map ? map[mat0 + blockIdx.y] : mat0 + blockIdx.y;
  const int n = ws.n;
  double* A = Ab + q * sA;
  double* W = ws.W + q * (long long)ws.ld * NB;
  double* pnorm = ws.pnorm + q * ws.npart;
  double* pdot = ws.pdot + q * ws.npart;
  double* p12 = ws.p12 + q * ws.nrb * 2 * NB;
  double* ypart = ws.ypart + q * (long long)ws.ncb * n;
  double* scal = ws.scal + q * 8;
  double* dv = ws.d + q * n;
  double* ev = ws.e + q * n;
  double* tauv = ws.tau + q * n;
  unsigned* bar = ws.bar + q;
  unsigned target = bar_base;
  const int RPB = (n + G - 1) / G;            // rows owned per CTA (<= PT by construction)
  const int rown = (tid < RPB) ? cta * RPB + tid : n;  // my row (n = none)

  const int ncols = pw + (last_panel ? 1 : 0);
  for (int ci = 0; ci < ncols; ++ci) {
    const int i = j0 + ci, li = ci;
    const bool tail_only = (ci == pw);  // last panel: only the diagonal entry of column n-1
    // ================= P1: finalise w_{i-1}, update column i, partial norm =====================
    double alpha_prev = 0.0;
    if (li > 0) alpha_prev = -0.5 * __ldcg(tauv + i - 1) * group_sum(pdot, G, &s_tmp);
    if (tid < li) {
      sVi[tid] = __ldcg(A + (size_t)(j0 + tid) * lda + i);
      sWi[tid] = (tid == li - 1) ? __ldcg(scal + 1) + alpha_prev : __ldcg(W + (size_t)tid * ws.ld + i);
    }
    __syncthreads();
    double part = 0.0;
    if (rown >= i && rown < n) {
      const int r = rown;
      double a = A[(size_t)i * lda + r];
      if (li > 0) {
        const double vprev = A[(size_t)(i - 1) * lda + r];
        const double wfin = W[(size_t)(li - 1) * ws.ld + r] + alpha_prev * vprev;
        W[(size_t)(li - 1) * ws.ld + r] = wfin;
        double acc = 0.0;
        for (int j = 0; j < li - 1; ++j)
          acc += A[(size_t)(j0 + j) * lda + r] * sWi[j] + W[(size_t)j * ws.ld + r] * sVi[j];
        a -= acc;
        a -= vprev * sWi[li - 1] + wfin * sVi[li - 1];
      }
      A[(size_t)i * lda + r] = a;
      if (r == i) dv[i] = a;
      if (r == i + 1) scal[0] = a;
      if (r >= i + 2) part = a * a;
    }
    if (tail_only) break;
    part = block_sum(part, red);
    if (tid == 0) pnorm[cta] = part;
    target += G;
    group_barrier(bar, target);
    // ================= P2: y = A22 v (tiles) and V^T v, W^T v (row chunks) =========================
    {
      const double x2 = group_sum(pnorm, G, &s_tmp);
      if (tid == 0) {
        double beta, tau, scale;
        householder_scalars(__ldcg(scal), x2, beta, tau, scale);
        s_beta = beta; s_tau = tau; s_scale = scale;
      }
      __syncthreads();
    }
    const double scale = s_scale, tau = s_tau;
    const int first = i + 1, m = n - first;
    const int ntr = (m + SR - 1) / SR, ntc = (m + SC - 1) / SC;
    const int ntiles = ntr * ntc;
    for (int wk = cta; wk < ntiles + ntr; wk += G) {
      if (wk < ntiles) {
        const int tr = wk % ntr, tc = wk / ntr;
        const int cbase = first + tc * SC;
        __syncthreads();  // sv / sy reuse
        if (tid < SC) {
          const int c = cbase + tid;
          sv[tid] = (c < n) ? ((c == first) ? 1.0 : __ldcg(A + (size_t)i * lda + c) * scale) : 0.0;
        }
        __syncthreads();
        const int r = first + tr * SR + (tid & (SR - 1));
        const int half = tid >> 7;  // 0 / 1: columns [0,16) / [16,32) of the tile
        const int c0 = cbase + half * (SC / 2);
        double av[SC / 2];
#pragma unroll
        for (int k = 0; k < SC / 2; ++k)
          av[k] = (r < n && c0 + k < n) ? A[(size_t)(c0 + k) * lda + r] : 0.0;
        double y0 = 0.0, y1 = 0.0;
#pragma unroll
        for (int k = 0; k < SC / 2; k += 2) {
          y0 = fma(av[k], sv[half * (SC / 2) + k], y0);
          y1 = fma(av[k + 1], sv[half * (SC / 2) + k + 1], y1);
        }
        const double y = y0 + y1;
        if (half == 1) sy[tid & (SR - 1)] = y;
        __syncthreads();
        if (half == 0 && r < n) ypart[(size_t)tc * n + r] = y + sy[tid];
      } else {
        // V^T v and W^T v over row chunk rc: warp w takes columns j = w, w + 8, ...
        const int rc = wk - ntiles;
        double vr[SR / 32];
#pragma unroll
        for (int t = 0; t < SR / 32; ++t) {
          const int r = first + rc * SR + lane + 32 * t;
          vr[t] = (r < n) ? ((r == first) ? 1.0 : __ldcg(A + (size_t)i * lda + r) * scale) : 0.0;
        }
        for (int j = warp; j < li; j += PT / 32) {
          double s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int t = 0; t < SR / 32; ++t) {
            const int r = first + rc * SR + lane + 32 * t;
            if (r < n) {
              s1 = fma(__ldcg(A + (size_t)(j0 + j) * lda + r), vr[t], s1);
              s2 = fma(__ldcg(W + (size_t)j * ws.ld + r), vr[t], s2);
            }
          }
          s1 = warp_sum(s1);
          s2 = warp_sum(s2);
          if (lane == 0) {
            p12[(size_t)rc * 2 * NB + j] = s1;
            p12[(size_t)rc * 2 * NB + NB + j] = s2;
          }
        }
      }
    }
    target += G;
    group_barrier(bar, target);
    // ================= P3: w = tau (y - V p2 - W p1), partial w^T v ==================================
    if (tid < 2 * li) {
      const int j = tid % li, which = tid / li;
      double s = 0.0;
      for (int b = 0; b < ntr; ++b) s += __ldcg(p12 + (size_t)b * 2 * NB + which * NB + j);
      if (which == 0) sp1[j] = s; else sp2[j] = s;
    }
    __syncthreads();
    part = 0.0;
    if (rown >= first && rown < n) {
      const int r = rown;
      double y = 0.0;
      for (int c = 0; c < ntc; ++c) y += __ldcg(ypart + (size_t)c * n + r);
      const double v = (r == first) ? 1.0 : A[(size_t)i * lda + r] * scale;
      double acc = 0.0;
      for (int j = 0; j < li; ++j)
        acc += A[(size_t)(j0 + j) * lda + r] * sp2[j] + W[(size_t)j * ws.ld + r] * sp1[j];
      const double w = tau * (y - acc);
      W[(size_t)li * ws.ld + r] = w;
      A[(size_t)i * lda + r] = v;
      if (r == first) scal[1] = w;
      part = w * v;
    }
    part = block_sum(part, red);
    if (tid == 0) {
      pdot[cta] = part;
      if (cta == 0) { ev[i] = s_beta; tauv[i] = tau; }
    }
    target += G;
    group_barrier(bar, target);
  }
  // ================= P4: finalise the last w of the panel (needed by the trailing GEMM update) =========
  if (!last_panel) {
    const int i = j0 + pw - 1, li = pw - 1;
    const double alpha = -0.5 * __ldcg(tauv + i) * group_sum(pdot, G, &s_tmp);
    if (rown >= i + 1 && rown < n) W[(size_t)li * ws.ld + rown] += alpha * A[(size_t)i * lda + rown];
  }
}

}  // namespace
}  // namespace gscfb
