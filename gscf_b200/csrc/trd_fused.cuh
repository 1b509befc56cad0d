// Fused Householder tridiagonalisation (included by tridiag.cu): ONE launch for the whole reduction,
// ONE cross-CTA exchange per column.
//
// A group of G CTAs owns one matrix, columns dealt out cyclically (column c -> CTA c mod G).  The
// owned columns live in shared memory when they fit (n = 1024 on 148 CTAs: 56 KiB each; n = 512 on
// a 16-CTA cluster: 128 KiB each), otherwise they are updated in place in global memory (L2).
//
// The classical unblocked step (LAPACK dsytd2) for column j is
//     y = A22 v_j,  w_j = tau_j (y - 1/2 tau_j (y^T v_j) v_j),  A22 -= v_j w_j^T + w_j v_j^T,
// followed by the Householder vector v_{j+1} of the updated column j+1.  Its global dependencies
// are (a) every entry of y, (b) the updated column j+1.  With full symmetric storage and column
// ownership y_c = A22(:,c)^T v_j is local to the owner of column c, so the only exchange is an
// all-gather of y (n - j - 1 numbers) plus the raw column j+1 from its owner.  Everything that
// follows (y^T v, w_j, the updated column, its norm, v_{j+1}, tau_{j+1}) is O(n) work and is
// recomputed redundantly - and bit-identically - by every CTA, so no second exchange is needed.
// The rank-2 update with (v_j, w_j) is then fused with the product for the NEXT column:
//     pass j+1:  for my columns c >= j+2:  A(:,c) -= v_j w_j(c) + w_j v_j(c);  y'_c = A(:,c)^T v_{j+1}
// i.e. one read+write sweep over the CTA's part of the trailing matrix per column, no deferred
// panel update, no V/W correction terms, no trailing GEMM.
//
// The exchange goes through L2: producers store y_c / the raw column into a double-buffered slot, one
// thread per CTA fences and bumps a monotonic arrival counter and spins on it, then every thread loads
// the slot once.  (Measured on B200, scripts/micro/exchange.cu: polling the data itself - "data as
// flag" - costs ~10k cycles per step with 148 CTAs hammering the same lines; the counter scheme ~3.5k.)
// The CTAs of a group must be co-resident: thread-block cluster launch (G <= 16, batches) or
// cooperative launch (G up to the SM count, single large matrices).  The wait is bounded (clock64
// timeout + global abort flag); an aborted reduction poisons d[0] with NaN so that the divide &
// conquer stage flags the matrix as failed.
#pragma once
#include <cooperative_groups.h>

namespace gscfb {
namespace {

constexpr int TF_MAXS = 48;  // owned columns per CTA (n <= 48 G)
constexpr int TF_CG = 8;     // columns per register group in the sweep (independent 16-byte loads in flight)

struct TfArgs {
  double* A; int lda; long long sA;
  double* X;             // [cap][2][xstride / 2] exchange slots (parity of the step): y at [c], raw column j+1 at
                         // [np + r], SYMH: row partials of CTA g at [(2 + g) np + r]
  long long xstride;     // doubles per matrix in X (>= 2 (2 + TF_SYMH_MAXG) np)
  unsigned* ctr;         // [cap] arrival counters, zero on entry
  double *d, *e, *tau;   // [cap][n]
  unsigned* abort_flag;  // one word, zero on entry
  const int* map; int mat0; int n;
  int dstride;           // elements between the d / e / tau arrays of consecutive matrices
  int jstop;             // 0: reduce everything.  > 0 (even): stop after reflector jstop-1, apply it, write the trailing
                         // matrix A(jstop:, jstop:) back - the caller continues on it with a cheaper configuration
  int cluster;           // launched as one thread-block cluster per matrix: the exchange uses the hardware cluster barrier
  int smax;              // stride of the per-warp partial sums: >= owned columns per CTA, multiple of 8
};

__device__ __forceinline__ double tf_ld(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void tf_st(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void tf_st2(double* p, double x, double y) {
  asm volatile("st.relaxed.gpu.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(x), "d"(y) : "memory");
}
// offset of entry r inside a step's chunked exchange slot (see the kernel)
// release fence of the exchange: acq_rel at gpu scope (a `__threadfence()` is the stronger fence.sc)
__device__ __forceinline__ void tf_fence_release() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
// arrive at the exchange counter: release fence + relaxed add, or (GSCF_TF_RED_RELEASE) one red.release.gpu
__device__ __forceinline__ void tf_arrive(unsigned* ctr) {
#ifdef GSCF_TF_RED_RELEASE
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
#else
  tf_fence_release();
  atomicAdd(ctr, 1u);
#endif
}
__device__ __forceinline__ unsigned tf_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool tf_is_sentinel(double v) { return __double_as_longlong(v) == -1LL; }

// sum over the block, valid in every thread; one __syncthreads (the caller alternates `red` buffers)
template <int T>
__device__ __forceinline__ double tf_block_sum(double v, double* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int w = 0; w < T / 32; ++w) s += red[w];
  return s;
}

constexpr int TF_SYMH_MAXG = 8;  // largest group of the symmetric-half variant (every CTA publishes a row-partial vector)
inline long long tf_x_doubles(int n) { return 2LL * (2 + TF_SYMH_MAXG) * ((n + 1) & ~1); }
inline int tf_smax(int n, int G) { return round_up(ceil_div(n, G), 8); }
inline size_t tf_smem_bytes(int n, int G, bool smem_a) {
  const int np = round_up(n, 2);
  const int S = ceil_div(n, G);
  // v (x2), w, [solo: y, raw], per-warp partials, reduction scratch, owned columns
  return ((size_t)(G == 1 ? 5 : 3) * np + 32 * tf_smax(n, G) + 32 + 32 + 8 + (smem_a ? (size_t)S * np : 0)) * sizeof(double);
}

// SOLO: one CTA owns the whole matrix (n <= ~150, batches of small problems): the exchange is two shared-memory
// vectors, no counter, no global round trip.
// REGA: the CTA's columns live in REGISTERS (n <= 1024 on >= n/8 CTAs: 8 columns x 4 rows per thread): the sweep was
// bound by shared-memory bandwidth (17 LDS/STS.128 per row pair), with the matrix in registers it only reads the
// three vectors.  Thread t owns rows 2 (t + T p), 2 (t + T p) + 1, p < RP = 1024 / (2 T), of every owned column.
constexpr int TF_RS = 8;  // owned columns per CTA in registers
constexpr int TF_RROWS = 1024;  // rows covered by the register variant: row pairs per thread = TF_RROWS / (2 T)
// SYMH (symmetric half, groups of <= TF_SYMH_MAXG CTAs working in L2 - batches of mid-size matrices, the C5 kernel): the
// sweep only touches the LOWER triangle of my columns (rows r >= c), i.e. half the L2 traffic of the full-column sweep
// (8 instead of 16 bytes per trailing-matrix element and column).  Element A(r, c), r > c, then serves both
//   y_c += A(r, c) v_r   (column part: reduced over the rows, published by the owner of column c as before)   and
//   y_r += A(r, c) v_c   (row part: accumulated in the registers of the thread that owns row r - no reduction - and
//                         published as one n-vector per CTA; every CTA adds the G vectors after the SAME exchange).
// The stale upper triangle is never read as data.
template <int KMAX, bool SMEM_A, int T, bool SOLO = false, bool REGA = false, bool SYMH = false, int MINB = 1>
__global__ void __launch_bounds__(T, MINB)
trd_fused_kernel(const TfArgs a) {
  static_assert(!SYMH || (!SMEM_A && !SOLO && !REGA), "SYMH is a variant of the global-memory sweep");
  const int G = gridDim.x, g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = T / 32;
  const long long q = a.map ? a.map[blockIdx.y + a.mat0] : (long long)blockIdx.y + a.mat0;
  const int n = a.n, lda = a.lda, np = (n + 1) & ~1;
  double* __restrict__ A = a.A + q * a.sA;
  double* X = a.X + q * a.xstride;
  const long long xhalf = a.xstride >> 1;
  unsigned* ctr = a.ctr + q;

  extern __shared__ __align__(16) double tf_smem[];
  double* sv0 = tf_smem;
  double* sv1 = sv0 + np;
  double* sw = sv1 + np;
  double* sy = sw + np;                // SOLO: y and the raw column
  double* sraw = sy + (SOLO ? np : 0);
  double* spart = sraw + (SOLO ? np : 0);  // [32][smax] per-warp partial y
  const int smax = a.smax;
  double* sred1 = spart + 32 * smax;
  double* sred2 = sred1 + 32;
  double* sbc = sred2 + 32;            // 0: y[first]  1: a'[first+1]
  double* sAc = sbc + 8;               // SMEM_A: owned columns, slot s at sAc + s*np

  const int S = g < n ? (n - g + G - 1) / G : 0;  // owned columns c = g + s G

  for (int r = tid; r < np; r += T) { sv0[r] = 0.0; sv1[r] = 0.0; sw[r] = 0.0; }
  if (tid < 8) sbc[tid] = 0.0;
  if (SMEM_A) {
    for (int s = 0; s < S; ++s) {
      const double* src = A + (size_t)(g + s * G) * lda;
      for (int r = tid; r < np; r += T) sAc[(size_t)s * np + r] = r < n ? src[r] : 0.0;
    }
  }
  constexpr int RP = REGA ? TF_RROWS / (2 * T) : 1;  // row pairs per thread (2 at T = 256, 1 at T = 512)
  double2 ra[REGA ? TF_RS : 1][RP];
  if (REGA) {
#pragma unroll
    for (int s = 0; s < TF_RS; ++s)
#pragma unroll
      for (int p = 0; p < RP; ++p) {
        const int c = g + s * G, r = 2 * (tid + T * p);
        double2 v = make_double2(0.0, 0.0);
        if (s < S && r < n) {
          v.x = A[(size_t)c * lda + r];
          if (r + 1 < n) v.y = A[(size_t)c * lda + r + 1];
        }
        ra[REGA ? s : 0][REGA ? p : 0] = v;
      }
  }
  __syncthreads();

  double* svc = sv0;  // v_j      (zero above row j+1, one at row j+1)
  double* svn = sv1;  // v_{j+1}
  double tau_j = 0.0;

#ifdef GSCF_TF_DEBUG_TIMING
  long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tbucket[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tfence = 0;
  long long tmark = clock64();
#define TF_TICK(k) do { const long long _t = clock64(); tacc[k] += _t - tmark; tmark = _t; } while (0)
#else
#define TF_TICK(k) do { } while (0)
#endif
  const int jlast = (a.jstop > 0 && a.jstop <= n - 2) ? a.jstop - 1 : n - 2;
  for (int j = -1; j <= jlast; ++j) {
    const int first = j + 1;
    double yv[KMAX], rv[KMAX];
    TF_TICK(7);
    if (j < 0) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int r = tid + k * T;
        yv[k] = 0.0;
        rv[k] = r < n ? A[r] : 0.0;  // column 0 as it came in
      }
    } else {
      // ================= sweep: apply reflector j-1 to my columns, multiply them with v_j ==================
      const int s0 = first > g ? (first - g + G - 1) / G : 0;
      const int r_lo = first & ~1;
      const int npairs = (np - r_lo) >> 1;
      double* Xy = X + (size_t)(j & 1) * xhalf;
      double* Xr = Xy + np;
      const bool pub_raw = (g + s0 * G == first);  // I own column j+1: its updated values are the raw column
      if (SYMH) {
        constexpr int KP = (KMAX + 1) / 2;  // row pairs per thread
        double2 rowacc[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) rowacc[k] = make_double2(0.0, 0.0);
        for (int sg = s0; sg < S; sg += TF_CG) {
          double acc[TF_CG], wc[TF_CG], vc[TF_CG], vjc[TF_CG];
          int cu[TF_CG];
#pragma unroll
          for (int u = 0; u < TF_CG; ++u) {
            const int slot = sg + u;
            const int c = g + slot * G;
            acc[u] = 0.0;
            cu[u] = slot < S ? c : 0x3fffffff;     // beyond my columns: every row is "above the diagonal"
            wc[u] = slot < S ? sw[c] : 0.0;    // w_{j-1}(c)
            vc[u] = slot < S ? svn[c] : 0.0;   // v_{j-1}(c)
            vjc[u] = slot < S ? svc[c] : 0.0;  // v_j(c)
          }
          // rows above the first column of the group are above the diagonal of all its columns
          const int p_start = max(0, (((g + sg * G) & ~1) - r_lo) >> 1);
#pragma unroll
          for (int k = 0; k < KP; ++k) {
            const int p = tid + k * T;
            if (p >= p_start && p < npairs) {
              const int r = r_lo + 2 * p;
              const double2 vp = *reinterpret_cast<const double2*>(svn + r);  // v_{j-1}
              const double2 wp = *reinterpret_cast<const double2*>(sw + r);   // w_{j-1}
              const double2 vn = *reinterpret_cast<const double2*>(svc + r);  // v_j
              double2 av[TF_CG];
#pragma unroll
              for (int u = 0; u < TF_CG; ++u)
                if (sg + u < S) av[u] = *reinterpret_cast<const double2*>(A + (size_t)(g + (sg + u) * G) * lda + r);
              const bool odd_tail = (r + 1 >= n);
#pragma unroll
              for (int u = 0; u < TF_CG; ++u) {
                if (sg + u < S && r + 1 >= cu[u]) {  // at least one of the two rows is on or below the diagonal
                  double* col = A + (size_t)(g + (sg + u) * G) * lda;
                  double2 x = av[u];
                  x.x = fma(-vp.x, wc[u], fma(-wp.x, vc[u], x.x));
                  x.y = fma(-vp.y, wc[u], fma(-wp.y, vc[u], x.y));
                  if (odd_tail) {
                    x.y = 0.0;
                    col[r] = x.x;
                  } else {
                    *reinterpret_cast<double2*>(col + r) = x;
                  }
                  if (pub_raw && sg + u == s0) *reinterpret_cast<double2*>(Xr + r) = x;
                  const bool d0 = r >= cu[u];                      // (r, c) on or below the diagonal; (r + 1, c) always is
                  acc[u] = fma(d0 ? x.x : 0.0, vn.x, fma(x.y, vn.y, acc[u]));
                  rowacc[k].x = fma(r > cu[u] ? x.x : 0.0, vjc[u], rowacc[k].x);
                  rowacc[k].y = fma(r + 1 > cu[u] ? x.y : 0.0, vjc[u], rowacc[k].y);
                }
              }
            }
          }
          {
            static_assert(TF_CG == 8, "butterfly below is written for 8 accumulators");
            const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
            double a4[4], a2[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const double send = h16 ? acc[i] : acc[i + 4], keep = h16 ? acc[i + 4] : acc[i];
              a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const double send = h8 ? a4[i] : a4[i + 2], keep = h8 ? a4[i + 2] : a4[i];
              a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            double a1 = (h4 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, h4 ? a2[0] : a2[1], 4);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
            const int u = (h16 ? 4 : 0) + (h8 ? 2 : 0) + (h4 ? 1 : 0);
            if ((lane & 3) == 0 && sg + u < S) spart[warp * smax + sg + u] = a1;
          }
        }
        // my row partials (complete over my columns, no reduction needed): one n-vector per CTA
        double* Xrow = Xy + (size_t)(2 + g) * np;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          const int p = tid + k * T;
          if (p < npairs) tf_st2(Xrow + r_lo + 2 * p, rowacc[k].x, rowacc[k].y);
        }
      } else if (REGA) {
        static_assert(!REGA || TF_RS == TF_CG, "one register group");
        double acc[TF_CG], wc[TF_CG], vc[TF_CG];
#pragma unroll
        for (int u = 0; u < TF_CG; ++u) {
          const int c = g + u * G;
          acc[u] = 0.0;
          wc[u] = u < S ? sw[c] : 0.0;
          vc[u] = u < S ? svn[c] : 0.0;
        }
#pragma unroll
        for (int p = 0; p < RP; ++p) {
          const int r = 2 * (tid + T * p);
          if (r + 1 >= first && r < np) {
            const double2 vp = *reinterpret_cast<const double2*>(svn + r);  // v_{j-1}
            const double2 wp = *reinterpret_cast<const double2*>(sw + r);   // w_{j-1}
            const double2 vn = *reinterpret_cast<const double2*>(svc + r);  // v_j
#pragma unroll
            for (int u = 0; u < TF_CG; ++u) {
              // no test per column: a finished column (u < s0) only sees zeros of v_{j-1} / w_{j-1} (or one last update it
              // no longer needs) and its sum is never read, a column beyond S is zero and stays zero - a branch per column
              // kept the compiler from interleaving the eight FMA chains
              double2 x = ra[REGA ? u : 0][REGA ? p : 0];
              x.x = fma(-vp.x, wc[u], fma(-wp.x, vc[u], x.x));
              x.y = fma(-vp.y, wc[u], fma(-wp.y, vc[u], x.y));
              ra[REGA ? u : 0][REGA ? p : 0] = x;
              if (pub_raw && u == s0) *reinterpret_cast<double2*>(Xr + r) = x;
              acc[u] = fma(x.x, vn.x, fma(x.y, vn.y, acc[u]));
            }
          }
        }
        {
          const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
          double a4[4], a2[2];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double send = h16 ? acc[i] : acc[i + 4], keep = h16 ? acc[i + 4] : acc[i];
            a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const double send = h8 ? a4[i] : a4[i + 2], keep = h8 ? a4[i + 2] : a4[i];
            a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
          double a1 = (h4 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, h4 ? a2[0] : a2[1], 4);
          a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
          a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
          const int u = (h16 ? 4 : 0) + (h8 ? 2 : 0) + (h4 ? 1 : 0);
          if ((lane & 3) == 0 && u >= s0 && u < S) spart[warp * smax + u] = a1;
        }
      } else if (SOLO) {
        // few rows (n <= ~150): one WARP per group of 8 columns, lanes over the row pairs - the block-wide mapping
        // below would leave three quarters of the threads without a row
        for (int sg = s0 + warp * TF_CG; sg < S; sg += NW * TF_CG) {
          double acc[TF_CG], wc[TF_CG], vc[TF_CG];
  #pragma unroll
          for (int u = 0; u < TF_CG; ++u) {
            const int slot = sg + u;
            const int c = g + slot * G;
            acc[u] = 0.0;
            wc[u] = slot < S ? sw[c] : 0.0;   // w_{j-1}(c)
            vc[u] = slot < S ? svn[c] : 0.0;  // v_{j-1}(c): after the swap svn holds the previous reflector
          }
          for (int p = lane; p < npairs; p += 32) {
            const int r = r_lo + 2 * p;
            const double2 vp = *reinterpret_cast<const double2*>(svn + r);  // v_{j-1}
            const double2 wp = *reinterpret_cast<const double2*>(sw + r);   // w_{j-1}
            const double2 vn = *reinterpret_cast<const double2*>(svc + r);  // v_j
            double2 av[TF_CG];
  #pragma unroll
            for (int u = 0; u < TF_CG; ++u) {
              const int slot = sg + u;
              if (slot < S) {
                const double* col = SMEM_A ? sAc + (size_t)slot * np : A + (size_t)(g + slot * G) * lda;
                av[u] = *reinterpret_cast<const double2*>(col + r);
              }
            }
            const bool odd_tail = !SMEM_A && (r + 1 >= n);  // pad row of an odd-order matrix in global memory
  #pragma unroll
            for (int u = 0; u < TF_CG; ++u) {
              const int slot = sg + u;
              if (slot < S) {
                double* col = SMEM_A ? sAc + (size_t)slot * np : A + (size_t)(g + slot * G) * lda;
                double2 x = av[u];
                x.x = fma(-vp.x, wc[u], fma(-wp.x, vc[u], x.x));
                x.y = fma(-vp.y, wc[u], fma(-wp.y, vc[u], x.y));
                if (odd_tail) {
                  x.y = 0.0;
                  col[r] = x.x;
                } else {
                  *reinterpret_cast<double2*>(col + r) = x;
                }
                if (pub_raw && slot == s0) *reinterpret_cast<double2*>((SOLO ? sraw : Xr) + r) = x;
                acc[u] = fma(x.x, vn.x, fma(x.y, vn.y, acc[u]));
              }
            }
          }
          {
            // 8 sums over the warp with 7 exchanges instead of 40: every butterfly step halves the number of values a
            // lane is responsible for (SHFL issues at one warp per clock per SM - 80 of them per warp dominated the sweep)
            static_assert(TF_CG == 8, "butterfly below is written for 8 accumulators");
            const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
            double a4[4], a2[2];
  #pragma unroll
            for (int i = 0; i < 4; ++i) {
              const double send = h16 ? acc[i] : acc[i + 4], keep = h16 ? acc[i + 4] : acc[i];
              a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
  #pragma unroll
            for (int i = 0; i < 2; ++i) {
              const double send = h8 ? a4[i] : a4[i + 2], keep = h8 ? a4[i + 2] : a4[i];
              a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            double a1 = (h4 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, h4 ? a2[0] : a2[1], 4);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
            const int u = (h16 ? 4 : 0) + (h8 ? 2 : 0) + (h4 ? 1 : 0);
            if ((lane & 3) == 0 && sg + u < S) sy[sg + u] = a1;  // the warp owns these columns: complete sums
          }
        }
      } else {
        for (int sg = s0; sg < S; sg += TF_CG) {
          double acc[TF_CG], wc[TF_CG], vc[TF_CG];
  #pragma unroll
          for (int u = 0; u < TF_CG; ++u) {
            const int slot = sg + u;
            const int c = g + slot * G;
            acc[u] = 0.0;
            wc[u] = slot < S ? sw[c] : 0.0;   // w_{j-1}(c)
            vc[u] = slot < S ? svn[c] : 0.0;  // v_{j-1}(c): after the swap svn holds the previous reflector
          }
          for (int p = tid; p < npairs; p += T) {
            const int r = r_lo + 2 * p;
            const double2 vp = *reinterpret_cast<const double2*>(svn + r);  // v_{j-1}
            const double2 wp = *reinterpret_cast<const double2*>(sw + r);   // w_{j-1}
            const double2 vn = *reinterpret_cast<const double2*>(svc + r);  // v_j
            double2 av[TF_CG];
  #pragma unroll
            for (int u = 0; u < TF_CG; ++u) {
              const int slot = sg + u;
              if (slot < S) {
                const double* col = SMEM_A ? sAc + (size_t)slot * np : A + (size_t)(g + slot * G) * lda;
                av[u] = *reinterpret_cast<const double2*>(col + r);
              }
            }
            const bool odd_tail = !SMEM_A && (r + 1 >= n);  // pad row of an odd-order matrix in global memory
  #pragma unroll
            for (int u = 0; u < TF_CG; ++u) {
              const int slot = sg + u;
              if (slot < S) {
                double* col = SMEM_A ? sAc + (size_t)slot * np : A + (size_t)(g + slot * G) * lda;
                double2 x = av[u];
                x.x = fma(-vp.x, wc[u], fma(-wp.x, vc[u], x.x));
                x.y = fma(-vp.y, wc[u], fma(-wp.y, vc[u], x.y));
                if (odd_tail) {
                  x.y = 0.0;
                  col[r] = x.x;
                } else {
                  *reinterpret_cast<double2*>(col + r) = x;
                }
                if (pub_raw && slot == s0) *reinterpret_cast<double2*>((SOLO ? sraw : Xr) + r) = x;
                acc[u] = fma(x.x, vn.x, fma(x.y, vn.y, acc[u]));
              }
            }
          }
          {
            // 8 sums over the warp with 7 exchanges instead of 40: every butterfly step halves the number of values a
            // lane is responsible for (SHFL issues at one warp per clock per SM - 80 of them per warp dominated the sweep)
            static_assert(TF_CG == 8, "butterfly below is written for 8 accumulators");
            const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
            double a4[4], a2[2];
  #pragma unroll
            for (int i = 0; i < 4; ++i) {
              const double send = h16 ? acc[i] : acc[i + 4], keep = h16 ? acc[i + 4] : acc[i];
              a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
  #pragma unroll
            for (int i = 0; i < 2; ++i) {
              const double send = h8 ? a4[i] : a4[i + 2], keep = h8 ? a4[i + 2] : a4[i];
              a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            double a1 = (h4 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, h4 ? a2[0] : a2[1], 4);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
            const int u = (h16 ? 4 : 0) + (h8 ? 2 : 0) + (h4 ? 1 : 0);
            if ((lane & 3) == 0 && sg + u < S) spart[warp * smax + sg + u] = a1;
          }
        }
      }
#ifdef GSCF_TF_DEBUG_TIMING
      if (j >= 0) tbucket[(j * 16) / n] += clock64() - tmark;
#endif
      TF_TICK(0);
      __syncthreads();
      TF_TICK(1);
      if (!SOLO) {
        for (int slot = s0 + tid; slot < S; slot += T) {
          double y = 0.0;
#pragma unroll
          for (int w = 0; w < NW; ++w) y += spart[w * smax + slot];
          Xy[g + slot * G] = y;
        }
      }
      // ================= exchange: arrive, wait for the other CTAs, load y and the raw column =================
      TF_TICK(2);
      // the y stores above (threads 0 .. S-s0-1) and the raw stores of the sweep (ordered by the barrier after the sweep)
      // must precede the release fence of thread 0: a warp barrier is enough while the y writers share its warp
      if (S <= 32) __syncwarp();
      else __syncthreads();
      if (!SOLO && a.cluster) {
        // barrier.cluster.arrive.release / wait.acquire: the y and raw stores of every CTA of the cluster are visible
        // to all of them afterwards; no L2 atomic, no polling
        cooperative_groups::this_cluster().sync();
      } else if (!SOLO) {
        if (tid == 0) {
#ifdef GSCF_TF_DEBUG_TIMING
          const long long tf0 = clock64();
#endif
          tf_arrive(ctr);
#ifdef GSCF_TF_DEBUG_TIMING
          tfence += clock64() - tf0;
#endif
          const unsigned target = (unsigned)G * (unsigned)(j + 1);
          if (*reinterpret_cast<volatile unsigned*>(a.abort_flag) == 0u) {
            const long long t0 = clock64();
            int spins = 0;
            while (tf_ld_acquire(ctr) < target) {
              if ((++spins & 255) == 0 &&
                  (clock64() - t0 > 2000000000LL || *reinterpret_cast<volatile unsigned*>(a.abort_flag) != 0u)) {
                atomicExch(a.abort_flag, 1u);
                break;
              }
            }
          }
        }
        __syncthreads();
      }
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int r = tid + k * T;
        const bool need = r >= first && r < n;
        yv[k] = need ? (SOLO ? sy[r] : tf_ld(Xy + r)) : 0.0;
        rv[k] = need ? (SOLO ? sraw[r] : tf_ld(Xr + r)) : 0.0;
        if (SYMH && need) {  // + the row parts of every CTA of the group, in a fixed order (bit-identical in all CTAs)
          for (int g2 = 0; g2 < G; ++g2) yv[k] += tf_ld(Xy + (size_t)(2 + g2) * np + r);
        }
      }
    }
    TF_TICK(3);
    // ================= redundant O(n) part: w_j, updated column j+1, v_{j+1}, tau_{j+1} ========================
    double part = 0.0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int r = tid + k * T;
      if (r >= first && r < n) {
        part = fma(yv[k], svc[r], part);
        if (r == first) sbc[0] = yv[k];
      }
    }
    const double pdot = tf_block_sum<T>(part, sred1);  // y^T v_j
    TF_TICK(4);
    const double alpha_w = 0.5 * tau_j * pdot;
    const double w_first = tau_j * (sbc[0] - alpha_w);  // v_j[first] == 1
    const int nf = first + 1;
    double wv[KMAX], av2[KMAX];
    part = 0.0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int r = tid + k * T;
      wv[k] = 0.0; av2[k] = 0.0;
      if (r >= first && r < n) {
        const double vr = svc[r];
        wv[k] = tau_j * (yv[k] - alpha_w * vr);
        av2[k] = rv[k] - vr * w_first - wv[k];
        if (r == first) {
          if (g == 0) (a.d + q * a.dstride)[first] = av2[k];
        } else if (r == nf) {
          sbc[1] = av2[k];
        } else {
          part = fma(av2[k], av2[k], part);
        }
      }
    }
    const double xn2 = tf_block_sum<T>(part, sred2);
    TF_TICK(5);
    double beta = 0.0, tau_n = 0.0, scale = 0.0;
    if (nf < n) {
      householder_scalars(sbc[1], xn2, beta, tau_n, scale);
      if (g == 0 && tid == 0) { (a.e + q * a.dstride)[first] = beta; (a.tau + q * a.dstride)[first] = tau_n; }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int r = tid + k * T;
      if (r < n) {
        const double vnew = r < nf ? 0.0 : (r == nf ? 1.0 : av2[k] * scale);
        svn[r] = vnew;
        sw[r] = wv[k];
        // reflector j in LAPACK layout (below its unit entry).  Stored one step late: the exchange of step j
        // proves that every CTA has finished reading column j as it came in (column 0 at j = -1 needs this).
        if (j >= 0 && r > first && ((r >> 6) % G) == g) A[(size_t)j * lda + r] = svc[r];
      }
    }
    __syncthreads();
    TF_TICK(6);
    double* t = svc; svc = svn; svn = t;
    tau_j = tau_n;
  }
  if (jlast < n - 2) {
    // hand-over: apply the pending reflector jstop-1 (svn = v, sw = w after the last swap) to my columns >= jstop and
    // write them to global memory; A(jstop:, jstop:) is then the complete matrix of the remaining reduction
    const int js = a.jstop;
    const int s0 = js > g ? (js - g + G - 1) / G : 0;
    for (int slot = s0; slot < S; ++slot) {
      const int c = g + slot * G;
      const double wc = sw[c], vc = svn[c];
      const double* src = SMEM_A ? sAc + (size_t)slot * np : A + (size_t)c * lda;
      double* dst = A + (size_t)c * lda;
      if (!REGA)
        for (int r = js + tid; r < n; r += T) dst[r] = fma(-svn[r], wc, fma(-sw[r], vc, src[r]));
    }
    if (REGA) {
#pragma unroll
      for (int u = 0; u < TF_RS; ++u)
#pragma unroll
        for (int p = 0; p < RP; ++p) {
          const int c = g + u * G, r = 2 * (tid + T * p);
          if (u >= s0 && u < S) {
            const double2 x = ra[REGA ? u : 0][REGA ? p : 0];
            const double wc = sw[c], vc = svn[c];
            if (r >= js && r < n) A[(size_t)c * lda + r] = fma(-svn[r], wc, fma(-sw[r], vc, x.x));
            if (r + 1 >= js && r + 1 < n) A[(size_t)c * lda + r + 1] = fma(-svn[r + 1], wc, fma(-sw[r + 1], vc, x.y));
          }
        }
    }
  }
#ifdef GSCF_TF_DEBUG_TIMING
  if ((tid == 0 || tid == T - 1) && (g == 0 || g == G - 1) && blockIdx.y == 0)
    printf("cta %d tid %d cycles/col: sweep %lld sync %lld publish %lld poll %lld dot %lld norm %lld vec %lld loop %lld\n", g, tid,
           tacc[0] / n, tacc[1] / n, tacc[2] / n, tacc[3] / n, tacc[4] / n, tacc[5] / n, tacc[6] / n, tacc[7] / n);
  if (tid == 0 && g == 0 && blockIdx.y == 0) {
    printf("fence cycles/col %lld\n", tfence / n);
    printf("sweep cycles/col by sixteenth of the reduction:");
    for (int b = 0; b < 16; ++b) printf(" %lld", tbucket[b] * 16 / n);
    printf("\n");
  }
#endif
  if (!SOLO && g == 0 && tid == 0 && *reinterpret_cast<volatile unsigned*>(a.abort_flag) != 0u)
    (a.d + q * a.dstride)[0] = __longlong_as_double(0x7ff8000000000000LL);
}

// ---- single CTA, matrix in registers (n <= 128: the C4 ensemble) ---------------------------------------------------------
// The shared-memory SOLO variant above re-reads and re-writes the whole trailing matrix through shared memory for every
// column (n^2 16 bytes at 128 B / clock = 2 000 cycles at n = 128 - measured 3 400 of 5 100 cycles per column).  Here
// the matrix lives in registers: with NMAX = 128, thread (cg = tid / 64, pr = tid % 64) keeps rows 2 pr, 2 pr + 1 of
// columns cg, cg + 4, ... (NU = 32 columns, 64 registers); a sweep is 192 FMAs per thread on registers plus one 32-value
// butterfly over the warp.  The layout is fixed, so the cost of a column step does not shrink with the trailing matrix:
// NMAX = 64 (thread (cg = tid / 32, pr = tid % 32), columns cg + 8 u, NU = 8) is the same kernel for orders <= 64, and
// the driver hands the trailing 64 x 64 block of a larger matrix over to it (jstop, see TfArgs) - measured on B200,
// 4096 matrices of order 128: see DESIGN.md section 5.
// T = 512 (NMAX = 128 only: 64 row pairs x 8 column groups, NU = 16): the 256-thread layout issues ~920 instructions per
// warp and column with two warps per scheduler and stalls on fixed-latency dependencies (ncu: 35 % issue utilisation,
// `profiles/r02_ncu_full_solo_reg.json`); twice the warps at half the registers hide them.
template <int NMAX, int T = 256>
__global__ void __launch_bounds__(T, NMAX == 64 ? 2 : 1)  // the 64-row layout runs two CTAs per SM (<= 128 registers)
trd_solo_reg_kernel(const TfArgs a) {
  static_assert(NMAX == 128 || NMAX == 64, "layouts: 64 row pairs x 4 column groups, or 32 row pairs x 8 column groups");
  static_assert(T == 256 || (T == 512 && NMAX == 128), "thread counts");
  constexpr int NWARP = T / 32;
  constexpr int RPAIRS = NMAX / 2;           // row pairs = threads per column group
  constexpr int NCG = T / RPAIRS;            // column groups (4 / 8)
  constexpr int NU = NMAX / NCG;             // columns per thread (32 / 16 / 8)
  constexpr int WPG = RPAIRS / 32;           // warps per column group (2 / 1)
  __shared__ __align__(16) double sv0[NMAX], sv1[NMAX], sw[NMAX], sraw[NMAX];
  // {w_{j-1}(c), v_{j-1}(c)} of the sweep, grouped by column group: entry (c % NCG) NU + c / NCG - a thread reads the pairs
  // of its NU columns as consecutive 16-byte (warp-uniform) loads instead of two strided 8-byte loads per column
  __shared__ __align__(16) double2 spair[NMAX];
  __shared__ double spart[NWARP][NU + 1];
  __shared__ double sred1[NWARP], sred2[NWARP], sbc[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cg = tid / RPAIRS, pr = tid % RPAIRS, r0 = 2 * pr;
  const long long q = a.map ? a.map[blockIdx.y + a.mat0] : (long long)blockIdx.y + a.mat0;
  const int n = a.n, lda = a.lda;
  double* __restrict__ A = a.A + q * a.sA;

  double2 ra[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const int c = cg + NCG * u;
    double2 v = make_double2(0.0, 0.0);
    if (c < n && r0 < n) {
      v.x = A[(size_t)c * lda + r0];
      if (r0 + 1 < n) v.y = A[(size_t)c * lda + r0 + 1];
    }
    ra[u] = v;
  }
  if (tid < NMAX) { sv0[tid] = 0.0; sv1[tid] = 0.0; sw[tid] = 0.0; sraw[tid] = 0.0; spair[tid] = make_double2(0.0, 0.0); }
  if (tid < 8) sbc[tid] = 0.0;
  __syncthreads();

  double* svc = sv0;
  double* svn = sv1;
  double tau_j = 0.0;
  auto block_sum8 = [&](double v, double* red) {
    v = warp_sum(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) s += red[w];
    return s;
  };
  const int jlast = (a.jstop > 0 && a.jstop <= n - 2) ? a.jstop - 1 : n - 2;
  for (int j = -1; j <= jlast; ++j) {
    const int first = j + 1;
    double yv = 0.0, rv = 0.0;  // row r = tid (threads >= n idle in the O(n) part)
    if (j < 0) {
      rv = tid < n ? A[tid] : 0.0;
    } else {
      // ---- sweep: rank-2 update with reflector j-1, product with v_j, all on registers ---------------------------
      const double2 vp = *reinterpret_cast<const double2*>(svn + r0);
      const double2 wp = *reinterpret_cast<const double2*>(sw + r0);
      const double2 vn = *reinterpret_cast<const double2*>(svc + r0);
      // Chunks of eight columns per thread, skipped as a whole (uniform test) once all their columns are finished; inside
      // a chunk no test per column: a finished column only sees the zeros of v_{j-1} / w_{j-1} (column `first - 1` one last
      // update it no longer needs) and its sum is never read, columns >= n are zero and stay zero.  (ncu of the per-column
      // test, `profiles/r02_ncu_full_solo_reg.json`: a quarter of the executed instructions were BSSY / BSYNC / BRA / ISETP,
      // and every column was its own basic block - no interleaving of the FMA chains, 2.3 warps per issue waiting.)
      // The warp-wide sums are formed chunk by chunk (8 values -> 1 per lane with three halving steps over lane bits
      // 16 / 8 / 4: every step halves the values a lane is responsible for), then the chunks are folded over lane bits
      // 2 / 1: only eight accumulators are live at a time (the 255-register build of the all-at-once butterfly had no
      // room to hoist the shared-memory loads of w(c) / v(c) away from their FMAs), a finished chunk costs nothing.
      constexpr int NCH = NU / 8;
      const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2, h1 = lane & 1;
      double bch[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        bch[ch] = 0.0;
        if (NCG * (8 * ch + 8) > first) {  // the chunk's last column, NCG (8 ch + 7) + NCG - 1, is still alive
          double acc[8];
          double2 wvc[8];
#pragma unroll
          for (int u8 = 0; u8 < 8; ++u8) wvc[u8] = spair[cg * NU + 8 * ch + u8];
#pragma unroll
          for (int u8 = 0; u8 < 8; ++u8) {
            const int u = 8 * ch + u8;
            const int c = cg + NCG * u;
            const double wc = wvc[u8].x, vc = wvc[u8].y;
            double2 x = ra[u];
            x.x = fma(-vp.x, wc, fma(-wp.x, vc, x.x));
            x.y = fma(-vp.y, wc, fma(-wp.y, vc, x.y));
            ra[u] = x;
            if (c == first) *reinterpret_cast<double2*>(sraw + r0) = x;
            acc[u8] = fma(x.x, vn.x, x.y * vn.y);
          }
          double a4[4], a2[2];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const double send = h16 ? acc[i] : acc[i + 4], keep = h16 ? acc[i + 4] : acc[i];
            a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const double send = h8 ? a4[i] : a4[i + 2], keep = h8 ? a4[i + 2] : a4[i];
            a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
          bch[ch] = (h4 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, h4 ? a2[0] : a2[1], 4);
        }
      }
      // bch[ch]: column u = 8 ch + w of my column group, w = 4 h16 + 2 h8 + h4, summed over 8 of the warp's 32 row pairs
      const int w8 = (h16 ? 4 : 0) + (h8 ? 2 : 0) + (h4 ? 1 : 0);
      if constexpr (NCH == 4) {
        double c2[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) c2[i] = (h2 ? bch[i + 2] : bch[i]) + __shfl_xor_sync(0xffffffffu, h2 ? bch[i] : bch[i + 2], 2);
        const double tot = (h1 ? c2[1] : c2[0]) + __shfl_xor_sync(0xffffffffu, h1 ? c2[0] : c2[1], 1);
        spart[warp][8 * ((h2 ? 2 : 0) + (h1 ? 1 : 0)) + w8] = tot;
      } else if constexpr (NCH == 2) {
        double tot = (h2 ? bch[1] : bch[0]) + __shfl_xor_sync(0xffffffffu, h2 ? bch[0] : bch[1], 2);
        tot += __shfl_xor_sync(0xffffffffu, tot, 1);
        if (!h1) spart[warp][8 * (h2 ? 1 : 0) + w8] = tot;
      } else {
        static_assert(NCH == 1 || NCH == 2 || NCH == 4, "8, 16 or 32 columns per thread");
        double tot = bch[0] + __shfl_xor_sync(0xffffffffu, bch[0], 2);
        tot += __shfl_xor_sync(0xffffffffu, tot, 1);
        if ((lane & 3) == 0) spart[warp][w8] = tot;
      }
      __syncthreads();
      if (tid < n) {
        const int c = tid, g4 = c % NCG, u = c / NCG;
        double y = 0.0;
        if (c >= first) {
#pragma unroll
          for (int w = 0; w < WPG; ++w) y += spart[WPG * g4 + w][u];
        }
        yv = y;
        rv = c >= first ? sraw[c] : 0.0;
      }
    }
    // ---- O(n) part (identical to trd_fused_kernel with one row per thread) ---------------------------------------
    const int r = tid;
    double part = 0.0;
    if (r >= first && r < n) {
      part = yv * svc[r];
      if (r == first) sbc[0] = yv;
    }
    const double pdot = block_sum8(part, sred1);
    const double alpha_w = 0.5 * tau_j * pdot;
    const double w_first = tau_j * (sbc[0] - alpha_w);
    const int nf = first + 1;
    double wv = 0.0, av2 = 0.0;
    part = 0.0;
    if (r >= first && r < n) {
      const double vr = svc[r];
      wv = tau_j * (yv - alpha_w * vr);
      av2 = rv - vr * w_first - wv;
      if (r == first) (a.d + q * a.dstride)[first] = av2;
      else if (r == nf) sbc[1] = av2;
      else part = av2 * av2;
    }
    const double xn2 = block_sum8(part, sred2);
    double beta = 0.0, tau_n = 0.0, scale = 0.0;
    if (nf < n) {
      householder_scalars(sbc[1], xn2, beta, tau_n, scale);
      if (tid == 0) { (a.e + q * a.dstride)[first] = beta; (a.tau + q * a.dstride)[first] = tau_n; }
    }
    if (r < n) {
      const double vnew = r < nf ? 0.0 : (r == nf ? 1.0 : av2 * scale);
      const double vcur = svc[r];
      svn[r] = vnew;
      sw[r] = wv;
      spair[(r % NCG) * NU + r / NCG] = make_double2(wv, vcur);  // (w_j, v_j): the reflector the next sweep applies
      if (j >= 0 && r > first) A[(size_t)j * lda + r] = vcur;  // reflector j, LAPACK layout
    }
    __syncthreads();
    double* t = svc; svc = svn; svn = t;
    tau_j = tau_n;
  }
  if (jlast < n - 2) {
    // hand-over: apply the pending reflector jstop-1 (svn = v, sw = w after the last swap) to the trailing block and write
    // it back; A(jstop:, jstop:) is then the complete (full symmetric) matrix of the remaining reduction
    const int js = a.jstop;
    const double2 vp = *reinterpret_cast<const double2*>(svn + r0);
    const double2 wp = *reinterpret_cast<const double2*>(sw + r0);
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int c = cg + NCG * u;
      if (c >= js && c < n) {
        const double wc = sw[c], vc = svn[c];
        const double2 x = ra[u];
        if (r0 >= js && r0 < n) A[(size_t)c * lda + r0] = fma(-vp.x, wc, fma(-wp.x, vc, x.x));
        if (r0 + 1 >= js && r0 + 1 < n) A[(size_t)c * lda + r0 + 1] = fma(-vp.y, wc, fma(-wp.y, vc, x.y));
      }
    }
  }
}

}  // namespace
}  // namespace gscfb
