// Symmetric eigensolver above the Jacobi range: Householder tridiagonalisation, divide & conquer on the tridiagonal
// matrix (stedc.cu), blocked back-transformation.  Batched over matrices.
//
//   A = Q T Q^T   (Q = H(0) H(1) ... H(n-2), H(i) = I - tau_i v_i v_i^T, LAPACK dsytrd "lower")
//   T = Z_T diag(w) Z_T^T     (stedc.cu)
//   Z = Q Z_T                 (compact-WY block reflectors, DMMA GEMMs)
//
// Tridiagonalisation - the production path is the fused one-launch reduction of trd_fused.cuh / trd_fused_blk.cuh
// (tridiagonalise_fused below picks the variant: single CTA in registers or shared memory, cluster per matrix,
// cooperative group with the columns in registers / shared memory / L2, blocked for matrices larger than L2).
// Fall-back for what those do not cover (and A/B switch GSCF_B200_TRD_LEGACY=1): three launches per column at the
// global dependencies of dlatrd,
//     trd_col_update : a_i <- A(:,i) - V W(i,:)^T - W V(i,:)^T, partial |a_i|^2       (needs w_{i-1})
//     trd_symv       : y = A22 v_i in 128x128 tiles + partial V^T v, W^T v              (needs v_i)
//     trd_assemble_w : w_i = tau (y - V (W^T v) - W (V^T v)), partial w^T v             (needs y)
//   followed per panel by the deferred rank-2k DMMA GEMM update of the trailing matrix.
// (The round-1 persistent-panel and cluster-panel kernels were superseded by the fused kernels and removed.)
#include <atomic>

#include "eig.cuh"
#include "gemm.cuh"
#include "stedc.cuh"

namespace gscfb {

namespace {

constexpr int NB = 32;    // tridiagonalisation panel width
constexpr int NBQ = 128;  // back-transformation block width (64: 20.0 ms, 128: see DESIGN.md section 5, n = 4096)
constexpr int RB = 128;   // rows per CTA in the vector kernels
constexpr int TS = 128;   // symv tile edge

struct TrdWs {
  double *d, *e, *tau;  // [cap][n]
  double* W;            // [cap][ld*NB]
  double* ypart;        // [cap][ncb*n]
  double* pnorm;        // [cap][nrb]
  double* pdot;         // [cap][nrb]
  double* p12;          // [cap][nrb*2*NB]
  double* scal;         // [cap][8]   0: a[i+1]  1: raw w[i+1]
  unsigned* bar;        // [cap] group barrier counters of the persistent panel kernel
  double* VW;           // [cap][ld * 4*NB]  [V|W] and [W|V] panels for the fused rank-2k update
  unsigned long long* amax;  // [cap] max |a_ij| (bit pattern)
  double* ascale;       // [cap] 1, or the norm an extreme matrix was divided by
  double* SK;           // split-k partials: min(cap,16) * 16 * 2 NBQ * n
  double* X;            // [cap][4 np] exchange slots of the fused tridiagonalisation (trd_fused.cuh)
  double* VPW;          // [cap][2][32][np] pending reflector panels of the blocked fused kernel
  double* PQ;           // [cap][2][256][64] partial dots of the blocked fused kernel
  int npart;            // entries per matrix in pnorm / pdot
  double* Vf;           // [cap][ld*n]  all reflectors, unit lower trapezoidal, zeros above
  double* Wf;           // [cap][ld*n]  V_b T_b per block of NBQ reflectors
  double* Gall;         // [cap][nbq][NBQ*NBQ]  V_b^T V_b
  double* Tall;         // [cap][nbq][NBQ*NBQ]  T_b
  GemmDesc* bdesc;      // [cap][nbq] descriptors of the per-block set-up GEMMs
  int nbq;              // blocks of nbw reflectors
  int nbw;              // block width of the back-transformation: 64 for n <= 256 (batches: small T factors, several
                        // CTAs per SM in larft), NBQ = 128 above
  double* Yw;           // [cap][2 NBQ * n]  (a merged pair of blocks: 2 NBQ rows)
  int n, ld, nrb, ncb, cap;
};
inline int ws_cap_hint(const TrdWs& ws) { return ws.cap; }

__device__ __forceinline__ void householder_scalars(double alpha, double xnorm2, double& beta, double& tau,
                                                     double& scale) {
  // LAPACK dlarfg.  Instead of its safe-minimum rescaling loop: a sum of squares below safmin / eps carries no
  // relative accuracy (subnormal squares), and a reflector built from it is not orthogonal - such a column (entries
  // below 1e-146 in a matrix that tridiag_dc_syev_batched has scaled to a sane norm) is treated as exactly zero.
  if (xnorm2 < 1e-292) {
    beta = alpha; tau = 0.0; scale = 0.0;
  } else {
    beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
    tau = (beta - alpha) / beta;
    scale = 1.0 / (alpha - beta);
  }
}

// ---- K1 ------------------------------------------------------------------------------------------
// Column i (absolute), li = i - j0 previous reflectors of the current panel live in A(:, j0..i-1)
// (explicit unit entries) and W(:, 0..li-1) (column li-1 not yet finalised: w += alpha v).
__global__ void __launch_bounds__(RB)
trd_col_update(double* __restrict__ Ab, int lda, long long sA, TrdWs ws, int i, int j0, int nrb_prev,
               const int* __restrict__ map) {
  __shared__ double sVi[NB], sWi[NB];
  __shared__ double s_alpha;
  __shared__ double red[33];
  const long long q = map ? map[blockIdx.y] : blockIdx.y;
  const int n = ws.n, li = i - j0;
  double* A = Ab + q * sA;
  double* W = ws.W + q * (long long)ws.ld * NB;
  const double* pdot = ws.pdot + q * ws.npart;
  const double* tauv = ws.tau + q * n;
  double* scal = ws.scal + q * 8;
  if (threadIdx.x == 0) {
    double alpha_prev = 0.0;
    if (li > 0) {
      double s = 0.0;
      for (int b = 0; b < nrb_prev; ++b) s += pdot[b];
      alpha_prev = -0.5 * tauv[i - 1] * s;
    }
    s_alpha = alpha_prev;
  }
  __syncthreads();
  const double alpha_prev = s_alpha;
  if (threadIdx.x < li) {
    const int j = threadIdx.x;
    sVi[j] = A[(size_t)(j0 + j) * lda + i];
    // row i of W: the newest column is taken from the raw value saved by trd_assemble_w
    sWi[j] = (j == li - 1) ? scal[1] + alpha_prev * 1.0 : W[(size_t)j * ws.ld + i];
  }
  __syncthreads();
  const int r = i + blockIdx.x * RB + threadIdx.x;
  double part = 0.0;
  if (r < n) {
    double a = A[(size_t)i * lda + r];
    if (li > 0) {
      // finalise w_{i-1}[r] (rows r >= i)
      const double vprev = A[(size_t)(i - 1) * lda + r];
      const double wfin = W[(size_t)(li - 1) * ws.ld + r] + alpha_prev * vprev;
      W[(size_t)(li - 1) * ws.ld + r] = wfin;
      for (int j = 0; j < li - 1; ++j)
        a -= A[(size_t)(j0 + j) * lda + r] * sWi[j] + W[(size_t)j * ws.ld + r] * sVi[j];
      a -= vprev * sWi[li - 1] + wfin * sVi[li - 1];
    }
    A[(size_t)i * lda + r] = a;
    if (r == i) (ws.d + q * n)[i] = a;
    if (r == i + 1) scal[0] = a;
    if (r >= i + 2) part = a * a;
  }
  part = block_sum(part, red);
  if (threadIdx.x == 0) (ws.pnorm + q * ws.npart)[blockIdx.x] = part;
}

// ---- K2 ------------------------------------------------------------------------------------------
// grid.x = ncb_act * nrb_act tiles of the trailing matrix (rows/cols >= i+1) followed by nrb_act
// extra CTAs that reduce V^T v and W^T v over their row block.
__global__ void __launch_bounds__(TS)
trd_symv(const double* __restrict__ Ab, int lda, long long sA, TrdWs ws, int i, int j0, int nrb_k1,
         int ntr, int ntc, const int* __restrict__ map) {
  __shared__ double sv[TS];
  __shared__ double s_scale;
  const long long q = map ? map[blockIdx.y] : blockIdx.y;
  const int n = ws.n, li = i - j0;
  const double* A = Ab + q * sA;
  const double* W = ws.W + q * (long long)ws.ld * NB;
  const double* scal = ws.scal + q * 8;
  if (threadIdx.x == 0) {
    const double* pn = ws.pnorm + q * ws.npart;
    double x2 = 0.0;
    for (int b = 0; b < nrb_k1; ++b) x2 += pn[b];
    double beta, tau, scale;
    householder_scalars(scal[0], x2, beta, tau, scale);
    s_scale = scale;
  }
  __syncthreads();
  const double scale = s_scale;
  const int first = i + 1;
  const int tile = blockIdx.x;
  if (tile < ntr * ntc) {
    const int tr = tile % ntr, tc = tile / ntr;
    const int c0 = first + tc * TS;
    const int c = c0 + threadIdx.x;
    sv[threadIdx.x] = (c < n) ? ((c == first) ? 1.0 : A[(size_t)i * lda + c] * scale) : 0.0;
    __syncthreads();
    const int r = first + tr * TS + threadIdx.x;
    const int cmax = min(TS, n - c0);
    double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
    if (r < n) {
      const double* a = A + (size_t)c0 * lda + r;
      int cc = 0;
      for (; cc + 4 <= cmax; cc += 4) {
        y0 = fma(a[(size_t)(cc + 0) * lda], sv[cc + 0], y0);
        y1 = fma(a[(size_t)(cc + 1) * lda], sv[cc + 1], y1);
        y2 = fma(a[(size_t)(cc + 2) * lda], sv[cc + 2], y2);
        y3 = fma(a[(size_t)(cc + 3) * lda], sv[cc + 3], y3);
      }
      for (; cc < cmax; ++cc) y0 = fma(a[(size_t)cc * lda], sv[cc], y0);
      (ws.ypart + q * (long long)ws.ncb * n)[(size_t)tc * n + r] = (y0 + y1) + (y2 + y3);
    }
  } else {
    // V^T v and W^T v partials of row block rb (warp w handles columns j = w, w+4, ...)
    const int rb = tile - ntr * ntc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double vr[TS / 32];
#pragma unroll
    for (int t = 0; t < TS / 32; ++t) {
      const int r = first + rb * TS + lane + 32 * t;
      vr[t] = (r < n) ? ((r == first) ? 1.0 : A[(size_t)i * lda + r] * scale) : 0.0;
    }
    double* p12 = ws.p12 + (q * ws.nrb + rb) * 2 * NB;
    for (int j = warp; j < li; j += TS / 32) {
      double s1 = 0.0, s2 = 0.0;
#pragma unroll
      for (int t = 0; t < TS / 32; ++t) {
        const int r = first + rb * TS + lane + 32 * t;
        if (r < n) {
          s1 = fma(A[(size_t)(j0 + j) * lda + r], vr[t], s1);
          s2 = fma(W[(size_t)j * ws.ld + r], vr[t], s2);
        }
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) {
        p12[j] = s1;       // (V^T v)_j
        p12[NB + j] = s2;  // (W^T v)_j
      }
    }
  }
}

// ---- K3 ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RB)
trd_assemble_w(double* __restrict__ Ab, int lda, long long sA, TrdWs ws, int i, int j0, int nrb_k1, int ntr,
               int ntc, const int* __restrict__ map) {
  __shared__ double sp1[NB], sp2[NB];
  __shared__ double s_sc[3];
  __shared__ double red[33];
  const long long q = map ? map[blockIdx.y] : blockIdx.y;
  const int n = ws.n, li = i - j0;
  double* A = Ab + q * sA;
  double* W = ws.W + q * (long long)ws.ld * NB;
  double* scal = ws.scal + q * 8;
  if (threadIdx.x == 0) {
    const double* pn = ws.pnorm + q * ws.npart;
    double x2 = 0.0;
    for (int b = 0; b < nrb_k1; ++b) x2 += pn[b];
    double beta, tau, scale;
    householder_scalars(scal[0], x2, beta, tau, scale);
    s_sc[0] = beta; s_sc[1] = tau; s_sc[2] = scale;
    if (blockIdx.x == 0) {
      (ws.e + q * n)[i] = beta;
      (ws.tau + q * n)[i] = tau;
    }
  }
  if (threadIdx.x < 2 * li) {
    const int j = threadIdx.x % li, which = threadIdx.x / li;
    const double* p12 = ws.p12 + q * ws.nrb * 2 * NB;
    double s = 0.0;
    for (int b = 0; b < ntr; ++b) s += p12[(size_t)b * 2 * NB + which * NB + j];
    if (which == 0) sp1[j] = s; else sp2[j] = s;
  }
  __syncthreads();
  const double tau = s_sc[1], scale = s_sc[2];
  const int first = i + 1;
  const int r = first + blockIdx.x * RB + threadIdx.x;
  double part = 0.0;
  if (r < n) {
    const double* yp = ws.ypart + q * (long long)ws.ncb * n;
    double y = 0.0;
    for (int c = 0; c < ntc; ++c) y += yp[(size_t)c * n + r];
    const double v = (r == first) ? 1.0 : A[(size_t)i * lda + r] * scale;
    double acc = y;
    for (int j = 0; j < li; ++j)
      acc -= A[(size_t)(j0 + j) * lda + r] * sp2[j] + W[(size_t)j * ws.ld + r] * sp1[j];
    const double w = tau * acc;
    W[(size_t)li * ws.ld + r] = w;
    A[(size_t)i * lda + r] = v;
    if (r == first) scal[1] = w;  // raw w_i[i+1]: row i+1 of W is needed by the next column's update
    part = w * v;
  }
  part = block_sum(part, red);
  if (threadIdx.x == 0) (ws.pdot + q * ws.npart)[blockIdx.x] = part;
}

// w_last += alpha v_last (rows >= i+1) at the end of a panel
__global__ void __launch_bounds__(RB)
trd_finalize_w(const double* __restrict__ Ab, int lda, long long sA, TrdWs ws, int i, int j0, int nrb_prev,
               const int* __restrict__ map) {
  __shared__ double s_alpha;
  const long long q = map ? map[blockIdx.y] : blockIdx.y;
  const int n = ws.n, li = i - j0;  // i = last column of the panel
  const double* A = Ab + q * sA;
  double* W = ws.W + q * (long long)ws.ld * NB;
  if (threadIdx.x == 0) {
    const double* pdot = ws.pdot + q * ws.npart;
    double s = 0.0;
    for (int b = 0; b < nrb_prev; ++b) s += pdot[b];
    s_alpha = -0.5 * (ws.tau + q * n)[i] * s;
  }
  __syncthreads();
  const int r = i + 1 + blockIdx.x * RB + threadIdx.x;
  if (r < n) W[(size_t)li * ws.ld + r] += s_alpha * A[(size_t)i * lda + r];
}

// ---- scaling of extreme matrices (dsyev: anrm outside [rmin, rmax]) ---------------------------------------------------
// max |a_ij| per matrix as the bit pattern of a non-negative double (ordered like the integers); NaN sorts above inf
__global__ void __launch_bounds__(256)
absmax_kernel(const double* __restrict__ Ab, int lda, long long sA, int n, unsigned long long* __restrict__ amax,
              const int* __restrict__ map) {
  const long long q = map ? map[blockIdx.y] : blockIdx.y;
  const double* A = Ab + q * sA;
  unsigned long long m = 0ull;
  const long long total = (long long)lda * n;  // the pad rows are part of the buffer (zero or finite garbage is harmless:
  (void)total;                                  // only rows < n are read)
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)n * n;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx / n), r = (int)(idx - (long long)c * n);
    const unsigned long long b = (unsigned long long)__double_as_longlong(fabs(A[(size_t)c * lda + r]));
    m = b > m ? b : m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
    m = t > m ? t : m;
  }
  if ((threadIdx.x & 31) == 0 && m) atomicMax(amax + q, m);
}

// A <- A / anrm when anrm is outside [1e-100, 1e100]; ascale = the factor the eigenvalues are multiplied with at the end
__global__ void __launch_bounds__(256)
scale_extreme_kernel(double* __restrict__ Ab, int lda, long long sA, int n, const unsigned long long* __restrict__ amax,
                     double* __restrict__ ascale, const int* __restrict__ map) {
  const long long q = map ? map[blockIdx.y] : blockIdx.y;
  const double a = __longlong_as_double((long long)amax[q]);
  const bool ext = a > 0.0 && isfinite(a) && (a < 1e-100 || a > 1e100);
  if (blockIdx.x == 0 && threadIdx.x == 0) ascale[q] = ext ? a : 1.0;
  if (!ext) return;
  double* A = Ab + q * sA;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)n * n;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx / n), r = (int)(idx - (long long)c * n);
    A[(size_t)c * lda + r] /= a;
  }
}

// ---- back-transformation helpers ---------------------------------------------------------------------
// T (pw x pw upper triangular, the dlarft factor) from G = V^T V and tau through T^-1 = diag(1/tau) + striu(G):
// column c of T solves the triangular system by back substitution,
//   T(c,c) = tau_c,   T(r,c) = -tau_r sum_{k=r+1..c} G(r,k) T(k,c),   r = c-1 .. 0
// (tau_r = 0, i.e. H_r = I, gives a zero row and column, as dlarft does).  The columns are independent: one quad
// of threads per column, no block-wide barrier inside the recurrence.
__global__ void __launch_bounds__(4 * NBQ) larft_kernel(TrdWs ws, const int* __restrict__ map) {
  // both triangles packed in shared memory (2 x 66 KB at NBQ = 128):
  //   Xp: column c of T, entries k = 0..c, at c (c + 1) / 2          Gp: row r of G, entries k = r+1..pw-1, at gbase(r)
  extern __shared__ double sm[];
  double* Xp = sm;
  const int nbw = ws.nbw;
  double* Gp = sm + nbw * (nbw + 1) / 2;
  __shared__ double stau[NBQ];
  const long long q = map ? map[blockIdx.y] : blockIdx.y;
  const int b = blockIdx.x, j0 = b * nbw, pw = min(nbw, ws.n - 1 - j0);
  const double* __restrict__ Gg = ws.Gall + (q * ws.nbq + b) * NBQ * NBQ;  // column-major, ld = pw, symmetric
  auto gbase = [pw](int r) { return r * (2 * pw - r - 1) / 2 - (r + 1); };  // Gp[gbase(r) + k] = G(r, k), k > r
  for (int idx = threadIdx.x; idx < pw * pw; idx += blockDim.x) {
    const int k = idx / pw, r = idx - k * pw;  // element (r, k) of the column-major block
    if (k > r) Gp[gbase(r) + k] = Gg[idx];
  }
  for (int idx = threadIdx.x; idx < pw; idx += blockDim.x) stau[idx] = (ws.tau + q * ws.n + j0)[idx];
  __syncthreads();
  const int c = threadIdx.x >> 2, part = threadIdx.x & 3;
  double* x = Xp + c * (c + 1) / 2;
  const unsigned qmask = 0xFu << (threadIdx.x & 28);  // the four lanes of my quad: quads of a warp run different trip counts
  if (c < pw) {
    if (part == 0) x[c] = stau[c];
    __syncwarp(qmask);
    for (int r = c - 1; r >= 0; --r) {
      const double* g = Gp + gbase(r);
      double s = 0.0;
      for (int k = r + 1 + part; k <= c; k += 4) s = fma(g[k], x[k], s);
      s += __shfl_xor_sync(qmask, s, 1);
      s += __shfl_xor_sync(qmask, s, 2);
      if (part == 0) x[r] = -stau[r] * s;
      __syncwarp(qmask);
    }
  }
  __syncthreads();
  double* Tg = ws.Tall + (q * ws.nbq + b) * NBQ * NBQ;
  for (int idx = threadIdx.x; idx < pw * pw; idx += blockDim.x) {
    const int cc = idx / pw, r = idx - cc * pw;
    Tg[idx] = r <= cc ? Xp[cc * (cc + 1) / 2 + r] : 0.0;
  }
}

}  // namespace
}  // namespace gscfb
#include "trd_fused.cuh"
#include "trd_fused_blk.cuh"
namespace gscfb {
namespace {

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

struct Carver {
  char* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t count) {
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off = align_up(off + count * sizeof(T));
    return p;
  }
};

void carve_trd(Carver& cv, int n, int cap, TrdWs& ws) {
  const int ld = padded_ld(n);
  ws.n = n; ws.ld = ld; ws.cap = cap;
  ws.nrb = ceil_div(n, RB);
  ws.ncb = ceil_div(n, 32);
  ws.npart = std::max(ws.nrb, 1024);
  ws.d = cv.take<double>((size_t)cap * n);
  ws.e = cv.take<double>((size_t)cap * n);
  ws.tau = cv.take<double>((size_t)cap * n);
  ws.W = cv.take<double>((size_t)cap * ld * NB);
  ws.ypart = cv.take<double>((size_t)cap * ws.ncb * n);
  ws.pnorm = cv.take<double>((size_t)cap * ws.npart);
  ws.pdot = cv.take<double>((size_t)cap * ws.npart);
  ws.bar = cv.take<unsigned>((size_t)4 * cap + 1);  // counters of up to three phases (trd_fused), [4cap] abort flag
  ws.VPW = cv.take<double>((size_t)cap * 2 * 32 * round_up(n, 2));
  ws.PQ = cv.take<double>((size_t)cap * 2 * 256 * 64);
  ws.VW = cv.take<double>((size_t)cap * ld * 4 * NB);
  ws.SK = cv.take<double>((size_t)std::min(cap, 16) * 16 * 2 * NBQ * n);
  ws.amax = cv.take<unsigned long long>((size_t)cap);
  ws.ascale = cv.take<double>((size_t)cap);
  ws.X = cv.take<double>((size_t)cap * tf_x_doubles(n));
  ws.p12 = cv.take<double>((size_t)cap * ws.nrb * 2 * NB);
  ws.scal = cv.take<double>((size_t)cap * 8);
  ws.nbw = n <= 256 ? 64 : NBQ;
  ws.nbq = std::max(1, ceil_div(n - 1, ws.nbw));
  ws.Vf = cv.take<double>((size_t)cap * ld * n);
  ws.Wf = cv.take<double>((size_t)cap * ld * n);
  ws.Gall = cv.take<double>((size_t)cap * ws.nbq * NBQ * NBQ);
  ws.Tall = cv.take<double>((size_t)cap * ws.nbq * NBQ * NBQ);
  ws.bdesc = cv.take<GemmDesc>((size_t)cap * ws.nbq);
  ws.Yw = cv.take<double>((size_t)cap * 2 * NBQ * n);
}

int trailing_update(int n, double* A, int lda, long long sA, const TrdWs& ws, int j0, int pw, int batch,
                    const int* map, cudaStream_t st);

int tridiagonalise_legacy(int n, double* A, int lda, long long sA, const TrdWs& ws, int batch, const int* map,
                          cudaStream_t st) {
  int nrb_prev_dot = 0;  // row blocks that wrote pdot in the previous K3
  if (n == 1) {
    trd_col_update<<<dim3(1, batch), RB, 0, st>>>(A, lda, sA, ws, 0, 0, 0, map);
    count_launch();
    GSCF_CHECK_LAUNCH();
    return GSCF_B200_OK;
  }
  for (int j0 = 0; j0 < n - 1; j0 += NB) {
    const int pw = std::min(NB, n - 1 - j0);
    for (int i = j0; i < j0 + pw; ++i) {
      const int nrb1 = ceil_div(n - i, RB);
      trd_col_update<<<dim3(nrb1, batch), RB, 0, st>>>(A, lda, sA, ws, i, j0, nrb_prev_dot, map);
      const int m = n - i - 1;  // trailing order
      const int ntr = ceil_div(m, TS), ntc = ceil_div(m, TS);
      const bool sample = (i % 8) == 4;  // bench.py roofline: every 8th symv launch is event-timed
      if (sample) phase_mark(PH_SYMV_SAMPLE, true, st);
      trd_symv<<<dim3(ntr * ntc + ntr, batch), TS, 0, st>>>(A, lda, sA, ws, i, j0, nrb1, ntr, ntc, map);
      if (sample) phase_mark(PH_SYMV_SAMPLE, false, st);
      const int nrb3 = ceil_div(m, RB);
      trd_assemble_w<<<dim3(nrb3, batch), RB, 0, st>>>(A, lda, sA, ws, i, j0, nrb1, ntr, ntc, map);
      count_launch(3);
      nrb_prev_dot = nrb3;
    }
    const int ilast = j0 + pw - 1;
    if (j0 + pw == n - 1) {
      // last panel: the diagonal entry of the final column still needs the deferred update
      trd_col_update<<<dim3(1, batch), RB, 0, st>>>(A, lda, sA, ws, n - 1, j0, nrb_prev_dot, map);
      count_launch();
    } else {
      const int m = n - ilast - 1;
      trd_finalize_w<<<dim3(ceil_div(m, RB), batch), RB, 0, st>>>(A, lda, sA, ws, ilast, j0, nrb_prev_dot, map);
      count_launch();
      // trailing update A22 -= V2 W2^T + W2 V2^T on rows/cols >= j0 + pw: one rank-2pw GEMM pass
      GSCF_TRY(trailing_update(n, A, lda, sA, ws, j0, pw, batch, map, st));
      nrb_prev_dot = 0;
    }
    GSCF_CHECK_LAUNCH();
  }
  return GSCF_B200_OK;
}


__global__ void pack_vw_kernel(const double* __restrict__ Ab, int lda, long long sA, TrdWs ws, int j0, int pw, int r0,
                               const int* __restrict__ map) {
  // VW = [V2 | W2], WV = [W2 | V2]  (rows r0.., pw columns each) so that the rank-2k update is ONE GEMM
  const long long q = map ? map[blockIdx.y] : blockIdx.y;
  const double* A = Ab + q * sA;
  const double* W = ws.W + q * (long long)ws.ld * NB;
  double* VW = ws.VW + q * (long long)ws.ld * 4 * NB;
  double* WV = VW + (long long)ws.ld * 2 * NB;
  const int m = ws.n - r0;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < m * pw; idx += gridDim.x * blockDim.x) {
    const int j = idx / m, rr = idx - j * m;
    const double v = A[(size_t)(j0 + j) * lda + r0 + rr];
    const double w = W[(size_t)j * ws.ld + r0 + rr];
    VW[(size_t)j * ws.ld + rr] = v;
    VW[(size_t)(pw + j) * ws.ld + rr] = w;
    WV[(size_t)j * ws.ld + rr] = w;
    WV[(size_t)(pw + j) * ws.ld + rr] = v;
  }
}

// A22 -= V2 W2^T + W2 V2^T = [V2 | W2] [W2 | V2]^T on rows / columns >= j0 + pw (one pass over A22)
int trailing_update(int n, double* A, int lda, long long sA, const TrdWs& ws, int j0, int pw, int batch,
                    const int* map, cudaStream_t st) {
  const int r0 = j0 + pw, mt = n - r0;
  if (mt <= 0) return GSCF_B200_OK;
  phase_mark(PH_TRAILING_GEMM, true, st);
  pack_vw_kernel<<<dim3(std::max(1, std::min(64, ceil_div(mt * pw, 256))), batch), 256, 0, st>>>(A, lda, sA, ws, j0, pw,
                                                                                             r0, map);
  count_launch();
  GSCF_CHECK_LAUNCH();
  GemmParams g = gemm_params(mt, mt, 2 * pw, -1.0, ws.VW, ws.ld, ws.VW + (size_t)ws.ld * 2 * NB, ws.ld, 1.0,
                             A + (size_t)r0 * lda + r0, lda);
  g.sA = g.sB = (long long)ws.ld * 4 * NB; g.sC = sA; g.batch_map = map;
  GSCF_TRY(launch_gemm(false, true, g, batch, mt, mt, st));
  phase_mark(PH_TRAILING_GEMM, false, st);
  return GSCF_B200_OK;
}

// ---- fused one-launch tridiagonalisation (trd_fused.cuh) ---------------------------------------------------
template <int KMAX, bool SMEM_A, int T, bool SOLO = false, bool REGA = false, bool SYMH = false, int MINB = 1>
int tf_launch(const TfArgs& a, int G, int cnt, bool cluster, size_t smem, cudaStream_t st) {
  auto kern = trd_fused_kernel<KMAX, SMEM_A, T, SOLO, REGA, SYMH, MINB>;
  // function attributes are per device (and this library may drive several GPUs / host threads from one process):
  // set them on the current device before every launch - two cheap host calls per tridiagonalisation
  GSCF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  GSCF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G, cnt);
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (cluster) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  } else {
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
  }
  cfg.attrs = attr;
  cfg.numAttrs = SOLO ? 0 : 1;  // a single CTA per matrix waits for nobody: ordinary launch, any batch size
  GSCF_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
  count_launch();
  return GSCF_B200_OK;
}

template <int KMAX, int T>
int tfb_launch_t(const TfbArgs& a, int G, int cnt, size_t smem, cudaStream_t st) {
  auto kern = trd_fused_blocked_kernel<KMAX, T>;
  GSCF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));  // per device
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G, cnt);
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GSCF_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, a));
  count_launch();
  return GSCF_B200_OK;
}
int tfb_launch(const TfbArgs& a, int G, int cnt, int T, int K, size_t smem, cudaStream_t st) {
  if (T == 512 && K <= 4) return tfb_launch_t<4, 512>(a, G, cnt, smem, st);
  if (T == 512) return tfb_launch_t<8, 512>(a, G, cnt, smem, st);
  return tfb_launch_t<8, 1024>(a, G, cnt, smem, st);
}

template <int KMAX, bool SMEM_A, int T>
int tf_max_clusters(int G, size_t smem) {
  auto kern = trd_fused_kernel<KMAX, SMEM_A, T>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(G, 1);
  cfg.blockDim = dim3(T);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = G; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int mc = 0;
  if (cudaOccupancyMaxActiveClusters(&mc, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return 0; }
  return mc;
}

// trailing order below which deferring the rank-2 updates stops paying (measured on B200: n = 4096 eigensolve 120.5 ms
// unblocked, 111.5 ms always blocked, 108.4 ms with the switch at 2500)
int tf_unblocked_below() {
  static int ub = -1;
  if (ub < 0) { const char* ue = getenv("GSCF_B200_TF_UNBLOCKED_BELOW"); ub = ue ? atoi(ue) : 2500; }
  return ub;
}

// One phase: the first `jstop` reflectors (0 = all) of the n x n matrices at A (already offset to the trailing block of an
// earlier phase), d / e / tau offset alike.  prefer_cluster: one 16-CTA cluster per matrix even for a single matrix.
int tf_phase(int n, int n_full, double* A, int lda, long long sA, const TrdWs& ws, double* d, double* e, double* tau,
             unsigned* counters, int batch, const int* map, int jstop, bool prefer_cluster, cudaStream_t st) {
  static int g_single = 0, force_global = 0, solo_ok = 1, t_small = 256, blocked = 1, cb = 1;
  static std::atomic<bool> init{false};
  int sms = 0;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);  // of the CURRENT device, every call
  }
  if (!init.load(std::memory_order_acquire)) {  // environment switches: idempotent, so a race only repeats the parse
    const char* ge = getenv("GSCF_B200_TF_G");
    g_single = ge ? atoi(ge) : 0;
    const char* fg = getenv("GSCF_B200_TF_GLOBAL");
    force_global = (fg && fg[0] == '1') ? 1 : 0;
    const char* se = getenv("GSCF_B200_TF_SOLO");
    solo_ok = (se && se[0] == '0') ? 0 : 1;
    const char* te = getenv("GSCF_B200_TF_T");
    t_small = te ? atoi(te) : 256;
    const char* be = getenv("GSCF_B200_TF_BLOCKED");
    blocked = (be && be[0] == '0') ? 0 : 1;
    const char* ce = getenv("GSCF_B200_TF_CLUSTER_BARRIER");
    cb = (ce && ce[0] == '0') ? 0 : 1;
    init.store(true, std::memory_order_release);
  }
  constexpr size_t kSmemMax = 220 * 1024;
  int g_smem = 0;  // fewest CTAs with the columns in shared memory
  if (!force_global && n <= 2048) {
    for (int G = 1; G <= sms; ++G)
      if (tf_smem_bytes(n, G, true) <= kSmemMax) { g_smem = G; break; }
  }
  // group size and launch mode
  int G = 0;
  bool cluster = false, smem_a = false;
  const bool solo = solo_ok && !force_global && tf_smem_bytes(n, 1, true) <= kSmemMax;
  static int batch_global = -1;
  // measured (128 matrices of order 512, whole eigensolve): 16-CTA clusters in shared memory 43.6 ms, clusters of 4
  // working in L2 39.5 ms (37 matrices in flight instead of 8)
  if (batch_global < 0) { const char* bg = getenv("GSCF_B200_TF_BATCH_GLOBAL"); batch_global = bg ? atoi(bg) : 4; }
  if (solo) {
    G = 1; smem_a = true;
  } else if (batch_global > 0 && batch * g_smem > sms && g_smem > batch_global && n <= 2048) {
    // many matrices, each too large for the shared memory of a few CTAs: small clusters working in L2, more in flight
    G = batch_global;
    cluster = true; smem_a = false;
  } else if (g_smem > 0 && g_smem <= 16 && (prefer_cluster || batch * g_smem > sms)) {
    // one thread-block cluster per matrix, scheduled by the hardware: batches that do not fit side by side, and the
    // middle phase of a single reduction (the cluster barrier is cheaper than the L2 counter exchange)
    G = g_smem <= 8 ? 8 : 16;
    cluster = true; smem_a = true;
  } else {
    const int per = std::max(1, sms / std::max(1, std::min(batch, sms)));  // CTAs available per matrix
    const int gmin = g_smem > 0 && g_smem <= per ? g_smem : 1;
    if (gmin > sms) return -1;
    G = std::max(gmin, std::min(per, ceil_div(n, 6)));
    if (g_single > 0 && batch == 1) G = std::max(gmin, std::min(sms, g_single));
    G = std::min(G, sms);
    smem_a = g_smem > 0 && G >= g_smem;
  }
  // columns in registers: <= 8 columns per CTA and <= 4 rows per thread (n = 1024 on 148 SMs)
  static int rega_ok = -1;
  if (rega_ok < 0) { const char* re = getenv("GSCF_B200_TF_REGA"); rega_ok = (re && re[0] == '0') ? 0 : 1; }
  const bool rega = rega_ok && smem_a && !solo && !cluster && ceil_div(n, G) <= TF_RS && round_up(n, 2) <= TF_RROWS;
  const size_t smem = tf_smem_bytes(n, G, smem_a && !rega);
  static int t_big = -1;
  if (t_big < 0) { const char* tb = getenv("GSCF_B200_TF_TBIG"); t_big = tb ? atoi(tb) : 0; }
  // symmetric-half sweep (trd_fused.cuh SYMH): the small-cluster-in-L2 configuration of batches of mid-size matrices
  static int symh_ok = -1;
  if (symh_ok < 0) { const char* sh = getenv("GSCF_B200_TF_SYMH"); symh_ok = (sh && sh[0] == '0') ? 0 : 1; }
  const bool symh = symh_ok && cluster && !smem_a && !solo && G <= TF_SYMH_MAXG && jstop == 0 && n <= 2048;
  const int T = symh ? 256
                     : (smem_a ? ((t_small == 128 && n <= 1024) ? 128 : 256) : ((n <= 4096 && t_big != 1024) ? 512 : 1024));
  const int K = ceil_div(n, T);
  // (KMAX, SMEM_A, T) instantiations
  auto launch = [&](const TfArgs& a, int cnt) -> int {
    if (solo && n <= 128) {
      static int solo_reg = -1;
      if (solo_reg < 0) { const char* sr = getenv("GSCF_B200_TF_SOLO_REG"); solo_reg = (sr && sr[0] == '0') ? 0 : 1; }
      if (solo_reg) {
        // the register layout is fixed per instantiation: orders <= 64 (and the trailing block a larger matrix hands over)
        // run the 64-row layout, whose column step costs less than half of the 128-row one
        static int solo_reg_t = -1;
        if (solo_reg_t < 0) { const char* srt = getenv("GSCF_B200_TF_SOLO_REG_T"); solo_reg_t = srt ? atoi(srt) : 256; }
        if (n <= 64) trd_solo_reg_kernel<64><<<dim3(1, cnt), 256, 0, st>>>(a);
        else if (solo_reg_t == 512) trd_solo_reg_kernel<128, 512><<<dim3(1, cnt), 512, 0, st>>>(a);
        else trd_solo_reg_kernel<128><<<dim3(1, cnt), 256, 0, st>>>(a);
        count_launch();
        GSCF_CHECK_LAUNCH();
        return GSCF_B200_OK;
      }
    }
    if (solo) {
      static int solo_t = -1;
      if (solo_t < 0) { const char* se2 = getenv("GSCF_B200_TF_SOLO_T"); solo_t = se2 ? atoi(se2) : 256; }
      if (solo_t == 512) return tf_launch<1, true, 512, true>(a, G, cnt, false, smem, st);
      return tf_launch<1, true, 256, true>(a, G, cnt, false, smem, st);
    }
    if (rega) {
      static int rega_t = -1;
      if (rega_t < 0) { const char* rt = getenv("GSCF_B200_TF_REGA_T"); rega_t = rt ? atoi(rt) : 256; }
      if (rega_t == 512) return tf_launch<2, false, 512, false, true>(a, G, cnt, cluster, smem, st);
      return tf_launch<4, false, 256, false, true>(a, G, cnt, cluster, smem, st);
    }
    if (symh) {  // small clusters in L2, lower triangle only
      // two CTAs per SM (128 registers per thread, a few spilled values) hide the exchange latency of one group behind
      // the sweep of another - measured on B200, 512 matrices of order 512 (C5): 39.0 ms instead of 47.7 ms per batch
      // (the full-column sweep: 61.5 ms); GSCF_B200_TF_SYMH_OCC=1 selects the one-CTA-per-SM build
      static int occ2 = -1;
      if (occ2 < 0) { const char* oe = getenv("GSCF_B200_TF_SYMH_OCC"); occ2 = (oe && oe[0] == '1') ? 0 : 1; }
      if (occ2 && K <= 2) return tf_launch<2, false, 256, false, false, true, 2>(a, G, cnt, cluster, smem, st);
      if (occ2 && K <= 4) return tf_launch<4, false, 256, false, false, true, 2>(a, G, cnt, cluster, smem, st);
      if (K <= 2) return tf_launch<2, false, 256, false, false, true>(a, G, cnt, cluster, smem, st);
      if (K <= 4) return tf_launch<4, false, 256, false, false, true>(a, G, cnt, cluster, smem, st);
      return tf_launch<8, false, 256, false, false, true>(a, G, cnt, cluster, smem, st);
    }
    if (smem_a) {
      if (T == 128) return tf_launch<8, true, 128>(a, G, cnt, cluster, smem, st);
      if (K <= 2) return tf_launch<2, true, 256>(a, G, cnt, cluster, smem, st);
      if (K <= 4) return tf_launch<4, true, 256>(a, G, cnt, cluster, smem, st);
      return tf_launch<8, true, 256>(a, G, cnt, cluster, smem, st);
    }
    if (T == 512 && K <= 4) return tf_launch<4, false, 512>(a, G, cnt, cluster, smem, st);
    if (T == 512) return tf_launch<8, false, 512>(a, G, cnt, cluster, smem, st);
    return tf_launch<8, false, 1024>(a, G, cnt, cluster, smem, st);
  };
  const bool use_blocked = blocked && !smem_a && !cluster && G <= 256 && TFB_PW <= 32 && n > tf_unblocked_below() &&
                           ceil_div(n, G) <= TF_MAXS && (jstop == 0 || n - jstop < tf_unblocked_below());
  if (cluster && smem_a) {
    int mc = T == 128 ? tf_max_clusters<8, true, 128>(G, smem) : K <= 2 ? tf_max_clusters<2, true, 256>(G, smem)
                    : (K <= 4 ? tf_max_clusters<4, true, 256>(G, smem) : tf_max_clusters<8, true, 256>(G, smem));
    if (mc < 1) return -1;
  }
  TfArgs a;
  a.A = A; a.lda = lda; a.sA = sA; a.X = ws.X; a.xstride = tf_x_doubles(n_full); a.d = d; a.e = e; a.tau = tau; a.dstride = n_full;
  a.abort_flag = ws.bar + 4 * ws.cap; a.ctr = counters; a.map = map; a.mat0 = 0; a.n = n; a.smax = tf_smax(n, G);
  a.jstop = jstop;
  a.cluster = (cluster && cb) ? 1 : 0;
  const int per_launch = (cluster || solo) ? 32768 : std::max(1, sms / G);
  for (int mat0 = 0; mat0 < batch; mat0 += per_launch) {
    a.mat0 = mat0;
    const int cnt = std::min(per_launch, batch - mat0);
    if (use_blocked) {
      TfbArgs ab;
      ab.b = a; ab.VP = ws.VPW; ab.WP = ws.VPW + (size_t)ws.cap * 32 * round_up(n_full, 2); ab.PQ = ws.PQ;
      ab.ctr2 = counters + ws.cap;
      ab.unblocked_below = tf_unblocked_below();
      GSCF_TRY(tfb_launch(ab, G, cnt, T, K, tfb_smem_bytes(n), st));
    } else {
      GSCF_TRY(launch(a, cnt));
    }
  }
  return GSCF_B200_OK;
}

// The reduction is recursive: once reflectors 0 .. j-1 are applied, what remains is the tridiagonalisation of the
// trailing block A(j:, j:).  The per-column cost is the exchange between the CTAs of a group (~3 600 cycles through L2
// for any group size, less inside a cluster, nothing for a single CTA), so the group shrinks with the matrix:
//   all SMs (matrix in shared memory or L2)  ->  one 16-CTA cluster (order <= ~590)  ->  one CTA (order <= ~144).
// Returns -1 when the fused kernels do not cover the case (caller falls back to the panel kernels).
int tridiagonalise_fused(int n, double* A, int lda, long long sA, const TrdWs& ws, int batch, const int* map,
                         cudaStream_t st) {
  static int enabled = -1, phases_on = 0, t_cluster = 592, t_solo = 144, handover_on = 1, t_smem = 1664, solo_split = 1,
             solo_tail = 1;
  if (enabled < 0) {
    const char* sse = getenv("GSCF_B200_TF_SOLO_SPLIT");
    solo_split = (sse && sse[0] == '0') ? 0 : 1;
    { const char* sr = getenv("GSCF_B200_TF_SOLO_REG"); if (sr && sr[0] == '0') solo_split = 0; }
    { const char* so = getenv("GSCF_B200_TF_SOLO"); if (so && so[0] == '0') solo_split = 0; }
    { const char* fg = getenv("GSCF_B200_TF_GLOBAL"); if (fg && fg[0] == '1') solo_split = 0; }
    // one or two large matrices: the last 128 columns leave the all-SM kernel (3.5 us per column: its floor is the L2
    // exchange between 148 CTAs) for the single-CTA register kernels (no exchange: ~1.7 and 0.7 us per column)
    const char* ste = getenv("GSCF_B200_TF_SOLO_TAIL");
    solo_tail = (ste && ste[0] == '0') ? 0 : solo_split;
    const char* env = getenv("GSCF_B200_TRD_FUSED");
    enabled = (env && env[0] == '0') ? 0 : 1;
    int dev = 0, coop = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop || sms < 1) enabled = 0;
    // measured at n = 1024 (B200): single phase 5.39 ms per eigensolve, SOLO tail only 5.38 ms, cluster + SOLO 5.93 ms -
    // the 16-CTA cluster's longer sweep (37 columns per CTA) costs more than its cheaper barrier saves, and one CTA
    // still spends ~6 000 cycles per column.  Off by default; GSCF_B200_TF_PHASES=1 enables the hand-over.
    const char* pe = getenv("GSCF_B200_TF_PHASES");
    phases_on = (pe && pe[0] == '1') ? 1 : 0;
    const char* tc = getenv("GSCF_B200_TF_PHASE_CLUSTER");
    if (tc) t_cluster = atoi(tc) & ~1;
    const char* ts = getenv("GSCF_B200_TF_PHASE_SOLO");
    if (ts) t_solo = atoi(ts) & ~1;
    const char* he = getenv("GSCF_B200_TF_HANDOVER");
    if (he) { handover_on = atoi(he) > 0 ? 1 : 0; if (atoi(he) > 1) t_smem = atoi(he) & ~1; }
  }
  if (!enabled || n < 3 || (lda & 1) || n > 8192) return -1;
  GSCF_CUDA_TRY(cudaMemsetAsync(ws.bar, 0, ((size_t)4 * ws.cap + 1) * sizeof(unsigned), st));
  phase_mark(PH_SYMV_SAMPLE, true, st);
  int off = 0, phase = 0;
  while (true) {
    const int ns = n - off;
    int jstop = 0;
    // a single matrix too large for shared memory: once the trailing block fits (order t_smem on all SMs) hand it to the
    // shared-memory kernel - its sweeps cost half of the L2 ones and its O(n) part handles fewer rows per thread
    if (handover_on && batch == 1 && phase == 0 && ns > t_smem + 128) jstop = (ns - t_smem) & ~1;
    else if (solo_split && ns > 96 && ns <= 128) jstop = (ns - 64) & ~1;  // register kernel: 128-row -> 64-row layout
    else if (solo_tail && batch <= 2 && ns > 192 && ns <= 2048) jstop = (ns - 128) & ~1;  // tail of a large matrix -> one CTA
    else if (phases_on && ns <= tf_unblocked_below() && ns <= 2048 && phase < 2) {
      if (t_cluster > t_solo && ns > t_cluster + 64 && batch <= 8) jstop = ns - t_cluster;
      else if (t_solo > 0 && ns > t_solo + 32) jstop = ns - t_solo;
      jstop &= ~1;
    }
    const bool prefer_cluster = phases_on && phase > 0 && batch <= 8 && ns <= t_cluster + 64;
    // counters: phase 0 uses [0, 2 cap) (y exchange + the blocked kernel's partial dots), phases 1, 2 one array each
    // (the single-CTA kernels of later phases use no counters)
    unsigned* counters = phase == 0 ? ws.bar : ws.bar + (size_t)std::min(phase + 1, 3) * ws.cap;
    const int rc = tf_phase(ns, n, A + (size_t)off * lda + off, lda, sA, ws, ws.d + off, ws.e + off, ws.tau + off, counters,
                            batch, map, jstop, prefer_cluster, st);
    if (rc != GSCF_B200_OK) {
      if (rc < 0 && phase == 0) return -1;
      if (rc < 0) { set_last_error("fused tridiagonalisation: no configuration for a trailing phase"); return GSCF_B200_ERR_CUDA; }
      return rc;
    }
    if (jstop == 0) break;
    off += jstop;
    ++phase;
  }
  phase_mark(PH_SYMV_SAMPLE, false, st);
  return GSCF_B200_OK;
}

// Dispatch: the fused one-launch reduction; the three-kernels-per-column scheme (the dependencies of dlatrd as separate
// launches + deferred rank-2k DMMA GEMM updates) covers what the fused kernels do not (n < 3, odd leading dimension,
// n > 8192, no cooperative launch) and stays selectable for A/B runs (GSCF_B200_TRD_LEGACY=1 or GSCF_B200_TRD_FUSED=0).
int tridiagonalise(int n, double* A, int lda, long long sA, const TrdWs& ws, int batch, const int* map,
                   cudaStream_t st) {
  static int legacy = -1;
  if (legacy < 0) {
    const char* env = getenv("GSCF_B200_TRD_LEGACY");
    legacy = (env && env[0] == '1') ? 1 : 0;
  }
  if (!legacy && n >= 3) {
    const int fs = tridiagonalise_fused(n, A, lda, sA, ws, batch, map, st);
    if (fs >= 0) return fs;
  }
  return tridiagonalise_legacy(n, A, lda, sA, ws, batch, map, st);
}

// ---- back-transformation ---------------------------------------------------------------------------------------
// All reflectors as a unit lower trapezoidal matrix with explicit zeros: Vf(:, c) = [0 .. 0, 1, v_c], unit at row c+1.
// cpb columns per block (blockIdx.y indexes column chunks): short columns (batches of small matrices) take one warp per
// column, a block per column was bound by the block scheduling rate there.
__global__ void extract_v_all_kernel(const double* __restrict__ Ab, int lda, long long sA, TrdWs ws, int cpb,
                                     const int* __restrict__ map) {
  const long long q = map ? map[blockIdx.z] : blockIdx.z;
  const int n = ws.n;
  if (cpb == 1) {
    const int c = blockIdx.y;
    const double* a = Ab + q * sA + (size_t)c * lda;
    double* v = ws.Vf + q * (long long)ws.ld * n + (size_t)c * ws.ld;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
      v[r] = (c >= n - 1 || r <= c) ? 0.0 : (r == c + 1 ? 1.0 : a[r]);
    return;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int c = blockIdx.y * cpb + warp; c < min(n, (int)(blockIdx.y + 1) * cpb); c += nwarp) {
    const double* a = Ab + q * sA + (size_t)c * lda;
    double* v = ws.Vf + q * (long long)ws.ld * n + (size_t)c * ws.ld;
    for (int r = lane; r < n; r += 32) v[r] = (c >= n - 1 || r <= c) ? 0.0 : (r == c + 1 ? 1.0 : a[r]);
  }
}

// descriptors of the per-block set-up GEMMs: which = 0: G_b = V_b^T V_b,  1: W_b = V_b T_b
__global__ void bt_desc_kernel(TrdWs ws, int which, int batch, const int* __restrict__ map) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * ws.nbq) return;
  const int z = idx / ws.nbq, b = idx - z * ws.nbq;
  const long long q = map ? map[z] : z;
  const int n = ws.n, j0 = b * ws.nbw, pw = min(ws.nbw, n - 1 - j0), m = n - j0 - 1;
  const double* Vb = ws.Vf + q * (long long)ws.ld * n + (size_t)j0 * ws.ld + j0 + 1;
  GemmDesc d;
  if (which == 0) {
    d.A = Vb; d.lda = ws.ld; d.B = Vb; d.ldb = ws.ld;
    d.C = ws.Gall + (q * ws.nbq + b) * NBQ * NBQ; d.ldc = pw;
    d.m = pw; d.n = pw; d.k = m;
  } else {
    d.A = Vb; d.lda = ws.ld; d.B = ws.Tall + (q * ws.nbq + b) * NBQ * NBQ; d.ldb = pw;
    d.C = ws.Wf + q * (long long)ws.ld * n + (size_t)j0 * ws.ld + j0 + 1; d.ldc = ws.ld;
    d.m = m; d.n = pw; d.k = pw;
  }
  d.alpha = 1.0; d.beta = 0.0;
  ws.bdesc[idx] = d;
}

// Merging two consecutive block reflectors into one of twice the width (k = 256 instead of 128 in the Z-dependent GEMMs,
// half as many of them):  (I - W0 V0^T)(I - W1 V1^T) = I - [W0 | W1 - W0 (V0^T W1)] [V0 | V1]^T.
// which = 2:  M = V0(j1+1:, :)^T W1           (nbw x pw1, into the G slot of block b0 - larft is done with it)
// which = 3:  W1' = W1 - W0 M, two descriptors per pair: rows j0+1 .. j1 (W1 is zero there: beta = 0) and rows j1+1 ..
__global__ void bt_pair_desc_kernel(TrdWs ws, int which, int batch, int npairs, const int* __restrict__ map) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * npairs) return;
  const int z = idx / npairs, p = idx - z * npairs;
  const long long q = map ? map[z] : z;
  const int n = ws.n, nbw = ws.nbw, b0 = 2 * p, j0 = b0 * nbw, j1 = j0 + nbw;
  const int pw1 = min(nbw, n - 1 - j1), m1 = n - j1 - 1;
  double* Vq = ws.Vf + q * (long long)ws.ld * n;
  double* Wq = ws.Wf + q * (long long)ws.ld * n;
  double* M = ws.Gall + (q * ws.nbq + b0) * NBQ * NBQ;
  if (which == 2) {
    GemmDesc d;
    d.A = Vq + (size_t)j0 * ws.ld + j1 + 1; d.lda = ws.ld;
    d.B = Wq + (size_t)j1 * ws.ld + j1 + 1; d.ldb = ws.ld;
    d.C = M; d.ldc = nbw;
    d.m = nbw; d.n = pw1; d.k = m1; d.alpha = 1.0; d.beta = 0.0;
    ws.bdesc[idx] = d;
  } else {
    GemmDesc top, bot;
    top.A = Wq + (size_t)j0 * ws.ld + j0 + 1; top.lda = ws.ld; top.B = M; top.ldb = nbw;
    top.C = Wq + (size_t)j1 * ws.ld + j0 + 1; top.ldc = ws.ld;
    top.m = nbw; top.n = pw1; top.k = nbw; top.alpha = -1.0; top.beta = 0.0;
    bot.A = Wq + (size_t)j0 * ws.ld + j1 + 1; bot.lda = ws.ld; bot.B = M; bot.ldb = nbw;
    bot.C = Wq + (size_t)j1 * ws.ld + j1 + 1; bot.ldc = ws.ld;
    bot.m = m1; bot.n = pw1; bot.k = nbw; bot.alpha = -1.0; bot.beta = 1.0;
    ws.bdesc[2 * idx] = top;
    ws.bdesc[2 * idx + 1] = bot;
  }
}

// Measured (B200): one matrix of order 1024 back-transformation 0.47 -> 0.32 ms, order 4096 11.5 -> 9.5 ms; batches are
// in the throughput regime, where the wider blocks gain nothing (4096 x 128: 29.7 -> 27.5 ms but the longer set-up slows
// the concurrent D&C stage by as much) or lose (512 x 512: 154.6 -> 173.8 ms): pairs only for fewer than 16 matrices.
bool bt_pairing(int batch) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("GSCF_B200_BT_PAIR"); on = !e ? 1 : (e[0] == '0' ? 0 : (e[0] == '2' ? 2 : 1)); }
  return on == 2 || (on == 1 && batch < 16);
}

// Z <- Q Z with the reflectors stored in A / tau:  Q = H(0) .. H(n-2) = prod_b (I - V_b T_b V_b^T).
// Everything that does not depend on Z is done up front for all blocks at once (V, G_b, T_b, W_b = V_b T_b: four
// launches); the sequential part is two GEMMs per block:  Y = V_b^T Z,  Z -= W_b Y.
// Z-independent part: V (explicit), G_b = V_b^T V_b, T_b, W_b = V_b T_b for all blocks.  Depends only on the
// tridiagonalisation, so the caller runs it on a side stream while the divide & conquer stage occupies the main one.
int back_transform_setup(int n, const double* A, int lda, long long sA, const TrdWs& ws, int batch, const int* map,
                         cudaStream_t st) {
  if (n < 2) return GSCF_B200_OK;
  const int nblocks = ws.nbq;
  if (n <= 256)
    extract_v_all_kernel<<<dim3(1, ceil_div(n, 32), batch), 256, 0, st>>>(A, lda, sA, ws, 32, map);
  else
    extract_v_all_kernel<<<dim3(std::max(1, std::min(8, ceil_div(n, 256))), n, batch), 256, 0, st>>>(A, lda, sA, ws, 1, map);
  bt_desc_kernel<<<ceil_div(batch * nblocks, 128), 128, 0, st>>>(ws, 0, batch, map);
  count_launch(2);
  GSCF_CHECK_LAUNCH();
  {
    GemmParams gp{};
    gp.descs = ws.bdesc; gp.kbatch = 1; gp.ksplit = 1;
    GSCF_TRY(launch_gemm(true, false, gp, batch * nblocks, ws.nbw, ws.nbw, st));
  }
  larft_kernel<<<dim3(nblocks, batch), 4 * ws.nbw, (size_t)ws.nbw * (ws.nbw + 1) * sizeof(double), st>>>(ws, map);  // two packed triangles
  bt_desc_kernel<<<ceil_div(batch * nblocks, 128), 128, 0, st>>>(ws, 1, batch, map);
  count_launch(2);
  GSCF_CHECK_LAUNCH();
  {
    GemmParams gp{};
    gp.descs = ws.bdesc; gp.kbatch = 1; gp.ksplit = 1;
    GSCF_TRY(launch_gemm(false, false, gp, batch * nblocks, n, ws.nbw, st));
  }
  const int npairs = nblocks / 2;
  if (bt_pairing(batch) && npairs > 0) {
    bt_pair_desc_kernel<<<ceil_div(batch * npairs, 128), 128, 0, st>>>(ws, 2, batch, npairs, map);
    count_launch();
    GSCF_CHECK_LAUNCH();
    GemmParams gm{};
    gm.descs = ws.bdesc; gm.kbatch = 1; gm.ksplit = 1;
    GSCF_TRY(launch_gemm(true, false, gm, batch * npairs, ws.nbw, ws.nbw, st));
    bt_pair_desc_kernel<<<ceil_div(batch * npairs, 128), 128, 0, st>>>(ws, 3, batch, npairs, map);
    count_launch();
    GSCF_CHECK_LAUNCH();
    GemmParams gw{};
    gw.descs = ws.bdesc; gw.kbatch = 1; gw.ksplit = 1;
    GSCF_TRY(launch_gemm(false, false, gw, batch * npairs * 2, n, ws.nbw, st));
  }
  return GSCF_B200_OK;
}

int back_transform(int n, const TrdWs& ws, double* Z, int ldz, long long sZ, int batch, const int* map, cudaStream_t st,
                   int nvec) {
  if (n < 2) return GSCF_B200_OK;
  const int nz = (nvec > 0 && nvec < n) ? nvec : n;  // columns of Z that are wanted
  const int nref = n - 1;  // reflectors 0..n-2
  const int nblocks = ws.nbq;
  constexpr int LDY = 2 * NBQ;
  // one (merged) block reflector: pw reflectors starting at column j0
  auto apply = [&](int j0, int pw) -> int {
    const int m = n - j0 - 1;  // rows j0+1 .. n-1
    const double* Vb = ws.Vf + (size_t)j0 * ws.ld + j0 + 1;
    const double* Wb = ws.Wf + (size_t)j0 * ws.ld + j0 + 1;
    // Y = V_b^T Z(j0+1:, :)   (pw x nz, ld LDY)
    GemmParams y = gemm_params(pw, nz, m, 1.0, Vb, ws.ld, Z + j0 + 1, ldz, 0.0, ws.Yw, LDY);
    y.sA = (long long)ws.ld * n; y.sB = sZ; y.sC = (long long)LDY * n; y.batch_map = map;
    const int splits = batch >= 16 ? 1 : std::max(1, std::min(8, m / 128));  // long-k TN product: split-k
    GSCF_TRY(launch_gemm_splitk(true, false, y, batch, splits, ws.SK, st));
    // Z(j0+1:, :) -= W_b Y
    GemmParams u = gemm_params(m, nz, pw, -1.0, Wb, ws.ld, ws.Yw, LDY, 1.0, Z + j0 + 1, ldz);
    u.sA = (long long)ws.ld * n; u.sB = (long long)LDY * n; u.sC = sZ; u.batch_map = map;
    return launch_gemm(false, false, u, batch, m, nz, st);
  };
  if (bt_pairing(batch) && nblocks >= 2) {
    // blocks (2p, 2p+1) were merged by the set-up; an odd last block stands alone
    if (nblocks & 1) {
      const int j0 = (nblocks - 1) * ws.nbw;
      GSCF_TRY(apply(j0, std::min(ws.nbw, nref - j0)));
    }
    for (int p = nblocks / 2 - 1; p >= 0; --p) {
      const int j0 = 2 * p * ws.nbw, j1 = j0 + ws.nbw;
      GSCF_TRY(apply(j0, ws.nbw + std::min(ws.nbw, nref - j1)));
    }
  } else {
    for (int b = nblocks - 1; b >= 0; --b) {
      const int j0 = b * ws.nbw;
      GSCF_TRY(apply(j0, std::min(ws.nbw, nref - j0)));
    }
  }
  return GSCF_B200_OK;
}


// ---- back-transformation of small matrices (n <= 128: the C4 ensemble) ------------------------------------------------
// The blocked route above costs a batch of small matrices ten launches (V, V^T V, T, W = V T on the side stream, then two
// GEMMs per block of 64 reflectors), each of which streams the whole batch through HBM once or twice (4096 x 128 KB = 0.5 GB
// per operand, far beyond L2): measured 2.1 ms for the four GEMMs + 1.6 ms of set-up per eigensolve of 4096 matrices, the
// GEMMs at 9 - 20 TFLOP/s (`profiles/r02_ncu_sequence_eig_n128.txt`).  Here ONE CTA per matrix applies the reflectors one by
// one (LAPACK dorm2r: Z <- H_0 (H_1 ( ... H_{n-2} Z))) with all reflectors in shared memory and Z in registers: the batch is
// read once and written once, no T factors.  Thread (c = tid % 128, g = tid / 128) keeps rows 32 g .. 32 g + 31 of column c
// of Z; a reflector costs one partial dot per thread, one barrier, four shared-memory reads and one update pass.  The
// reflectors are stored with their explicit zeros and unit entry, so the passes need no row predicates; a row group that
// lies entirely above the reflector skips both passes (uniform per warp).
constexpr int BTS_T = 512;
constexpr int BTS_NMAX = 128;
// shared memory: reflectors [n][np] (the same region first / last stages Z for coalesced global accesses, column stride
// n | 1), tau [n], s_j = v_{j-1}^T v_j [n], partial dots [2 buffers][2 reflectors][4 row groups][128 columns]
inline size_t bts_smem_bytes(int n) {
  const int np = round_up(n, 32);  // whole row groups: the passes read 32 rows without a bound check, the padding is zero
  const size_t region = std::max((size_t)n * np, (size_t)n * (n | 1));
  return (region + 2 * (size_t)n + 2 * 2 * 4 * BTS_NMAX) * sizeof(double);
}
// Two reflectors per barrier:  H_b H_a Z = Z - v_a y_a - v_b y_b  with  y_a = tau_a v_a^T Z,
// y_b = tau_b (v_b^T Z - (v_b^T v_a) y_a): both partial dots come out of one pass over the thread's rows of Z, the scalar
// products s = v_b^T v_a of consecutive reflectors are formed once up front (one warp per pair).
__global__ void __launch_bounds__(BTS_T, 1)
bt_small_kernel(const double* __restrict__ Ab, int lda, long long sA, const double* __restrict__ taub, int n,
                double* __restrict__ Zb, int ldz, long long sZ, int nz, const int* __restrict__ map) {
  extern __shared__ __align__(16) double bsm[];
  const int np = (n + 31) & ~31;
  const int zs = n | 1;                      // odd column stride of the staged Z: conflict-free for lanes along the columns
  const size_t region = max((size_t)n * np, (size_t)n * zs);
  double* sV = bsm;                          // [n][np]: column j = reflector j = [0 .. 0, 1, v_j, 0 ..], unit entry at row j + 1
  double* stau = sV + region;                // [n]
  double* sdot = stau + n;                   // [n]: sdot[j] = v_{j-1}^T v_j
  double* sred = sdot + n;                   // [2][2][4][128]
  const long long q = map ? map[blockIdx.x] : blockIdx.x;
  const double* __restrict__ A = Ab + q * sA;
  double* __restrict__ Z = Zb + q * sZ;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARP = BTS_T / 32;
  const int c = tid & (BTS_NMAX - 1), g = tid >> 7, r0 = 32 * g;
  const bool mine = c < nz;
  // ---- Z: global -> shared (lanes along the rows) -> registers (lanes along the columns) ----------------------------
  for (int col = warp; col < nz; col += NWARP)
    for (int r = lane; r < n; r += 32) sV[(size_t)col * zs + r] = Z[(size_t)col * ldz + r];
  __syncthreads();
  double z[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) z[i] = (mine && r0 + i < n) ? sV[(size_t)c * zs + r0 + i] : 0.0;
  __syncthreads();
  // ---- reflectors -----------------------------------------------------------------------------------------------
  for (int j = warp; j < n; j += NWARP) {
    const double* a = A + (size_t)j * lda;
    for (int r = lane; r < np; r += 32)
      sV[(size_t)j * np + r] = (j < n - 1 && r < n && r > j + 1) ? a[r] : ((j < n - 1 && r == j + 1) ? 1.0 : 0.0);
  }
  for (int j = tid; j < n; j += BTS_T) stau[j] = j < n - 1 ? (taub + q * n)[j] : 0.0;
  __syncthreads();
  for (int j = 1 + warp; j < n - 1; j += NWARP) {
    double sdt = 0.0;
    for (int r = lane; r < np; r += 32) sdt = fma(sV[(size_t)(j - 1) * np + r], sV[(size_t)j * np + r], sdt);
    sdt = warp_sum(sdt);
    if (lane == 0) sdot[j] = sdt;
  }
  __syncthreads();
  int buf = 0;
  int j = n - 2;
  for (; j >= 1; j -= 2) {
    // pair (a = j, b = j - 1): H_a first
    const double ta = stau[j], tb = stau[j - 1];
    const bool active = r0 + 31 >= j && r0 < n;  // my row group reaches below the leading zeros of reflector b
    const double2* va = reinterpret_cast<const double2*>(sV + (size_t)j * np + r0);
    const double2* vb = reinterpret_cast<const double2*>(sV + (size_t)(j - 1) * np + r0);
    double pa0 = 0.0, pa1 = 0.0, pb0 = 0.0, pb1 = 0.0;
    if (active) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const double2 a0 = va[i / 2], a1 = va[i / 2 + 1], b0 = vb[i / 2], b1 = vb[i / 2 + 1];
        pa0 = fma(a0.x, z[i], pa0); pa1 = fma(a0.y, z[i + 1], pa1);
        pb0 = fma(b0.x, z[i], pb0); pb1 = fma(b0.y, z[i + 1], pb1);
        pa0 = fma(a1.x, z[i + 2], pa0); pa1 = fma(a1.y, z[i + 3], pa1);
        pb0 = fma(b1.x, z[i + 2], pb0); pb1 = fma(b1.y, z[i + 3], pb1);
      }
    }
    double* red = sred + buf * (2 * 4 * BTS_NMAX);
    red[g * BTS_NMAX + c] = pa0 + pa1;
    red[4 * BTS_NMAX + g * BTS_NMAX + c] = pb0 + pb1;
    __syncthreads();
    buf ^= 1;
    if (active) {
      const double ya = ta * ((red[c] + red[BTS_NMAX + c]) + (red[2 * BTS_NMAX + c] + red[3 * BTS_NMAX + c]));
      const double* rb = red + 4 * BTS_NMAX;
      const double yb = tb * (((rb[c] + rb[BTS_NMAX + c]) + (rb[2 * BTS_NMAX + c] + rb[3 * BTS_NMAX + c])) - sdot[j] * ya);
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const double2 a = va[i / 2], b = vb[i / 2];
        z[i] = fma(-b.x, yb, fma(-a.x, ya, z[i]));
        z[i + 1] = fma(-b.y, yb, fma(-a.y, ya, z[i + 1]));
      }
    }
  }
  if (j == 0) {  // an even number of rows leaves reflector 0 alone
    const double t0 = stau[0];
    const bool active = r0 + 31 >= 1 && r0 < n;
    const double2* v2 = reinterpret_cast<const double2*>(sV + r0);
    double p0 = 0.0, p1 = 0.0;
    if (active) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const double2 a = v2[i / 2];
        p0 = fma(a.x, z[i], p0); p1 = fma(a.y, z[i + 1], p1);
      }
    }
    double* red = sred + buf * (2 * 4 * BTS_NMAX);
    red[g * BTS_NMAX + c] = p0 + p1;
    __syncthreads();
    if (active) {
      const double y = t0 * ((red[c] + red[BTS_NMAX + c]) + (red[2 * BTS_NMAX + c] + red[3 * BTS_NMAX + c]));
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const double2 a = v2[i / 2];
        z[i] = fma(-a.x, y, z[i]);
        z[i + 1] = fma(-a.y, y, z[i + 1]);
      }
    }
  }
  // ---- Z: registers -> shared -> global -------------------------------------------------------------------------
  __syncthreads();  // every thread is done with the reflectors
  if (mine) {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (r0 + i < n) sV[(size_t)c * zs + r0 + i] = z[i];
  }
  __syncthreads();
  for (int col = warp; col < nz; col += NWARP)
    for (int r = lane; r < n; r += 32) Z[(size_t)col * ldz + r] = sV[(size_t)col * zs + r];
}

// Lanes along the rows (n = 128, all columns of Z): a WARP owns eight whole columns of Z - lane l keeps rows l, l + 32,
// l + 64, l + 96 of each - so the sums over the rows are warp-level (halving butterflies: eight partials -> one per lane with
// seven exchanges, two more for the full sum) and the reflector loop has NO block barrier; a lane reads four entries of a
// reflector instead of having 32 of them broadcast, and Z is read and written in place with coalesced accesses (no staging).
// Row slices that lie entirely above a reflector pair are skipped (uniform).
// Measured on B200 (4096 matrices of order 128, whole eigensolve): 14.59 ms against 13.29 ms with bt_small_kernel - the 18
// double shuffles and the dependent chain per pair (dots -> butterfly -> shared memory -> update) cost more than the block
// barrier and the broadcast loads they replace.  Correct (residuals as the other routes), NOT the default; kept behind
// GSCF_B200_BT_SMALL=2 as the measured alternative.
constexpr int BTW_T = 512;
inline size_t btw_smem_bytes() { return ((size_t)128 * 128 + 2 * 128 + (BTW_T / 32) * 16) * sizeof(double); }
__global__ void __launch_bounds__(BTW_T, 1)
bt_small_warp_kernel(const double* __restrict__ Ab, int lda, long long sA, const double* __restrict__ taub,
                     double* __restrict__ Zb, int ldz, long long sZ, const int* __restrict__ map) {
  constexpr int n = 128, np = 128, NWARP = BTW_T / 32;
  extern __shared__ __align__(16) double bsm[];
  double* sV = bsm;                 // [n][np]: column j = reflector j with explicit zeros and unit entry (row j + 1)
  double* stau = sV + n * np;       // [n]
  double* sdot = stau + n;          // [n]: sdot[j] = v_{j-1}^T v_j
  double* sy = sdot + n;            // [NWARP][16]: the sixteen sums of a warp's reflector pair (a: 0..7, b: 8..15)
  const long long q = map ? map[blockIdx.x] : blockIdx.x;
  const double* __restrict__ A = Ab + q * sA;
  double* __restrict__ Z = Zb + q * sZ;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double z[4][8];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) z[i][c] = Z[(size_t)(8 * warp + c) * ldz + lane + 32 * i];
  for (int j = warp; j < n; j += NWARP) {
    const double* a = A + (size_t)j * lda;
    for (int r = lane; r < np; r += 32)
      sV[j * np + r] = (j < n - 1 && r > j + 1) ? a[r] : ((j < n - 1 && r == j + 1) ? 1.0 : 0.0);
  }
  for (int j = tid; j < n; j += BTW_T) stau[j] = j < n - 1 ? (taub + q * n)[j] : 0.0;
  __syncthreads();
  for (int j = 1 + warp; j < n - 1; j += NWARP) {
    double sdt = 0.0;
    for (int r = lane; r < np; r += 32) sdt = fma(sV[(j - 1) * np + r], sV[j * np + r], sdt);
    sdt = warp_sum(sdt);
    if (lane == 0) sdot[j] = sdt;
  }
  __syncthreads();
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
  double* my = sy + warp * 16;
  // sums over the warp of eight per-lane values; lane (h16, h8, h4) returns the total of value 4 h16 + 2 h8 + h4
  auto reduce8 = [&](const double (&v)[8]) {
    double a4[4], a2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double send = h16 ? v[i] : v[i + 4], keep = h16 ? v[i + 4] : v[i];
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
    return a1;
  };
  const int slot = ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);  // 2 h8 + h4: column of the half this lane ends up with
  for (int j = n - 2; j >= 0; j -= 2) {
    // pair (a = j, b = j - 1), H_a first; the last reflector (j = 0) stands alone: b = none
    const bool pair = j >= 1;
    const double ta = stau[j], tb = pair ? stau[j - 1] : 0.0;
    const int i0 = (pair ? j : j + 1) >> 5;  // first row slice with a non-zero entry of either reflector
    double va[4], vb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      va[i] = i >= i0 ? sV[j * np + lane + 32 * i] : 0.0;
      vb[i] = (pair && i >= i0) ? sV[(j - 1) * np + lane + 32 * i] : 0.0;
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      double v[8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        double pa = 0.0, pb = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i >= i0) {
            pa = fma(va[i], z[i][4 * half + c], pa);
            pb = fma(vb[i], z[i][4 * half + c], pb);
          }
        v[c] = pa;
        v[4 + c] = pb;
      }
      const double tot = reduce8(v);
      if ((lane & 3) == 0) my[(h16 ? 8 : 0) + 4 * half + slot] = tot;
    }
    __syncwarp();
    const double sab = pair ? sdot[j] : 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const double ya = ta * my[c];
      const double yb = tb * (my[8 + c] - sab * ya);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i >= i0) z[i][c] = fma(-vb[i], yb, fma(-va[i], ya, z[i][c]));
    }
    __syncwarp();  // every lane has read the sums before the next pair overwrites them
  }
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int i = 0; i < 4; ++i) Z[(size_t)(8 * warp + c) * ldz + lane + 32 * i] = z[i][c];
}

// n <= 128 and the switch GSCF_B200_BT_SMALL (default on)
// GSCF_B200_BT_SMALL: 0 = blocked route, 1 (default) = bt_small_kernel, 2 = bt_small_warp_kernel where it applies
int bt_small_mode() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("GSCF_B200_BT_SMALL"); on = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1; }
  return on;
}
bool bt_small_route(int n) { return bt_small_mode() != 0 && n >= 2 && n <= BTS_NMAX; }

int back_transform_small(int n, const double* A, int lda, long long sA, const TrdWs& ws, double* Z, int ldz, long long sZ,
                         int batch, const int* map, cudaStream_t st, int nvec) {
  const int nz = (nvec > 0 && nvec < n) ? nvec : n;
  const size_t smem = bts_smem_bytes(n);
  static PerDeviceOnce once;
  int dev;
  if (once.need(dev)) {
    GSCF_CUDA_TRY(cudaFuncSetAttribute(bt_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)bts_smem_bytes(BTS_NMAX)));
    GSCF_CUDA_TRY(cudaFuncSetAttribute(bt_small_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)btw_smem_bytes()));
    once.done(dev);
  }
  if (bt_small_mode() == 2 && n == 128 && nz == n) {
    bt_small_warp_kernel<<<batch, BTW_T, btw_smem_bytes(), st>>>(A, lda, sA, ws.tau, Z, ldz, sZ, map);
    count_launch();
    GSCF_CHECK_LAUNCH();
    return GSCF_B200_OK;
  }
  bt_small_kernel<<<batch, BTS_T, smem, st>>>(A, lda, sA, ws.tau, n, Z, ldz, sZ, nz, map);
  count_launch();
  GSCF_CHECK_LAUNCH();
  return GSCF_B200_OK;
}

}  // namespace

size_t tridiag_dc_worksize(int n, int batch) {
  Carver cv{nullptr};
  TrdWs ws;
  carve_trd(cv, n, batch, ws);
  StedcWs dws;
  return stedc_carve(nullptr, cv.off, n, batch, dws) + 1024;
}

int tridiag_dc_syev_batched(int n, double* A, int lda, long long strideA, double* w, long long stride_w,
                            double* Z, int ldz, long long strideZ, int batch, const int* batch_map, void* work,
                            size_t work_bytes, int* info_dev, cudaStream_t st, int nvec) {
  if (batch <= 0 || n <= 0) return GSCF_B200_OK;
  // the workspace is indexed like the matrices (by batch_map[z] when given): capacity = what
  // the caller sized it for; we only check the lower bound for `batch` entries
  if (!work || work_bytes < tridiag_dc_worksize(n, batch)) {
    set_last_error("tridiag_dc_syev_batched: workspace missing or too small");
    return GSCF_B200_ERR_INVALID_ARG;
  }
  GSCF_CUDA_TRY(cudaFuncSetAttribute(larft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     NBQ * (NBQ + 1) * (int)sizeof(double)));  // per device: set before every use
  // capacity: derive from the workspace size so that batch_map indices stay in range
  int cap = batch;
  while (tridiag_dc_worksize(n, cap * 2) <= work_bytes && cap < (1 << 20)) cap *= 2;
  while (tridiag_dc_worksize(n, cap + 1) <= work_bytes) ++cap;
  Carver cv{(char*)work};
  TrdWs ws;
  carve_trd(cv, n, cap, ws);
  StedcWs dws;
  stedc_carve((char*)work, cv.off, n, cap, dws);
  phase_mark(PH_TRIDIAG, true, st);
  {
    // dsyev's scaling: a matrix whose norm is far from 1 is divided by it, the eigenvalues are multiplied back at the end
    // (squares of entries below 1e-154 / above 1e154 leave the double range in the reflector norms)
    GSCF_CUDA_TRY(cudaMemsetAsync(ws.amax, 0, (size_t)cap * sizeof(unsigned long long), st));
    const int gx = std::max(1, std::min(ceil_div(n * n, 256 * 16), std::max(1, 148 * 8 / batch)));
    absmax_kernel<<<dim3(gx, batch), 256, 0, st>>>(A, lda, strideA, n, ws.amax, batch_map);
    scale_extreme_kernel<<<dim3(gx, batch), 256, 0, st>>>(A, lda, strideA, n, ws.amax, ws.ascale, batch_map);
    count_launch(2);
    GSCF_CHECK_LAUNCH();
    dws.outer = ws.ascale;
  }
  GSCF_TRY(tridiagonalise(n, A, lda, strideA, ws, batch, batch_map, st));
  phase_mark(PH_TRIDIAG, false, st);
  // fork: the block factors of the back-transformation on a side stream, concurrently with the D&C stage
  // per host thread: engines on distinct streams may be driven from distinct threads
  static thread_local cudaStream_t side = nullptr;
  static thread_local cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  static thread_local int overlap = -1;
  if (overlap < 0) {
    const char* oe = getenv("GSCF_B200_BT_OVERLAP");
    overlap = (oe && oe[0] == '0') ? 0 : 1;
    if (overlap && (cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking) != cudaSuccess ||
                    cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess)) {
      cudaGetLastError();
      overlap = 0;
    }
  }
  const bool small_bt = bt_small_route(n);
  if (small_bt) {
    // one kernel after the D&C stage, nothing to prepare (see bt_small_kernel)
    phase_mark(PH_STEDC, true, st);
    GSCF_TRY(stedc_batched(n, ws.d, ws.e, n, w, stride_w, Z, ldz, strideZ, dws, batch, batch_map, info_dev, st, nvec));
    phase_mark(PH_STEDC, false, st);
    phase_mark(PH_BACKTRANSFORM, true, st);
    GSCF_TRY(back_transform_small(n, A, lda, strideA, ws, Z, ldz, strideZ, batch, batch_map, st, nvec));
    phase_mark(PH_BACKTRANSFORM, false, st);
    return GSCF_B200_OK;
  }
  if (overlap) {
    GSCF_CUDA_TRY(cudaEventRecord(ev_fork, st));
    GSCF_CUDA_TRY(cudaStreamWaitEvent(side, ev_fork, 0));
    GSCF_TRY(back_transform_setup(n, A, lda, strideA, ws, batch, batch_map, side));
    GSCF_CUDA_TRY(cudaEventRecord(ev_join, side));
  }
  phase_mark(PH_STEDC, true, st);
  GSCF_TRY(stedc_batched(n, ws.d, ws.e, n, w, stride_w, Z, ldz, strideZ, dws, batch, batch_map, info_dev, st, nvec));
  phase_mark(PH_STEDC, false, st);
  phase_mark(PH_BACKTRANSFORM, true, st);
  if (overlap) GSCF_CUDA_TRY(cudaStreamWaitEvent(st, ev_join, 0));
  else GSCF_TRY(back_transform_setup(n, A, lda, strideA, ws, batch, batch_map, st));
  GSCF_TRY(back_transform(n, ws, Z, ldz, strideZ, batch, batch_map, st, nvec));
  phase_mark(PH_BACKTRANSFORM, false, st);
  return GSCF_B200_OK;
}

}  // namespace gscfb
