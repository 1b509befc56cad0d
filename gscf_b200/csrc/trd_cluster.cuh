// Thread-block-cluster panel kernel of the Householder tridiagonalisation (included by tridiag.cu).
//
// For matrices that are L2-resident (n <= 2048) the per-column cost of the reduction is not
// bandwidth but the latency of its three global dependencies.  Here one *cluster* of C CTAs (C = 16
// for a single matrix, 8 in batches) owns a matrix: the dependencies become hardware cluster
// barriers (barrier.cluster) and everything that crosses CTAs (column a_i, partial norms / dots,
// V^T v, W^T v, row i+1 of V and W) is exchanged through distributed shared memory.  Each CTA owns
// a block of RPB = ceil(n/C) rows for the whole panel and keeps its rows of the panel columns of A,
// of V and of W in shared memory: inside the column loop there is NO global store (a release
// barrier would have to drain it to L2 - measured: ERRBAR + UCGABAR_WAIT dominated the first
// version) and the only global loads are the trailing-matrix product y = A22 v (row block x all
// columns, L2 -> SM, SYMV_U independent loads in flight per thread).  V, W, d, e, tau are written
// back once per panel.
#pragma once
#include <cooperative_groups.h>

namespace gscfb {
namespace {

namespace cg = cooperative_groups;

constexpr int CT = 512;        // threads per CTA
constexpr int SYMV_U = 16;     // independent loads in flight per thread in the trailing product
constexpr int CL_MAXN = 2048;  // largest order handled by the cluster kernel
constexpr int CL_MAXRW = 8;    // warps that own rows (RPB <= 256)

struct ClusterPub {            // published to the other CTAs of the cluster
  double norm_part[CL_MAXRW];
  double dot_part[CL_MAXRW];
  double head;
  double pad;
};

__global__ void __launch_bounds__(CT)
trd_cluster_panel_kernel(double* __restrict__ Ab, int lda, long long sA, TrdWs ws, int j0, int pw, int last_panel,
                         const int* __restrict__ map) {
  cg::cluster_group cluster = cg::this_cluster();
  const int C = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long q = map ? map[blockIdx.y] : blockIdx.y;
  const int n = ws.n;
  double* A = Ab + q * sA;
  double* W = ws.W + q * (long long)ws.ld * NB;

  const int RPB = (n + C - 1) / C;
  const int nrw = (RPB + 31) >> 5;  // warps owning rows
  int TPR = CT / RPB;               // threads per row in the symv (power of two, <= 16)
  TPR = TPR >= 16 ? 16 : (TPR >= 8 ? 8 : (TPR >= 4 ? 4 : (TPR >= 2 ? 2 : 1)));

  // dynamic shared memory (doubles): sAp, sVp, sWp [ (NB+1) * RPB ] each (column j at [j*RPB + rl]),
  // sa[RPB], sv[n + pad], sy[16 * RPB], sp12[2 NB], srowV/srowW[NB + 1]
  extern __shared__ __align__(16) double smem[];
  double* sAp = smem;
  double* sVp = sAp + (size_t)(NB + 1) * RPB;
  double* sWp = sVp + (size_t)(NB + 1) * RPB;
  double* sa = sWp + (size_t)(NB + 1) * RPB;
  double* sv = sa + RPB;
  double* sy = sv + n + SYMV_U * 16;
  double* sp12 = sy + (size_t)16 * RPB;
  double* srowV = sp12 + 2 * NB;
  double* srowW = srowV + (NB + 1);
  __shared__ ClusterPub pub;
  __shared__ double s_p1[NB], s_p2[NB], s_Vi[NB], s_Wi[NB];
  __shared__ double s_d[NB + 1], s_e[NB], s_tau[NB];
  __shared__ double s_bc[4];  // broadcast: alpha_prev, beta, tau, scale

  const int rl = tid % RPB;  // local row handled in the symv
  const int qs = tid / RPB;  // column slice in the symv (valid if < TPR)
  const int myrow = (tid < RPB) ? rank * RPB + tid : n;  // row owned in P1 / P3 (n = none)
  const int ncols = pw + (last_panel ? 1 : 0);

  // ---- prologue: my rows of the panel columns of A (one L2 round trip for the whole panel) ---------------
  for (int idx = tid; idx < ncols * RPB; idx += CT) {
    const int j = idx / RPB, t = idx - j * RPB;
    const int r = rank * RPB + t;
    sAp[(size_t)j * RPB + t] = (r < n) ? A[(size_t)(j0 + j) * lda + r] : 0.0;
  }
  if (tid < CL_MAXRW) { pub.norm_part[tid] = 0.0; pub.dot_part[tid] = 0.0; }
  __syncthreads();

#ifdef GSCF_TRD_DEBUG_TIMING
  long long tacc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long tmark = clock64();
#define TRD_TICK(k) do { const long long _t = clock64(); tacc[k] += _t - tmark; tmark = _t; } while (0)
#else
#define TRD_TICK(k) do { } while (0)
#endif
  double tau_prev = 0.0;  // tau of the previous column (every thread computes the scalars redundantly)
  for (int ci = 0; ci < ncols; ++ci) {
    const int i = j0 + ci, li = ci;
    const bool tail_only = (ci == pw);
    // ================= P1: finalise w_{i-1}, update column i ===============================================
    double alpha_prev = 0.0;
    if (li > 0) {
      // only warp 0 talks to the other CTAs (every warp doing so made the remote reads of the same
      // few words collide: ~1800 cycles instead of ~500); the result is broadcast through smem
      if (warp == 0) {
        const int owner = i / RPB;  // the CTA owning row i published row i of V and W (newest W entry raw)
        double rv = 0.0, rw = 0.0;
        if (lane < li) {
          rv = cluster.map_shared_rank(srowV, owner)[lane];
          rw = cluster.map_shared_rank(srowW, owner)[lane];
        }
        double s = 0.0;
        for (int k = lane; k < C * nrw; k += 32) s += cluster.map_shared_rank(&pub, k / nrw)->dot_part[k % nrw];
        const double ap = -0.5 * tau_prev * warp_sum(s);  // -1/2 tau_{i-1} w'^T v
        if (lane < li) {
          s_Vi[lane] = rv;
          s_Wi[lane] = rw + (lane == li - 1 ? ap : 0.0);
        }
        if (lane == 0) s_bc[0] = ap;
      }
      TRD_TICK(6);
      __syncthreads();
      alpha_prev = s_bc[0];
      TRD_TICK(7);
    }
    double part = 0.0;
    if (tid < RPB) {
      double a = 0.0;
      if (myrow >= i && myrow < n) {
        a = sAp[(size_t)ci * RPB + tid];
        if (li > 0) {
          const double vprev = sVp[(size_t)(li - 1) * RPB + tid];
          const double wfin = sWp[(size_t)(li - 1) * RPB + tid] + alpha_prev * vprev;
          sWp[(size_t)(li - 1) * RPB + tid] = wfin;
          double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;  // independent chains: DFMA latency, not count, bounds this
          int j = 0;
          for (; j + 2 <= li - 1; j += 2) {
            c0 = fma(sVp[(size_t)j * RPB + tid], s_Wi[j], c0);
            c1 = fma(sWp[(size_t)j * RPB + tid], s_Vi[j], c1);
            c2 = fma(sVp[(size_t)(j + 1) * RPB + tid], s_Wi[j + 1], c2);
            c3 = fma(sWp[(size_t)(j + 1) * RPB + tid], s_Vi[j + 1], c3);
          }
          for (; j < li - 1; ++j) {
            c0 = fma(sVp[(size_t)j * RPB + tid], s_Wi[j], c0);
            c1 = fma(sWp[(size_t)j * RPB + tid], s_Vi[j], c1);
          }
          a -= (c0 + c1) + (c2 + c3);
          a -= vprev * s_Wi[li - 1] + wfin * s_Vi[li - 1];
        }
        if (myrow == i) s_d[ci] = a;
        if (myrow == i + 1) pub.head = a;
        if (myrow >= i + 2) part = a * a;
      }
      sa[tid] = a;
    }
    if (tail_only) break;
    if (warp < nrw) {
      part = warp_sum(part);
      if (lane == 0) pub.norm_part[warp] = part;
    }
    TRD_TICK(0);
    cluster.sync();  // #1: a_i, partial norms, head published
    TRD_TICK(1);
    // ================= P2: v, y = A22 v, V^T v, W^T v ========================================================
    const int first = i + 1;
    double beta, tau, scale;
    if (warp == 0) {
      double s = 0.0;
      for (int k = lane; k < C * nrw; k += 32) s += cluster.map_shared_rank(&pub, k / nrw)->norm_part[k % nrw];
      const double head = cluster.map_shared_rank(&pub, first / RPB)->head;
      s = warp_sum(s);
      householder_scalars(head, s, beta, tau, scale);
      if (lane == 0) { s_e[ci] = beta; s_tau[ci] = tau; s_bc[1] = beta; s_bc[2] = tau; s_bc[3] = scale; }
    }
    __syncthreads();
    beta = s_bc[1]; tau = s_bc[2]; scale = s_bc[3];
    TRD_TICK(8);
    {
      int cidx = tid / RPB, lidx = tid - cidx * RPB;  // (owner CTA, local row) of element idx, updated incrementally
      const int dc = CT / RPB, dl = CT - dc * RPB;
      for (int idx = tid; idx < n + SYMV_U * 16; idx += CT) {
        double v = 0.0;
        if (idx == first) v = 1.0;
        else if (idx > first && idx < n) v = cluster.map_shared_rank(sa, cidx)[lidx] * scale;
        sv[idx] = v;
        cidx += dc; lidx += dl;
        if (lidx >= RPB) { lidx -= RPB; ++cidx; }
      }
    }
    __syncthreads();
    TRD_TICK(9);
    {
      const int r = rank * RPB + rl;
      double y0 = 0.0, y1 = 0.0, y2 = 0.0, y3 = 0.0;
      if (qs < TPR && r >= first && r < n) {
        const double* arow = A + r;
        for (int c = first + qs; c < n; c += SYMV_U * TPR) {
          double a[SYMV_U];
#pragma unroll
          for (int k = 0; k < SYMV_U; ++k) {
            const int cc = c + k * TPR;
            a[k] = cc < n ? arow[(size_t)cc * lda] : 0.0;
          }
#pragma unroll
          for (int k = 0; k < SYMV_U; k += 4) {
            y0 = fma(a[k + 0], sv[c + (k + 0) * TPR], y0);
            y1 = fma(a[k + 1], sv[c + (k + 1) * TPR], y1);
            y2 = fma(a[k + 2], sv[c + (k + 2) * TPR], y2);
            y3 = fma(a[k + 3], sv[c + (k + 3) * TPR], y3);
          }
        }
      }
      if (qs < TPR) sy[(size_t)qs * RPB + rl] = (y0 + y1) + (y2 + y3);
    }
    TRD_TICK(10);
    for (int j = warp; j < li; j += CT / 32) {
      double s1 = 0.0, s2 = 0.0;
      for (int t = lane; t < RPB; t += 32) {
        const int r = rank * RPB + t;
        if (r >= first && r < n) {
          s1 = fma(sVp[(size_t)j * RPB + t], sv[r], s1);
          s2 = fma(sWp[(size_t)j * RPB + t], sv[r], s2);
        }
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) { sp12[j] = s1; sp12[NB + j] = s2; }
    }
    TRD_TICK(2);
    cluster.sync();  // #2: V^T v / W^T v partials published; sy complete within the CTA
    TRD_TICK(3);
    // ================= P3: w = tau (y - V p2 - W p1), partial w^T v, publish row i+1 =========================
    if (li > 0) {
      // one half-warp per (which, j): its 16 lanes read the 16 CTAs' partials in parallel
      for (int item = warp * 2 + (lane >> 4); item < 2 * li; item += (CT / 32) * 2) {
        const int j = item % li, which = item / li, c = lane & 15;
        double s = (c < C) ? cluster.map_shared_rank(sp12, c)[which * NB + j] : 0.0;
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (c == 0) { if (which == 0) s_p1[j] = s; else s_p2[j] = s; }
      }
      __syncthreads();
    }
    part = 0.0;
    if (tid < RPB) {
      double v = 0.0, w = 0.0;
      if (myrow >= first && myrow < n) {
        double y = 0.0;
        for (int k = 0; k < TPR; ++k) y += sy[(size_t)k * RPB + tid];
        v = sv[myrow];
        double c0 = 0.0, c1 = 0.0, c2 = 0.0, c3 = 0.0;
        int j = 0;
        for (; j + 2 <= li; j += 2) {
          c0 = fma(sVp[(size_t)j * RPB + tid], s_p2[j], c0);
          c1 = fma(sWp[(size_t)j * RPB + tid], s_p1[j], c1);
          c2 = fma(sVp[(size_t)(j + 1) * RPB + tid], s_p2[j + 1], c2);
          c3 = fma(sWp[(size_t)(j + 1) * RPB + tid], s_p1[j + 1], c3);
        }
        for (; j < li; ++j) {
          c0 = fma(sVp[(size_t)j * RPB + tid], s_p2[j], c0);
          c1 = fma(sWp[(size_t)j * RPB + tid], s_p1[j], c1);
        }
        w = tau * (y - ((c0 + c1) + (c2 + c3)));
        part = w * v;
      }
      sVp[(size_t)li * RPB + tid] = v;  // rows above the reflector hold zeros
      sWp[(size_t)li * RPB + tid] = w;
    }
    {  // publish row i+1 of V and W for the next column (W entry li still raw): the owner's warp, one lane per entry
      const int t = first - rank * RPB;
      if (t >= 0 && t < RPB && warp == (t >> 5)) {
        __syncwarp();
        if (lane <= li) {
          srowV[lane] = sVp[(size_t)lane * RPB + t];
          srowW[lane] = sWp[(size_t)lane * RPB + t];
        }
      }
    }
    if (warp < nrw) {
      part = warp_sum(part);
      if (lane == 0) pub.dot_part[warp] = part;
    }
    tau_prev = tau;
    TRD_TICK(4);
    cluster.sync();  // #3: partial dots and row i+1 published
    TRD_TICK(5);
  }
  // ================= epilogue: finalise the last w, write the panel back =====================================
  if (!last_panel) {
    const int li = pw - 1;
    if (warp == 0) {
      double s = 0.0;
      for (int k = lane; k < C * nrw; k += 32) s += cluster.map_shared_rank(&pub, k / nrw)->dot_part[k % nrw];
      s = -0.5 * tau_prev * warp_sum(s);
      if (lane == 0) s_bc[0] = s;
    }
    __syncthreads();
    const double alpha = s_bc[0];
    if (tid < RPB) sWp[(size_t)li * RPB + tid] += alpha * sVp[(size_t)li * RPB + tid];
  }
  __syncthreads();
  for (int idx = tid; idx < pw * RPB; idx += CT) {
    const int j = idx / RPB, t = idx - j * RPB;
    const int r = rank * RPB + t;
    if (r < n) {
      if (r > j0 + j) A[(size_t)(j0 + j) * lda + r] = sVp[(size_t)j * RPB + t];  // v below the diagonal
      W[(size_t)j * ws.ld + r] = sWp[(size_t)j * RPB + t];
    }
  }
  {
    double* dv = ws.d + q * n;
    double* ev = ws.e + q * n;
    double* tauv = ws.tau + q * n;
    if (tid < ncols) {
      const int i = j0 + tid;
      if (i / RPB == rank) dv[i] = s_d[tid];
      if (rank == 0 && tid < pw) { ev[i] = s_e[tid]; tauv[i] = s_tau[tid]; }
    }
  }
#ifdef GSCF_TRD_DEBUG_TIMING
  if (tid == 0 && (rank == 0 || rank == C - 1) && j0 == 0)
    printf("rank %d cycles/col: P1 %lld (dotred %lld rows %lld) sync1 %lld P2 %lld (scal %lld gather %lld symv %lld) sync2 %lld P3 %lld sync3 %lld\n", rank, (tacc[0] + tacc[6] + tacc[7]) / pw, tacc[6] / pw, tacc[7] / pw,
           tacc[1] / pw, (tacc[2] + tacc[8] + tacc[9] + tacc[10]) / pw, tacc[8] / pw, tacc[9] / pw, tacc[10] / pw, tacc[3] / pw, tacc[4] / pw, tacc[5] / pw);
#endif
  cluster.sync();  // nobody may exit while its shared memory can still be read remotely
}

inline size_t trd_cluster_smem_bytes(int n, int C) {
  const int RPB = (n + C - 1) / C;
  return ((size_t)3 * (NB + 1) * RPB + RPB + n + SYMV_U * 16 + (size_t)16 * RPB + 2 * NB + 2 * (NB + 1)) *
         sizeof(double);
}

}  // namespace
}  // namespace gscfb
