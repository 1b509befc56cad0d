// Blocked variant of the fused tridiagonalisation (trd_fused.cuh) for matrices that do not fit in shared memory.
//
// Same ownership (column c -> CTA c mod G), same single all-gather of y per column.  What changes is the traffic:
// the unblocked kernel rewrites the whole trailing matrix for every column (16 (n-j)^2 bytes per step through
// L2 / HBM: 5.6 TB/s while the matrix exceeds L2, ~8.5 TB/s after).  Here up to PW reflectors stay *pending*:
//   * ordinary step j (k = j - j0 pending reflectors j0..j-1): my columns are only READ,
//         y_c = A0(:,c)^T v_j  -  sum_{i<k} [ V_i(c) (W_i^T v_j) + W_i(c) (V_i^T v_j) ],
//     except the few "eager" columns (j, j0+PW] that will become reflectors before the next boundary: their owner
//     keeps them up to date (one rank-2 update per step, as in the unblocked kernel), so the raw column j+1 it
//     publishes is exact.  The 2k dot products W_i^T v_j, V_i^T v_j are a second, small all-gather of per-CTA
//     partial sums; it is issued before the sweep and collected after it, i.e. overlapped with the matrix read.
//   * boundary step (k == PW): one read+write sweep applies all pending reflectors (rank-2PW update, FP64 FMA
//     bound) fused with the product for v_j.  Reflectors j0..j-2 come from the global panel buffers VP/WP (made
//     visible by the exchanges since they were written), reflector j-1 from shared memory.
// Traffic per column drops from 16 (n-j)^2 to ~8 (n-j)^2 (1 + 2/PW) bytes.
#pragma once

namespace gscfb {
namespace {

#ifndef GSCF_TFB_PW
#define GSCF_TFB_PW 32
#endif
constexpr int TFB_PW = GSCF_TFB_PW;  // pending reflectors between two boundary sweeps (<= 32: workspace and lane mappings)
static_assert(TFB_PW >= 4 && TFB_PW <= 32, "TFB_PW");

struct TfbArgs {
  TfArgs b;
  double* VP;      // [cap][PW][np]
  double* WP;      // [cap][PW][np]
  double* PQ;      // [cap][2][2 PW][gridDim.x] partial dots
  unsigned* ctr2;  // [cap]
  int unblocked_below;  // trailing order at which the kernel stops deferring updates
};

inline size_t tfb_smem_bytes(int n) {
  const int np = round_up(n, 2);
  return ((size_t)3 * np + 32 * TF_MAXS + 32 + 32 + 8 + 2 * TFB_PW * TF_MAXS + 2 * TFB_PW) * sizeof(double);
}

// Eight sums over the warp with 7 exchanges instead of 40 (every butterfly step halves the values a lane is responsible
// for): lanes 0, 4, ... 28 end with the sum of acc[lane / 4] and store it to out[lane / 4] if that index is < nvalid.
__device__ __forceinline__ void tfb_butterfly8(const double (&acc)[TF_CG], int lane, int nvalid, double* out) {
  static_assert(TF_CG == 8, "written for 8 accumulators");
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
  if ((lane & 3) == 0 && u < nvalid) out[u] = a1;
}

template <int KMAX, int T>
__global__ void __launch_bounds__(T, 1)
trd_fused_blocked_kernel(const TfbArgs ab) {
  const TfArgs& a = ab.b;
  constexpr int PW = TFB_PW;
  const int G = gridDim.x, g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = T / 32;
  const long long q = a.map ? a.map[blockIdx.y + a.mat0] : (long long)blockIdx.y + a.mat0;
  const int n = a.n, lda = a.lda, np = (n + 1) & ~1;
  double* __restrict__ A = a.A + q * a.sA;
  double* X = a.X + q * a.xstride;
  const long long xhalf = a.xstride >> 1;
  unsigned* ctr = a.ctr + q;
  unsigned* ctr2 = ab.ctr2 + q;
  double* VP = ab.VP + q * (long long)PW * np;
  double* WP = ab.WP + q * (long long)PW * np;
  double* PQ = ab.PQ + q * 2LL * G * 2 * PW;

  extern __shared__ __align__(16) double tf_smem[];
  double* sv0 = tf_smem;
  double* sv1 = sv0 + np;
  double* sw = sv1 + np;
  double* spart = sw + np;             // [32][TF_MAXS] per-warp partial y
  double* sred1 = spart + 32 * TF_MAXS;
  double* sred2 = sred1 + 32;
  double* sbc = sred2 + 32;            // 0: y[first]  1: a'[first+1]
  double* sVo = sbc + 8;               // [PW][TF_MAXS] V_i at my columns
  double* sWo = sVo + PW * TF_MAXS;    // [PW][TF_MAXS] W_i at my columns
  double* sPQ = sWo + PW * TF_MAXS;    // [2 PW]: P_i = W_i^T v_j, Q_i = V_i^T v_j

  const int S = g < n ? (n - g + G - 1) / G : 0;  // owned columns c = g + s G
  const int a_unblocked_below = ab.unblocked_below;

  for (int r = tid; r < np; r += T) { sv0[r] = 0.0; sv1[r] = 0.0; sw[r] = 0.0; }
  for (int r = tid; r < 2 * PW * TF_MAXS; r += T) sVo[r] = 0.0;
  if (tid < 8) sbc[tid] = 0.0;
  __syncthreads();

  // row r = tid + k T: my column slot (or -1) and whether I store it to the global panels (64-row chunks dealt cyclically)
  int own_slot[KMAX];
  unsigned chunk_mine = 0;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int r = tid + k * T;
    own_slot[k] = (r < n && r % G == g) ? r / G : -1;
    if (r < n && ((r >> 6) % G) == g) chunk_mine |= 1u << k;
  }
  double* svc = sv0;  // v_j      (zero above row j+1, one at row j+1)
  double* svn = sv1;  // v_{j-1} during the sweep, then v_{j+1}
  double tau_j = 0.0;
  int j0 = 0;         // first pending reflector
  unsigned nA = 0;    // partial-dot exchanges so far

  auto wait_counter = [&](unsigned* c, unsigned target) {
    if (*reinterpret_cast<volatile unsigned*>(a.abort_flag) == 0u) {
      const long long t0 = clock64();
      int spins = 0;
      while (tf_ld_acquire(c) < target) {
        if ((++spins & 255) == 0 &&
            (clock64() - t0 > 2000000000LL || *reinterpret_cast<volatile unsigned*>(a.abort_flag) != 0u)) {
          atomicExch(a.abort_flag, 1u);
          break;
        }
      }
    }
  };

#ifdef GSCF_TF_DEBUG_TIMING
  long long tacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // [0..8) deferred regime, [8..16) unblocked regime
  long long tmark = clock64();
  int treg = 0, tcols[2] = {0, 0};
#define TFB_TICK(k) do { const long long _t = clock64(); tacc[treg + (k)] += _t - tmark; tmark = _t; } while (0)
#else
#define TFB_TICK(k) do { } while (0)
#endif
  // jstop (see trd_fused.cuh): only honoured in the unblocked regime, where exactly one reflector is pending
  const int jlast = (a.jstop > 0 && a.jstop <= n - 2 && n - a.jstop < a_unblocked_below) ? a.jstop - 1 : n - 2;
  for (int j = -1; j <= jlast; ++j) {
    const int first = j + 1;
    double yv[KMAX], rv[KMAX];
    TFB_TICK(7);
#ifdef GSCF_TF_DEBUG_TIMING
    treg = (n - j > a_unblocked_below) ? 0 : 8;
    ++tcols[treg >> 3];
#endif
    if (j < 0) {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int r = tid + k * T;
        yv[k] = 0.0;
        rv[k] = r < n ? A[r] : 0.0;  // column 0 as it came in
      }
    } else {
      const int kp = j - j0;                 // pending reflectors j0 .. j-1
      // deferring pays while the trailing matrix streams from HBM; once it is L2-resident the second exchange and
      // the FMA-bound boundary sweeps cost more than the saved writes: panel width 1 = the unblocked algorithm
      const int pw_eff = (n - j > a_unblocked_below) ? PW : 1;
      const bool boundary = (kp >= pw_eff);
      const int s0 = first > g ? (first - g + G - 1) / G : 0;
      const int jend = j0 + PW;              // eager columns: (j, jend] (already hold reflectors j0 .. j-2)
      int se = jend >= g ? (jend - g) / G + 1 : 0;  // first slot with c > jend
      se = max(s0, min(S, se));
      const int r_lo = first & ~1;
      const int npairs = (np - r_lo) >> 1;
      double* Xy = X + (size_t)(j & 1) * xhalf;
      double* Xr = Xy + np;
      const bool pub_raw = (g + s0 * G == first);
      double* PQs = PQ + (size_t)(nA & 1) * G * 2 * PW;
      const bool exch_a = !boundary && kp > 0;
      // ---- partial dots of v_j with the pending reflectors over my index set (second, overlapped all-gather) ----
      if (exch_a) {
        {  // 8 lanes per dot product (2 kp <= 62 dots, T >= 512)
          const int e = tid >> 3, sub = tid & 7;
          double s = 0.0;
          if (e < 2 * kp) {
            const int i = e < kp ? e : e - kp;
            const double* src = (e < kp ? sWo : sVo) + i * TF_MAXS;
            for (int sl = s0 + sub; sl < S; sl += 8) s = fma(src[sl], svc[g + sl * G], s);
          }
          s += __shfl_xor_sync(0xffffffffu, s, 1);
          s += __shfl_xor_sync(0xffffffffu, s, 2);
          s += __shfl_xor_sync(0xffffffffu, s, 4);
          if (e < 2 * kp && sub == 0) PQs[(size_t)(e < kp ? e : PW + (e - kp)) * G + g] = s;  // [dot][cta]: coalesced for the readers
        }
        __syncthreads();
        if (tid == 0) {
          tf_fence_release();
          atomicAdd(ctr2, 1u);
        }
        ++nA;
      }
      TFB_TICK(0);
      // ================= eager columns (j, jend]: rank-2 update with reflector j-1, product with v_j ===========
      for (int slot = s0; slot < se; ++slot) {
        const int c = g + slot * G;
        const double wc = sw[c], vc = svn[c];
        double* col = A + (size_t)c * lda;
        double acc = 0.0;
        for (int p = tid; p < npairs; p += T) {
          const int r = r_lo + 2 * p;
          const double2 vp = *reinterpret_cast<const double2*>(svn + r);
          const double2 wp = *reinterpret_cast<const double2*>(sw + r);
          const double2 vn = *reinterpret_cast<const double2*>(svc + r);
          double2 xx = *reinterpret_cast<const double2*>(col + r);
          xx.x = fma(-vp.x, wc, fma(-wp.x, vc, xx.x));
          xx.y = fma(-vp.y, wc, fma(-wp.y, vc, xx.y));
          if (r + 1 >= n) { xx.y = 0.0; col[r] = xx.x; }
          else *reinterpret_cast<double2*>(col + r) = xx;
          if (c == first) *reinterpret_cast<double2*>(Xr + r) = xx;
          acc = fma(xx.x, vn.x, fma(xx.y, vn.y, acc));
        }
        acc = warp_sum(acc);
        if (lane == 0) spart[warp * TF_MAXS + slot] = acc;
      }
      if (boundary) {
        // ================= boundary sweep: rank-2PW update of all my columns fused with y = A v_j ================
        for (int sg = se; sg < S; sg += TF_CG) {
          double acc[TF_CG], wc[TF_CG], vc[TF_CG];
#pragma unroll
          for (int u = 0; u < TF_CG; ++u) {
            const int slot = sg + u;
            const int c = g + slot * G;
            acc[u] = 0.0;
            wc[u] = slot < S ? sw[c] : 0.0;
            vc[u] = slot < S ? svn[c] : 0.0;
          }
          for (int p = tid; p < npairs; p += T) {
            const int r = r_lo + 2 * p;
            const double2 vp = *reinterpret_cast<const double2*>(svn + r);  // v_{j-1}
            const double2 wp = *reinterpret_cast<const double2*>(sw + r);   // w_{j-1}
            const double2 vn = *reinterpret_cast<const double2*>(svc + r);  // v_j
            double2 x[TF_CG];
#pragma unroll
            for (int u = 0; u < TF_CG; ++u) {
              const int slot = sg + u;
              x[u] = make_double2(0.0, 0.0);
              if (slot < S) x[u] = *reinterpret_cast<const double2*>(A + (size_t)(g + slot * G) * lda + r);
            }
#pragma unroll 2
            for (int i = 0; i < kp - 1; ++i) {
              const double2 pv = __ldcg(reinterpret_cast<const double2*>(VP + (size_t)i * np + r));  // L2: written by other CTAs
              const double2 pw2 = __ldcg(reinterpret_cast<const double2*>(WP + (size_t)i * np + r));
              const double* wo = sWo + i * TF_MAXS + sg;
              const double* vo = sVo + i * TF_MAXS + sg;
#pragma unroll
              for (int u = 0; u < TF_CG; ++u) {
                const double wcu = wo[u], vcu = vo[u];  // zero beyond S (never written)
                x[u].x = fma(-pv.x, wcu, fma(-pw2.x, vcu, x[u].x));
                x[u].y = fma(-pv.y, wcu, fma(-pw2.y, vcu, x[u].y));
              }
            }
            const bool odd_tail = (r + 1 >= n);
#pragma unroll
            for (int u = 0; u < TF_CG; ++u) {
              const int slot = sg + u;
              if (slot < S) {
                double* col = A + (size_t)(g + slot * G) * lda;
                double2 xx = x[u];
                xx.x = fma(-vp.x, wc[u], fma(-wp.x, vc[u], xx.x));
                xx.y = fma(-vp.y, wc[u], fma(-wp.y, vc[u], xx.y));
                if (odd_tail) { xx.y = 0.0; col[r] = xx.x; }
                else *reinterpret_cast<double2*>(col + r) = xx;
                if (pub_raw && slot == s0) *reinterpret_cast<double2*>(Xr + r) = xx;
                acc[u] = fma(xx.x, vn.x, fma(xx.y, vn.y, acc[u]));
              }
            }
          }
          tfb_butterfly8(acc, lane, S - sg, spart + warp * TF_MAXS + sg);
        }
        j0 = j;
        TFB_TICK(1);
      } else {
        // ================= all other columns: read only ===========================================================
        // two row pairs per thread and pass: 16 independent 16-byte loads in flight per thread (the sweep is bound by
        // memory latency, not bandwidth: ncu long_scoreboard, profiles/r01_ncu_full_trd_fused.json "4096")
        for (int sg = se; sg < S; sg += TF_CG) {
          double acc[TF_CG];
#pragma unroll
          for (int u = 0; u < TF_CG; ++u) acc[u] = 0.0;
          const double* cb = A + (size_t)(g + sg * G) * lda + r_lo;
          const size_t cs = (size_t)G * lda;
          const int umax = S - 1 - sg;  // beyond S: re-read the last column, result discarded
          for (int p = tid; p < npairs; p += 2 * T) {
            const int p2 = p + T;
            const bool h2 = p2 < npairs;
            const int q2 = h2 ? p2 : p;
            double2 x0[TF_CG], x1[TF_CG];
#pragma unroll
            for (int u = 0; u < TF_CG; ++u) x0[u] = *reinterpret_cast<const double2*>(cb + min(u, umax) * cs + 2 * p);
#pragma unroll
            for (int u = 0; u < TF_CG; ++u) x1[u] = *reinterpret_cast<const double2*>(cb + min(u, umax) * cs + 2 * q2);
            double2 v0 = *reinterpret_cast<const double2*>(svc + r_lo + 2 * p);
            double2 v1 = *reinterpret_cast<const double2*>(svc + r_lo + 2 * q2);
            if (r_lo + 2 * p + 1 >= n) v0.y = 0.0;   // odd tail: the pad row is not part of the matrix
            if (r_lo + 2 * q2 + 1 >= n) v1.y = 0.0;
            if (!h2) { v1.x = 0.0; v1.y = 0.0; }
#pragma unroll
            for (int u = 0; u < TF_CG; ++u) {
              acc[u] = fma(x0[u].x, v0.x, fma(x0[u].y, v0.y, acc[u]));
              acc[u] = fma(x1[u].x, v1.x, fma(x1[u].y, v1.y, acc[u]));
            }
          }
          tfb_butterfly8(acc, lane, S - sg, spart + warp * TF_MAXS + sg);
        }
      }
      TFB_TICK(2);
      // ---- collect the partial dots ---------------------------------------------------------------------------
      if (exch_a) {
        if (tid == 0) wait_counter(ctr2, (unsigned)G * nA);
        __syncthreads();
        {  // every lane issues all its loads before the first sum (G <= 256: at most 8 per dot and lane)
          constexpr int EMAX = (2 * PW + NW - 1) / NW;
          double part[EMAX];
#pragma unroll
          for (int m = 0; m < EMAX; ++m) {
            const int e = warp + m * NW;
            const int d = e < kp ? e : PW + (e - kp);
            double s = 0.0;
            if (e < 2 * kp) {
              double v[8];
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const int gg = lane + 32 * u;
                v[u] = gg < G ? __ldcg(PQs + (size_t)d * G + gg) : 0.0;
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) s += v[u];
            }
            part[m] = s;
          }
#pragma unroll
          for (int m = 0; m < EMAX; ++m) {
            const int e = warp + m * NW;
            const double s = warp_sum(part[m]);
            if (lane == 0 && e < 2 * kp) sPQ[e < kp ? e : PW + (e - kp)] = s;
          }
        }
      }
      __syncthreads();
      TFB_TICK(3);
      {  // 16 lanes per column: lane 0 sums the per-warp partials, all lanes share the correction terms
        const int slot = s0 + (tid >> 4), sub = tid & 15;
        const bool act = slot < S;
        double y = 0.0;
        if (act && sub == 0) {
#pragma unroll
          for (int w = 0; w < NW; ++w) y += spart[w * TF_MAXS + slot];
        }
        if (exch_a) {
          double c0 = 0.0;
          if (act && slot >= se)
            for (int i = sub; i < kp; i += 16)
              c0 = fma(sVo[i * TF_MAXS + slot], sPQ[i], fma(sWo[i * TF_MAXS + slot], sPQ[PW + i], c0));
          c0 += __shfl_xor_sync(0xffffffffu, c0, 1);
          c0 += __shfl_xor_sync(0xffffffffu, c0, 2);
          c0 += __shfl_xor_sync(0xffffffffu, c0, 4);
          c0 += __shfl_xor_sync(0xffffffffu, c0, 8);
          y -= c0;
        }
        if (act && sub == 0) Xy[g + slot * G] = y;
      }
      // ================= exchange: arrive, wait for the other CTAs, load y and the raw column =================
      TFB_TICK(4);
      __syncthreads();
      if (tid == 0) {
        tf_fence_release();
        atomicAdd(ctr, 1u);
        wait_counter(ctr, (unsigned)G * (unsigned)(j + 1));
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int r = tid + k * T;
        const bool need = r >= first && r < n;
        yv[k] = need ? tf_ld(Xy + r) : 0.0;
        rv[k] = need ? tf_ld(Xr + r) : 0.0;
      }
    }
    TFB_TICK(5);
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
    double beta = 0.0, tau_n = 0.0, scale = 0.0;
    if (nf < n) {
      householder_scalars(sbc[1], xn2, beta, tau_n, scale);
      if (g == 0 && tid == 0) { (a.e + q * a.dstride)[first] = beta; (a.tau + q * a.dstride)[first] = tau_n; }
    }
    const int ip = j - j0;  // pending index of reflector j (>= 0 for j >= 0)
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int r = tid + k * T;
      if (r < n) {
        const double vnew = r < nf ? 0.0 : (r == nf ? 1.0 : av2[k] * scale);
        const double vcur = svc[r];
        svn[r] = vnew;
        sw[r] = wv[k];
        if (j >= 0) {
          if (own_slot[k] >= 0) {  // my column index: remember reflector j at it
            sVo[ip * TF_MAXS + own_slot[k]] = vcur;
            sWo[ip * TF_MAXS + own_slot[k]] = wv[k];
          }
          if (chunk_mine >> k & 1) {
            if (r > first) A[(size_t)j * lda + r] = vcur;  // LAPACK layout for the back-transformation (one step late)
            VP[(size_t)ip * np + r] = vcur;
            WP[(size_t)ip * np + r] = wv[k];
          }
        }
      }
    }
    __syncthreads();
    TFB_TICK(6);
    double* t = svc; svc = svn; svn = t;
    tau_j = tau_n;
  }
  if (jlast < n - 2) {
    // hand-over: apply the pending reflector jstop-1 (svn = v, sw = w after the last swap) to my columns >= jstop in place
    const int js = a.jstop;
    const int s0 = js > g ? (js - g + G - 1) / G : 0;
    for (int slot = s0; slot < S; ++slot) {
      const int c = g + slot * G;
      const double wc = sw[c], vc = svn[c];
      double* col = A + (size_t)c * lda;
      for (int r = js + tid; r < n; r += T) col[r] = fma(-svn[r], wc, fma(-sw[r], vc, col[r]));
    }
  }
#ifdef GSCF_TF_DEBUG_TIMING
  if ((tid == 0 || tid == T - 1) && (g == 0 || g == G - 1) && blockIdx.y == 0)
    for (int rg = 0; rg < 2; ++rg) {
      const int nc = max(1, tcols[rg]);
      const long long* t = tacc + 8 * rg;
      printf("cta %d tid %d %s (%d cols) cycles/col: pubA %lld boundary %lld sweep %lld waitA %lld yfin %lld exchB %lld post %lld loop %lld\n",
             g, tid, rg ? "unblocked" : "deferred", tcols[rg], t[0] / nc, t[1] / nc, t[2] / nc, t[3] / nc, t[4] / nc, t[5] / nc,
             t[6] / nc, t[7] / nc);
    }
#endif
  if (g == 0 && tid == 0 && *reinterpret_cast<volatile unsigned*>(a.abort_flag) != 0u)
    (a.d + q * a.dstride)[0] = __longlong_as_double(0x7ff8000000000000LL);
}

}  // namespace
}  // namespace gscfb
