// Divide & conquer eigensolver for symmetric tridiagonal matrices on the GPU (Cuppen's method with
// Gu-Eisenstat stabilisation - the dstedc algorithm family), batched over matrices and over the
// merges of one tree level.
//
// The tree is balanced: 2^L leaves with boundaries bnd(i) = floor(i n / 2^L); all merges of a
// level have (almost) the same size and run in the same launches.  Per level:
//   dc_prepare   one CTA per merge: z vector, rank sort, deflation (parallel block scans; the sequential dlaed2 logic
//                of dc_core.hpp only when a pair needs a Givens rotation), column typing, the rotations
//   dc_secular   one warp per root of the secular equation; writes delta(j,i) = d_j - lam_i
//   dc_zhat      one warp per pole: Gu-Eisenstat recomputation of z
//   dc_vectors   eigenvectors of the rank-one problem, rows grouped by column type
//   dc_gather    non-deflated columns -> Gq, deflated columns/eigenvalues -> output; GEMM descriptors
//   grouped DMMA GEMM  Q_top * U_top and Q_bot * U_bot with device-side sizes (after deflation)
// Leaves (order <= 32 in batches, <= 16 for a few large matrices) are solved by dc_leaf_tql2_kernel: implicit QL
// iteration, one warp per leaf.  T is scaled to unit max-norm first (dc_scale_kernel), like dstedc.
#include "stedc.cuh"

#include "dc_core.hpp"
#include "eig.cuh"

namespace gscfb {

namespace {

constexpr int kLeaf = 32;  // maximum leaf order

__host__ __device__ inline int bnd(int i, int n, int L) { return (int)(((long long)i * n) >> L); }

// Total order on doubles for the rank sorts (-NaN < -inf < ... < +inf < +NaN; -0 < +0): with a plain `<` a NaN in the
// spectrum (garbage input) would give two entries the same rank, leave a slot of the permutation unwritten and send
// the gather out of bounds.  For ordinary numbers the order is the usual one.
__device__ __forceinline__ long long dc_sort_key(double x) {
  const long long b = __double_as_longlong(x);
  return b < 0 ? (b ^ 0x7fffffffffffffffLL) : b;
}

struct NodeGeom {
  int lo, mid, hi;
};
__device__ inline NodeGeom node_geom(int t, int h, int n, int L) {
  NodeGeom g;
  g.lo = bnd(t << h, n, L);
  g.mid = bnd((t << h) + (1 << (h - 1)), n, L);
  g.hi = bnd((t + 1) << h, n, L);
  return g;
}

// ---- scaling ------------------------------------------------------------------------------------------
// dstedc scales T to unit max-norm first: the deflation tolerance of a merge, 8 eps max(|d|, |z|) with |z| <= 1, is an
// absolute quantity (unscaled, a matrix of norm 1e-3 would lose three digits and one of norm 1e-150 everything).
__global__ void __launch_bounds__(256)
dc_scale_kernel(double* __restrict__ d, double* __restrict__ e, long long dstride, StedcWs ws,
                const int* __restrict__ map) {
  __shared__ double red[8];
  const long long q = map ? map[blockIdx.x] : blockIdx.x;
  const int n = ws.n;
  double* dq = d + q * dstride;
  double* eq = e + q * dstride;
  double m = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    m = fmax(m, fabs(dq[i]));
    if (i < n - 1) m = fmax(m, fabs(eq[i]));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = 0.0;
  for (int w = 0; w < 8; ++w) m = fmax(m, red[w]);
  const double s = (m > 0.0 && isfinite(m)) ? m : 1.0;
  // entries below 1e-140 of the norm become exact zeros: the tridiagonalisation of a (numerically) low-rank matrix
  // leaves a cascade of ever smaller noise (1e-16, 1e-32, ... 1e-160) whose squares are subnormal, and the rotations
  // of the QL leaves computed from them are not orthogonal.  A zero e splits the problem instead.
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double dv = dq[i] / s;
    dq[i] = fabs(dv) < 1e-140 ? 0.0 : dv;   // NaN compares false: kept, flagged at the end
    if (i < n - 1) {
      const double ev = eq[i] / s;
      eq[i] = fabs(ev) < 1e-140 ? 0.0 : ev;
    }
  }
  if (threadIdx.x == 0) ws.scale[q] = s * (ws.outer ? ws.outer[q] : 1.0);
}

// ---- leaves ------------------------------------------------------------------------------------------
__global__ void dc_split_kernel(const double* __restrict__ d, const double* __restrict__ e, long long dstride,
                                StedcWs ws, const int* __restrict__ map) {
  const long long q = map ? map[blockIdx.y] : blockIdx.y;
  const int n = ws.n, L = ws.L;
  const double* dq = d + q * dstride;
  const double* eq = e + q * dstride;
  double* dm = ws.dmod + q * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double v = dq[i];
    // every leaf boundary b (0 < b < n) is a split point of exactly one merge: subtract |e_{b-1}|
    // from d[b-1] and d[b]
    // i == b - 1  <=> i + 1 is a boundary;  i == b <=> i is a boundary
    for (int s = 1; s < (1 << L); ++s) {
      const int b = bnd(s, n, L);
      if (b == i + 1 || b == i) v -= fabs(eq[b - 1]);
    }
    dm[i] = v;
  }
}

// Leaf solver: implicit-shift QL iteration (the EISPACK tql2 / LAPACK dsteqr recurrence) on the m <= 32 tridiagonal
// leaf, one WARP per leaf.  The scalar recurrence is uniform across the warp; lane k owns row k of the eigenvector
// matrix, so a plane rotation of two columns is one shared-memory read-modify-write per lane.  Writes the leaf's
// eigenvectors into the block diagonal of Z0 and its eigenvalues (any order) into lam[0].
constexpr int kLeafWarps = 8;
__global__ void __launch_bounds__(kLeafWarps * 32)
dc_leaf_tql2_kernel(const double* __restrict__ e, long long dstride, StedcWs ws, double* __restrict__ Z0, int ldz,
                    long long sZ, int nbatch, const int* __restrict__ map, int zs) {
  extern __shared__ __align__(16) double lsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long gl = (long long)blockIdx.x * kLeafWarps + warp;  // leaves of all matrices, flattened
  if (gl >= (long long)ws.nleaf * nbatch) return;
  const int zb = (int)(gl / ws.nleaf), leaf = (int)(gl - (long long)zb * ws.nleaf);
  const long long q = map ? map[zb] : zb;
  const int n = ws.n, L = ws.L;
  const int lo = bnd(leaf, n, L), hi = bnd(leaf + 1, n, L), m = hi - lo;
  // zs >= m: column stride of the leaf's eigenvector block = the largest leaf order of this launch, so that the shared
  // memory of a block follows the leaf size (leaves of 16: 18 KB instead of 70 KB per block - 8 instead of 3 blocks per SM;
  // 512 matrices of order 128 are 512 blocks, which were 1.15 waves of 444 slots)
  double* Zs = lsm + (size_t)warp * (zs * zs + 2 * zs);  // Zs[col * zs + row]
  double* sd = Zs + zs * zs;
  double* se = sd + zs;
  const double* dm = ws.dmod + q * n + lo;
  const double* eq = e + q * dstride + lo;
  if (lane < m)
    for (int c = 0; c < m; ++c) Zs[c * zs + lane] = (c == lane) ? 1.0 : 0.0;
  if (lane < m) {
    sd[lane] = dm[lane];
    se[lane] = lane < m - 1 ? eq[lane] : 0.0;
  }
  __syncwarp();
  const double eps = 2.220446049250313e-16;
  for (int l = 0; l < m; ++l) {
    for (int iter = 0; iter < 60; ++iter) {
      // first negligible off-diagonal entry at or after l: every lane tests its own entry, one ballot (the serial scan
      // was a chain of up to m dependent shared-memory loads per QL iteration)
      const bool small = lane < m - 1 && fabs(se[lane]) <= eps * (fabs(sd[lane]) + fabs(sd[lane + 1]));
      const unsigned cand = __ballot_sync(0xffffffffu, small) & (0xffffffffu << l);
      const int mm = cand ? min(m - 1, __ffs(cand) - 1) : m - 1;
      if (mm == l) break;
      const double dl = sd[l], el = se[l];
      double g = (sd[l + 1] - dl) / (2.0 * el);
      double r = hypot(g, 1.0);
      g = sd[mm] - dl + el / (g + copysign(r, g));
      __syncwarp();
      double sn = 1.0, cs = 1.0, p = 0.0;
      int i = mm - 1;
      bool underflow = false;
      // the entries of step i are loaded one step ahead (d[i + 1] of the next step is this step's d[i]): their
      // shared-memory latency is off the recurrence; a sweep only writes entries it has already read
      double ei = se[i], di1 = sd[i + 1], di = sd[i];
      for (; i >= l; --i) {
        double ein = 0.0, din = 0.0;
        if (i > l) { ein = se[i - 1]; din = sd[i - 1]; }
        __syncwarp();  // every lane has read this step's entries before lane 0 overwrites them
        const double f = sn * ei, b = cs * ei;
        // r = sqrt(x), 1 / r as ONE reciprocal square root (1 ulp) and a product: the sqrt + divide chain was a third of
        // the ~440 cycles of a rotation step, and a leaf is nothing but ~1.5 m^2 such steps in sequence.  Entries are
        // bounded by the leaf's norm; an underflow of x to 0 takes the exit below.
        const double x = fma(f, f, g * g);
        const double rinv = rsqrt(x);
        r = x * rinv;
        if (x == 0.0) r = 0.0;
        if (lane == 0) se[i + 1] = r;
        if (r == 0.0) {
          if (lane == 0) { sd[i + 1] = di1 - p; se[mm] = 0.0; }
          underflow = true;
          break;
        }
        sn = f * rinv;
        cs = g * rinv;
        g = di1 - p;
        r = (di - g) * sn + 2.0 * cs * b;
        p = sn * r;
        if (lane == 0) sd[i + 1] = g + p;
        g = cs * r - b;
        if (lane < m) {
          const double z1 = Zs[(i + 1) * zs + lane], z0 = Zs[i * zs + lane];
          Zs[(i + 1) * zs + lane] = sn * z0 + cs * z1;
          Zs[i * zs + lane] = cs * z0 - sn * z1;
        }
        di1 = di; di = din; ei = ein;
      }
      __syncwarp();
      if (!underflow) {
        if (lane == 0) { sd[l] = dl - p; se[l] = g; se[mm] = 0.0; }
      }
      __syncwarp();
    }
  }
  __syncwarp();
  double* Zq = Z0 + q * sZ;
  double* lam = ws.lam[0] + q * n;
  for (int c = 0; c < m; ++c)
    if (lane < m) Zq[(size_t)(lo + c) * ldz + lo + lane] = Zs[c * zs + lane];
  if (lane < m) lam[lo + lane] = sd[lane];
}

__global__ void zero_matrix_kernel(double* __restrict__ Z, int ldz, long long sZ, int n,
                                   const int* __restrict__ map) {
  const long long q = map ? map[blockIdx.y] : blockIdx.y;
  double* Zq = Z + q * sZ;
  const long long total = (long long)n * n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / n), r = (int)(i - (long long)c * n);
    Zq[(size_t)c * ldz + r] = 0.0;
  }
}

// ---- merge: prepare / deflate ------------------------------------------------------------------------
// dynamic shared memory: k * (4 doubles + 4 ints) when it fits, otherwise the global work arrays
__global__ void __launch_bounds__(1024)
dc_prepare_kernel(StedcWs ws, const double* __restrict__ e, long long dstride, int h, int lev_in,
                  double* __restrict__ Zin, int ldin, long long sin, int use_smem,
                  const int* __restrict__ map) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int t = blockIdx.x, zb = blockIdx.y;
  const long long q = map ? map[zb] : zb;
  const int n = ws.n, L = ws.L;
  const NodeGeom g = node_geom(t, h, n, L);
  const int lo = g.lo, n1 = g.mid - g.lo, k = g.hi - g.lo;
  double* Zq = Zin + q * sin;
  const double* lam_in = ws.lam[lev_in] + q * n + lo;
  double *sd, *sz, *sdl, *sw;
  int *sidx, *sct, *snd, *sdc;
  if (use_smem) {
    sd = reinterpret_cast<double*>(smraw);
    sz = sd + k; sdl = sz + k; sw = sdl + k;
    sidx = reinterpret_cast<int*>(sw + k);
    sct = sidx + k; snd = sct + k; sdc = snd + k;
  } else {
    sd = ws.dwork + q * n + lo; sz = ws.z + q * n + lo; sdl = ws.dlamda + q * n + lo; sw = ws.w + q * n + lo;
    sidx = ws.idx + q * n + lo; sct = ws.ctype + q * n + lo; snd = ws.ndcol + q * n + lo; sdc = ws.dcol + q * n + lo;
  }
  long long* skey = reinterpret_cast<long long*>(sw);  // sort keys; sw (pole weights) is only written by the deflation
  const double ecoup = (e + q * dstride)[g.mid - 1];
  const double sgn = ecoup < 0.0 ? -1.0 : 1.0;
  const double rho = 2.0 * fabs(ecoup);
  const double rs2 = 0.70710678118654752440;  // 1/sqrt(2)
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    sd[j] = lam_in[j];
    const double zz = j < n1 ? Zq[(size_t)(lo + j) * ldin + (g.mid - 1)] : sgn * Zq[(size_t)(lo + j) * ldin + g.mid];
    sz[j] = zz * rs2;
    sct[j] = j < n1 ? 1 : 3;
    skey[j] = dc_sort_key(lam_in[j]);
  }
  __syncthreads();
  // rank sort (ties by index) -> idx
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const long long dj = skey[j];
    int r = 0;
    for (int i = 0; i < k; ++i) {
      const long long di = skey[i];
      r += (di < dj || (di == dj && i < j)) ? 1 : 0;
    }
    sidx[r] = j;
  }
  __syncthreads();
  __shared__ int s_K, s_nrot;
  int* meta = ws.meta + ((size_t)q * ws.nleaf + t) * 8;
  int* gpos = ws.gpos + q * n + lo;
  int* rot_a = ws.rot_a + q * n + lo;
  int* rot_b = ws.rot_b + q * n + lo;
  double* rot_c = ws.rot_c + q * n + lo;
  double* rot_s = ws.rot_s + q * n + lo;
  // ---- deflation.  Fast path (all threads): entries with a negligible z component deflate, the others become
  // poles in sorted order, provided no pair of neighbouring poles is close enough for a Givens rotation
  // (dlaed2's second test).  A rotation changes z and d of the pair, which feeds the test of the next pair,
  // so as soon as one is needed thread 0 redoes the merge with the sequential dc_deflate.
  __shared__ double s_red[32];
  __shared__ int s_scan[32];
  __shared__ int s_any;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nwarp = blockDim.x >> 5;
  auto block_max = [&](double v) {
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) s_red[wid] = v;
    __syncthreads();
    double m = 0.0;
    for (int w = 0; w < nwarp; ++w) m = fmax(m, s_red[w]);
    return m;
  };
  // exclusive scan of one int per thread over the block; *total = sum
  auto block_scan = [&](int v, int* total) {
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_scan[wid] = inc;
    __syncthreads();
    int base = 0, tot = 0;
    for (int w = 0; w < nwarp; ++w) {
      const int c = s_scan[w];
      if (w < wid) base += c;
      tot += c;
    }
    *total = tot;
    return base + inc - v;
  };
  double dmx = 0.0, zmx = 0.0;
  for (int j = tid; j < k; j += blockDim.x) { dmx = fmax(dmx, fabs(sd[j])); zmx = fmax(zmx, fabs(sz[j])); }
  dmx = block_max(dmx);
  zmx = block_max(zmx);
  const double tol = 8.0 * kDcEps * fmax(dmx, zmx);
  if (tid == 0) s_any = 0;
  // contiguous chunk of sorted positions per thread
  const int per = (k + blockDim.x - 1) / blockDim.x;
  const int j_lo = min(k, tid * per), j_hi = min(k, j_lo + per);
  int cnt = 0;
  for (int jj = j_lo; jj < j_hi; ++jj) cnt += (rho * fabs(sz[sidx[jj]]) > tol) ? 1 : 0;
  int Kp = 0;
  int pos = block_scan(cnt, &Kp);
  const bool all_deflated = rho * zmx <= tol;
  if (all_deflated) Kp = 0;
  {
    int ppos = pos;
    for (int jj = j_lo; jj < j_hi; ++jj) {
      const int nj = sidx[jj];
      if (!all_deflated && rho * fabs(sz[nj]) > tol) snd[ppos++] = nj;
      else sdc[jj - (all_deflated ? 0 : ppos)] = nj;
    }
  }
  __syncthreads();
  for (int pp = 1 + tid; pp < Kp; pp += blockDim.x) {
    const int pj = snd[pp - 1], nj = snd[pp];
    double sv = sz[pj], cv = sz[nj];
    const double tau = hypot(cv, sv);
    const double t = sd[nj] - sd[pj];
    cv /= tau;
    sv = -sv / tau;
    if (fabs(t * cv * sv) <= tol) s_any = 1;
  }
  __syncthreads();
  if (s_any) {
    if (threadIdx.x == 0) {
      int K = 0, nrot = 0;
      dc_deflate(k, sd, sz, rho, sidx, sct, &K, sdl, sw, snd, sdc, &nrot, rot_a, rot_b, rot_c, rot_s);
      // group the poles by column type: [1 | 2 | 3]
      int c1 = 0, c2 = 0, c3 = 0;
      for (int p = 0; p < K; ++p) {
        const int ty = sct[snd[p]];
        c1 += ty == 1; c2 += ty == 2; c3 += ty == 3;
      }
      int o1 = 0, o2 = c1, o3 = c1 + c2;
      for (int p = 0; p < K; ++p) {
        const int ty = sct[snd[p]];
        gpos[p] = ty == 1 ? o1++ : (ty == 2 ? o2++ : o3++);
      }
      s_K = K; s_nrot = nrot;
      meta[0] = K; meta[1] = nrot; meta[2] = c1; meta[3] = c2; meta[4] = c3;
      (ws.rho + (size_t)q * ws.nleaf)[t] = rho;
    }
  } else {
    // poles in sorted order; stable partition by column type (1: top rows, 3: bottom rows; no rotations -> no type 2)
    const int pper = (Kp + blockDim.x - 1) / blockDim.x;
    const int p_lo = min(Kp, tid * pper), p_hi = min(Kp, p_lo + pper);
    int c1l = 0;
    for (int pp = p_lo; pp < p_hi; ++pp) {
      const int nj = snd[pp];
      sdl[pp] = sd[nj];
      sw[pp] = sz[nj];
      c1l += sct[nj] == 1 ? 1 : 0;
    }
    int c1 = 0;
    int o1 = block_scan(c1l, &c1);
    int o3 = c1 + (p_lo - o1);
    for (int pp = p_lo; pp < p_hi; ++pp) gpos[pp] = sct[snd[pp]] == 1 ? o1++ : o3++;
    if (tid == 0) {
      s_K = Kp; s_nrot = 0;
      meta[0] = Kp; meta[1] = 0; meta[2] = c1; meta[3] = 0; meta[4] = Kp - c1;
      (ws.rho + (size_t)q * ws.nleaf)[t] = rho;
    }
  }
  __syncthreads();
  const int K = s_K, nrot = s_nrot;
  if (use_smem) {
    double* gd = ws.dwork + q * n + lo;
    double* gdl = ws.dlamda + q * n + lo;
    double* gw = ws.w + q * n + lo;
    int* gnd = ws.ndcol + q * n + lo;
    int* gdc = ws.dcol + q * n + lo;
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      gd[j] = sd[j];
      if (j < K) { gdl[j] = sdl[j]; gw[j] = sw[j]; gnd[j] = snd[j]; }
      if (j < k - K) gdc[j] = sdc[j];
    }
  }
  // Givens rotations of the deflated near-degenerate pairs, in order
  for (int r = 0; r < nrot; ++r) {
    const int a = rot_a[r], b = rot_b[r];
    const double c = rot_c[r], s = rot_s[r];
    double* x = Zq + (size_t)(lo + a) * ldin + lo;
    double* y = Zq + (size_t)(lo + b) * ldin + lo;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      const double xv = x[i], yv = y[i];
      x[i] = c * xv + s * yv;
      y[i] = c * yv - s * xv;
    }
    __syncthreads();
  }
}

// ---- merge: secular equation --------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dc_secular_kernel(StedcWs ws, int h, int lev_out, const int* __restrict__ map) {
  const int t = blockIdx.y, zb = blockIdx.z;
  const long long q = map ? map[zb] : zb;
  const int n = ws.n;
  const NodeGeom g = node_geom(t, h, n, ws.L);
  const int lo = g.lo;
  const int K = (ws.meta + ((size_t)q * ws.nleaf + t) * 8)[0];
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // one warp per root
  if (i >= K) return;
  const double rho = (ws.rho + (size_t)q * ws.nleaf)[t];
  double* S = ws.S + q * (long long)ws.ld * n + (size_t)lo * ws.ld + lo;  // node block, [j*ld + i]
  const double lam = secular_root_warp(K, i, ws.dlamda + q * n + lo, ws.w + q * n + lo, rho, S + i, ws.ld);
  if ((threadIdx.x & 31) == 0) (ws.lam[lev_out] + q * n + lo)[i] = lam;
}

// One THREAD per root, for batches of many small merges (C4: 4096 x n = 128): with K <= 128 poles a warp per root spends
// most of its instructions on the four butterfly sums and on scalar work that 32 lanes repeat (measured: the two
// dc_secular launches of a C4 iteration were 2.6 of 9 ms of its D&C stage).  Poles and z live in shared memory (every lane
// reads the same entry: broadcast), consecutive threads write consecutive rows of delta.  Same iteration and stopping
// rule (dc_core.hpp:secular_root, the routine the host unit test checks against dlaed4-style references).
__global__ void __launch_bounds__(128)
dc_secular_thread_kernel(StedcWs ws, int h, int lev_out, int kcap, const int* __restrict__ map) {
  extern __shared__ __align__(16) double ssec[];  // d[kcap] | z[kcap]
  const int t = blockIdx.y, zb = blockIdx.z;
  const long long q = map ? map[zb] : zb;
  const int n = ws.n;
  const NodeGeom g = node_geom(t, h, n, ws.L);
  const int lo = g.lo;
  const int K = (ws.meta + ((size_t)q * ws.nleaf + t) * 8)[0];
  if ((int)(blockIdx.x * blockDim.x) >= K) return;  // whole block beyond the (deflated) problem size
  double* sd = ssec;
  double* sz = ssec + kcap;
  const double* dl = ws.dlamda + q * n + lo;
  const double* wl = ws.w + q * n + lo;
  for (int j = threadIdx.x; j < K; j += blockDim.x) { sd[j] = dl[j]; sz[j] = wl[j]; }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  const double rho = (ws.rho + (size_t)q * ws.nleaf)[t];
  double* S = ws.S + q * (long long)ws.ld * n + (size_t)lo * ws.ld + lo;  // node block, [j*ld + i]
  const double lam = secular_root(K, i, sd, sz, rho, S + i, ws.ld, nullptr);
  (ws.lam[lev_out] + q * n + lo)[i] = lam;
}

// zhat_j = sign(w_j) sqrt( - delta(j,j) * prod_{i != j} delta(j,i) / (d_j - d_i) )
__global__ void __launch_bounds__(256)
dc_zhat_kernel(StedcWs ws, int h, const int* __restrict__ map) {
  const int t = blockIdx.y, zb = blockIdx.z;
  const long long q = map ? map[zb] : zb;
  const int n = ws.n;
  const NodeGeom g = node_geom(t, h, n, ws.L);
  const int lo = g.lo;
  const int K = (ws.meta + ((size_t)q * ws.nleaf + t) * 8)[0];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * (blockDim.x >> 5) + warp;
  if (j >= K) return;
  const double* S = ws.S + q * (long long)ws.ld * n + (size_t)lo * ws.ld + lo;
  const double* dl = ws.dlamda + q * n + lo;
  const double dj = dl[j];
  // four factors per division (the kernel was bound by its K^2 FP64 divisions: 0.6 of 6.7 ms of a C4 divide & conquer): T
  // has unit max-norm and deflation keeps |d_j - d_i| and |delta| above ~1e-30, so products of four cannot leave the range
  double p = 1.0;
  for (int i0 = lane; i0 < K; i0 += 128) {
    double num = 1.0, den = 1.0;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + 32 * u;
      if (i < K) {
        num *= S[(size_t)j * ws.ld + i];
        if (i != j) den *= dj - dl[i];
      }
    }
    p *= num / den;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) p *= __shfl_xor_sync(0xffffffffu, p, o);
  if (lane == 0) (ws.zhat + q * n + lo)[j] = copysign(sqrt(fabs(p)), (ws.w + q * n + lo)[j]);
}

// U^T[gpos[j]][i] = zhat_j / delta(j,i) / ||.||.  CTA = 32 eigenvectors (lane) x 8 warps splitting the poles.
__global__ void __launch_bounds__(1024)
dc_vectors_kernel(StedcWs ws, int h, int partial, const int* __restrict__ map) {
  __shared__ double snrm[32][33];
  const int nw = blockDim.x >> 5;  // 8 warps, or 32 when a level has few large merges (K / 32 CTAs would not fill the GPU)
  const int t = blockIdx.y, zb = blockIdx.z;
  const long long q = map ? map[zb] : zb;
  const int n = ws.n;
  const NodeGeom g = node_geom(t, h, n, ws.L);
  const int lo = g.lo;
  const int K = (ws.meta + ((size_t)q * ws.nleaf + t) * 8)[0];
  // partial spectrum (top-level merge only): eigenvectors of the first Ksel roots, see dc_select_kernel
  const int Kv = partial ? (ws.meta + ((size_t)q * ws.nleaf + t) * 8)[5] : K;
  if (blockIdx.x * 32 >= Kv) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const double* S = ws.S + q * (long long)ws.ld * n + (size_t)lo * ws.ld + lo;
  double* U = ws.U + q * (long long)ws.ld * n + (size_t)lo * ws.ld + lo;
  const double* zh = ws.zhat + q * n + lo;
  const int* gpos = ws.gpos + q * n + lo;
  // pass 1: unnormalised components (one division each) into U, norms; pass 2: scale in place (same thread, same entries)
  double nrm = 0.0;
  if (i < Kv)
    for (int j = warp; j < K; j += nw) {
      const double u = zh[j] / S[(size_t)j * ws.ld + i];
      U[(size_t)gpos[j] * ws.ld + i] = u;
      nrm = fma(u, u, nrm);
    }
  snrm[warp][lane] = nrm;
  __syncthreads();
  double tot = 0.0;
  for (int w = 0; w < nw; ++w) tot += snrm[w][lane];
  const double inv = 1.0 / sqrt(tot);
  if (i < Kv)
    for (int j = warp; j < K; j += nw) U[(size_t)gpos[j] * ws.ld + i] *= inv;
}

// Partial spectrum (n_eigenpairs = nvec < n, ScfBase.hh:87,315-316): at the TOP-LEVEL merge only the eigenvectors of
// the nvec lowest eigenvalues of the whole matrix are wanted.  All K roots are still needed (the Gu-Eisenstat z-hat is a
// product over every root), but eigenvectors - the K x K matrix U and the Q U products, three quarters of the flops of
// the whole divide & conquer - only for the roots whose global rank is below nvec.  Roots ascend with their index, so
// these are the first Ksel roots; the rank counts the deflated eigenvalues below a root (ties: the root first, the order
// dc_rank_kernel uses).  Ksel goes to meta[5] of the (single) top-level node.
__global__ void __launch_bounds__(256)
dc_select_kernel(StedcWs ws, int lev_out, int nvec, const int* __restrict__ map) {
  __shared__ int s_cnt;
  const long long q = map ? map[blockIdx.x] : blockIdx.x;
  const int n = ws.n;
  int* meta = ws.meta + (size_t)q * ws.nleaf * 8;  // node 0 of the top level
  const int K = meta[0];
  const double* lam = ws.lam[lev_out] + q * n;
  const int* dcol = ws.dcol + q * n;
  const double* dwork = ws.dwork + q * n;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int mine = 0;
  for (int i = threadIdx.x; i < K; i += blockDim.x) {
    const long long li = dc_sort_key(lam[i]);
    int below = 0;
    for (int tdx = 0; tdx < n - K; ++tdx) below += dc_sort_key(dwork[dcol[tdx]]) < li ? 1 : 0;
    if (i + below < nvec) ++mine;
  }
  if (mine) atomicAdd(&s_cnt, mine);
  __syncthreads();
  if (threadIdx.x == 0) meta[5] = s_cnt;
}

// Gather non-deflated columns (grouped by type) into Gq, copy deflated columns/eigenvalues to the
// output, and write the two GEMM descriptors of this merge.
__global__ void __launch_bounds__(128)
dc_gather_kernel(StedcWs ws, int h, int lev_out, const double* __restrict__ Zin, int ldin, long long sin,
                 double* __restrict__ Zout, int ldout, long long sout, int nodes, int partial, int cpb,
                 const int* __restrict__ map) {
  const int t = blockIdx.y, zb = blockIdx.z;
  const long long q = map ? map[zb] : zb;
  const int n = ws.n;
  const NodeGeom g = node_geom(t, h, n, ws.L);
  const int lo = g.lo, n1 = g.mid - g.lo, k = g.hi - g.lo;
  const int* meta = ws.meta + ((size_t)q * ws.nleaf + t) * 8;
  const int K = meta[0];
  const double* Zq = Zin + q * sin;
  double* Zo = Zout + q * sout;
  double* Gq = ws.Gq + q * (long long)ws.ld * n + (size_t)lo * ws.ld + lo;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const int c1 = meta[2], c2 = meta[3], c3 = meta[4];
    GemmDesc* dsc = ws.descs + ((size_t)zb * nodes + t) * 2;
    const double* U = ws.U + q * (long long)ws.ld * n + (size_t)lo * ws.ld + lo;
    GemmDesc top, bot;
    top.A = Gq; top.lda = ws.ld; top.B = U; top.ldb = ws.ld;
    top.C = Zo + (size_t)lo * ldout + lo; top.ldc = ldout;
    const int Kn = partial ? meta[5] : K;  // partial spectrum: only the columns of the selected roots
    top.m = n1; top.n = Kn; top.k = c1 + c2; top.alpha = 1.0; top.beta = 0.0;
    bot.A = Gq + (size_t)c1 * ws.ld + n1; bot.lda = ws.ld; bot.B = U + (size_t)c1 * ws.ld; bot.ldb = ws.ld;
    bot.C = Zo + (size_t)lo * ldout + lo + n1; bot.ldc = ldout;
    bot.m = k - n1; bot.n = Kn; bot.k = c2 + c3; bot.alpha = 1.0; bot.beta = 0.0;
    dsc[0] = top;
    dsc[1] = bot;
  }
  // cpb == 1: one column per block (large merges: the whole block streams one long column); cpb > 1: one column per
  // WARP, cpb columns per block - with short columns (batches of small merges) a block per column was bound by the
  // block scheduling rate (4096 x 129 blocks of 1 KB each: 0.35 - 0.4 ms per launch at n = 128)
  const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c_begin = cpb == 1 ? blockIdx.x : blockIdx.x * cpb + warp;
  const int c_end = cpb == 1 ? blockIdx.x + 1 : min(k, (int)(blockIdx.x + 1) * cpb);
  const int c_step = cpb == 1 ? 1 : nwarp;
  const int i0 = cpb == 1 ? threadIdx.x : lane, istep = cpb == 1 ? blockDim.x : 32;
  for (int c = c_begin; c < c_end && c < k; c += c_step) {
    if (c < K) {
      const int src = (ws.ndcol + q * n + lo)[c];
      const int dst = (ws.gpos + q * n + lo)[c];
      const double* x = Zq + (size_t)(lo + src) * ldin + lo;
      double* y = Gq + (size_t)dst * ws.ld;
      for (int i = i0; i < k; i += istep) y[i] = x[i];
    } else {
      const int tdx = c - K;
      const int src = (ws.dcol + q * n + lo)[tdx];
      const double* x = Zq + (size_t)(lo + src) * ldin + lo;
      double* y = Zo + (size_t)(lo + c) * ldout + lo;
      for (int i = i0; i < k; i += istep) y[i] = x[i];
      if (i0 == 0) (ws.lam[lev_out] + q * n + lo)[c] = (ws.dwork + q * n + lo)[src];
    }
  }
}

// ---- final ascending sort -----------------------------------------------------------------------------
// wpj == 0: one thread per entry (batches of small matrices); wpj == 1: one WARP per entry, the lanes split the n
// comparisons (one large matrix: 1024 threads walking 1024 entries each took 54 us of a 5.7 ms iteration at n = 1024)
__global__ void dc_rank_kernel(StedcWs ws, int lev, int* __restrict__ info, int wpj, const int* __restrict__ map) {
  const long long q = map ? map[blockIdx.y] : blockIdx.y;
  const int n = ws.n;
  const double* lam = ws.lam[lev] + q * n;
  int* rank = ws.rank + q * n;
  if (wpj) {
    const int lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
    for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n; j += nw) {
      const long long lj = dc_sort_key(lam[j]);
      int r = 0;
      for (int i = lane; i < n; i += 32) {
        const long long li = dc_sort_key(lam[i]);
        r += (li < lj || (li == lj && i < j)) ? 1 : 0;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
      if (lane == 0) rank[j] = r;
    }
  } else {
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
      const long long lj = dc_sort_key(lam[j]);
      int r = 0;
      for (int i = 0; i < n; ++i) {
        const long long li = dc_sort_key(lam[i]);
        r += (li < lj || (li == lj && i < j)) ? 1 : 0;
      }
      rank[j] = r;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && info) info[q] = 0;
}

__global__ void dc_permute_kernel(StedcWs ws, int lev, const double* __restrict__ Zin, int ldin, long long sin,
                                  double* __restrict__ Zout, int ldout, long long sout, double* __restrict__ w,
                                  long long stride_w, int* __restrict__ info, int nvec, int cpb,
                                  const int* __restrict__ map) {
  const long long q = map ? map[blockIdx.y] : blockIdx.y;
  const int n = ws.n;
  // one column per block, or (short columns: batches of small matrices) one per warp and cpb per block - see dc_gather
  const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j_begin = cpb == 1 ? blockIdx.x : blockIdx.x * cpb + warp;
  const int j_end = cpb == 1 ? blockIdx.x + 1 : min(n, (int)(blockIdx.x + 1) * cpb);
  const int j_step = cpb == 1 ? 1 : nwarp;
  const int i0 = cpb == 1 ? threadIdx.x : lane, istep = cpb == 1 ? blockDim.x : 32;
  for (int j = j_begin; j < j_end && j < n; j += j_step) {
    const int r = (ws.rank + q * n)[j];
    const double* x = Zin + q * sin + (size_t)j * ldin;
    double* y = Zout + q * sout + (size_t)r * ldout;
    if (r < nvec)  // partial spectrum: the other eigenvectors were never formed
      for (int i = i0; i < n; i += istep) y[i] = x[i];
    if (i0 == 0) {
      const double l = (ws.lam[lev] + q * n)[j] * ws.scale[q];
      (w + q * stride_w)[r] = l;
      if (!isfinite(l) && info) info[q] = 1;
    }
  }
}

}  // namespace

size_t stedc_carve(char* base, size_t offset, int n, int cap, StedcWs& ws) {
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = up(offset);
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off = up(off + bytes);
    return p;
  };
  ws.n = n; ws.ld = padded_ld(n); ws.cap = cap;
  int L = 0;
  // leaves: 32 for batches (throughput: fewer merge levels), 16 for a few large matrices (latency: the QL iteration of
  // a leaf is a serial chain of ~m^2 rotations; measured n = 1024: 5.90 ms with 32, 5.64 ms with 16)
  static int leaf_env = -1;
  if (leaf_env < 0) { const char* le = getenv("GSCF_B200_DC_LEAF"); leaf_env = le ? std::max(4, std::min(kLeaf, atoi(le))) : 0; }
  // (batches of small matrices: measured at n = 128, whole D&C stage per SCF solve - 4096 matrices 143.5 ms with leaves of 32
  // against 156.9 ms with 16, 512 matrices 25.3 against 23.0 ms: below ~1000 matrices the serial QL chain of a 32-leaf is
  // no longer hidden behind other leaves)
  const int leaf_target = leaf_env > 0 ? leaf_env : ((cap <= 8 || (cap <= 1024 && n <= 256)) ? 16 : kLeaf);
  while (ceil_div(n, 1 << L) > leaf_target) ++L;
  ws.L = L; ws.nleaf = 1 << L;
  ws.lmax = ceil_div(n, 1 << L);
  const size_t cn = (size_t)cap * n;
  ws.lam[0] = (double*)take(cn * 8); ws.lam[1] = (double*)take(cn * 8);
  ws.dmod = (double*)take(cn * 8); ws.dwork = (double*)take(cn * 8); ws.z = (double*)take(cn * 8);
  ws.dlamda = (double*)take(cn * 8); ws.w = (double*)take(cn * 8); ws.zhat = (double*)take(cn * 8);
  ws.rot_c = (double*)take(cn * 8); ws.rot_s = (double*)take(cn * 8);
  ws.rho = (double*)take((size_t)cap * ws.nleaf * 8);
  ws.scale = (double*)take((size_t)cap * 8);
  ws.idx = (int*)take(cn * 4); ws.ctype = (int*)take(cn * 4); ws.ndcol = (int*)take(cn * 4);
  ws.dcol = (int*)take(cn * 4); ws.gpos = (int*)take(cn * 4); ws.rot_a = (int*)take(cn * 4);
  ws.rot_b = (int*)take(cn * 4); ws.rank = (int*)take(cn * 4);
  ws.meta = (int*)take((size_t)cap * ws.nleaf * 8 * 4);
  const size_t mat = (size_t)cap * ws.ld * n * 8;
  if (L > 0) {
    ws.Zb = (double*)take(mat); ws.S = (double*)take(mat); ws.U = (double*)take(mat); ws.Gq = (double*)take(mat);
  } else {
    ws.Zb = ws.S = ws.U = ws.Gq = nullptr;
    ws.Zb = (double*)take(mat);
  }
  ws.descs = (GemmDesc*)take((size_t)cap * ws.nleaf * sizeof(GemmDesc));
  return off;
}

int stedc_batched(int n, double* d, double* e, long long dstride, double* w, long long stride_w,
                  double* Z, int ldz, long long strideZ, const StedcWs& ws, int batch, const int* map,
                  int* info_dev, cudaStream_t st, int nvec) {
  if (batch <= 0) return GSCF_B200_OK;
  const int L = ws.L;
  const int nv = (nvec > 0 && nvec < n) ? nvec : n;
  const int gx = std::max(1, std::min(148 * 4 / std::max(1, batch), ceil_div(n * n, 256 * 8)));
  // buffers: X_0 (leaf eigenvectors) -> level 1 -> ... -> X_L -> sorted into Z
  // X_L must be the workspace buffer, so X_0 = Zb for even L and Z for odd L.
  double* bufp[2]; int bufld[2]; long long bufs[2];
  const int first = (L % 2 == 0) ? 1 : 0;  // index 0 = user Z, 1 = Zb
  bufp[0] = Z; bufld[0] = ldz; bufs[0] = strideZ;
  bufp[1] = ws.Zb; bufld[1] = ws.ld; bufs[1] = (long long)ws.ld * n;
  int cur = first;
  dc_scale_kernel<<<batch, 256, 0, st>>>(d, e, dstride, ws, map);
  count_launch(1);
  // ---- leaves -----------------------------------------------------------------------------------------
  dc_split_kernel<<<dim3(std::max(1, ceil_div(n, 256)), batch), 256, 0, st>>>(d, e, dstride, ws, map);
  count_launch(1);
  GSCF_CHECK_LAUNCH();
  // GSCF_B200_DC_LEAF_SMEM=full: the 32-column layout whatever the leaf size (A/B)
  static int leaf_full = -1;
  if (leaf_full < 0) { const char* lf = getenv("GSCF_B200_DC_LEAF_SMEM"); leaf_full = (lf && lf[0] == 'f') ? 1 : 0; }
  const int zs = leaf_full ? kLeaf : std::max(1, std::min(kLeaf, ws.lmax));
  const size_t leaf_smem = (size_t)kLeafWarps * (zs * zs + 2 * zs) * sizeof(double);
  const size_t leaf_smem_max = (size_t)kLeafWarps * (kLeaf * kLeaf + 2 * kLeaf) * sizeof(double);
  static PerDeviceOnce once;
  int dev;
  const bool configure = once.need(dev);
  if (configure) {
    GSCF_CUDA_TRY(cudaFuncSetAttribute(dc_leaf_tql2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)leaf_smem_max));
    GSCF_CUDA_TRY(cudaFuncSetAttribute(dc_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    once.done(dev);
  }
  zero_matrix_kernel<<<dim3(gx, batch), 256, 0, st>>>(bufp[cur], bufld[cur], bufs[cur], n, map);
  dc_leaf_tql2_kernel<<<(unsigned)ceil_div_ll((long long)ws.nleaf * batch, kLeafWarps), kLeafWarps * 32, leaf_smem, st>>>(
      e, dstride, ws, bufp[cur], bufld[cur], bufs[cur], batch, map, zs);
  count_launch(2);
  GSCF_CHECK_LAUNCH();
  int lev = 0;
  // ---- merges -----------------------------------------------------------------------------------------
  for (int h = 1; h <= L; ++h) {
    const int nodes = 1 << (L - h);
    const int kmax = ceil_div(n, nodes) + 1;
    const int n1max = ceil_div(kmax, 2) + 1;
    const int nxt = 1 - cur;
    const size_t smem = (size_t)kmax * 48;
    const int use_smem = smem <= 200 * 1024 ? 1 : 0;
    // the top-level merge writes every column of its (single) n x n block - roots by the GEMMs, deflated columns by the
    // gather; lower levels leave the blocks between the nodes to this zero fill
    if (h < L || nv < n) {
      zero_matrix_kernel<<<dim3(gx, batch), 256, 0, st>>>(bufp[nxt], bufld[nxt], bufs[nxt], n, map);
      count_launch(1);
    }
    // the O(k^2) rank sort of a large merge: 1024 threads when there are few merges to fill the GPU with
    const int prep_t = (kmax >= 512 && nodes * batch <= 148) ? 1024 : 256;
    dc_prepare_kernel<<<dim3(nodes, batch), prep_t, use_smem ? smem : 0, st>>>(ws, e, dstride, h, lev, bufp[cur],
                                                                            bufld[cur], bufs[cur], use_smem, map);
    // a warp per root (latency: one large merge) or a thread per root (throughput: many small merges)
    static int sec_mode = -1;  // 0 auto, 1 warp, 2 thread
    if (sec_mode < 0) {
      const char* se = getenv("GSCF_B200_DC_SECULAR");
      sec_mode = !se ? 0 : (se[0] == 'w' ? 1 : (se[0] == 't' ? 2 : 0));
    }
    // (measured at n = 128: 4096 matrices 138.3 -> 128.1 ms per solve with the thread kernel, 512 matrices 22.2 -> 24.7 ms:
    // below ~250k roots per launch the serial K-loop of a thread is no longer hidden)
    const bool sec_thread = kmax <= 256 && (sec_mode == 2 || (sec_mode == 0 && kmax <= 160 &&
                                                              (long long)nodes * batch * kmax >= 262144));
    if (sec_thread)
      dc_secular_thread_kernel<<<dim3(ceil_div(kmax, 128), nodes, batch), 128, (size_t)2 * kmax * sizeof(double), st>>>(
          ws, h, 1 - lev, kmax, map);
    else
      dc_secular_kernel<<<dim3(ceil_div(kmax, 8), nodes, batch), 256, 0, st>>>(ws, h, 1 - lev, map);
    dc_zhat_kernel<<<dim3(ceil_div(kmax, 8), nodes, batch), 256, 0, st>>>(ws, h, map);
    const int partial = (h == L && nv < n) ? 1 : 0;  // top-level merge of a partial-spectrum solve
    if (partial) {
      dc_select_kernel<<<batch, 256, 0, st>>>(ws, 1 - lev, nv, map);
      count_launch();
    }
    const int vec_t = (kmax >= 512 && (long long)ceil_div(kmax, 32) * nodes * batch <= 148) ? 1024 : 256;
    dc_vectors_kernel<<<dim3(ceil_div(kmax, 32), nodes, batch), vec_t, 0, st>>>(ws, h, partial, map);
    const int gcpb = kmax <= 256 ? 16 : 1;  // columns per block of the gather
    dc_gather_kernel<<<dim3(ceil_div(kmax, gcpb), nodes, batch), 128, 0, st>>>(
        ws, h, 1 - lev, bufp[cur], bufld[cur], bufs[cur], bufp[nxt], bufld[nxt], bufs[nxt], nodes, partial, gcpb, map);
    count_launch(5);
    GSCF_CHECK_LAUNCH();
    GemmParams gp{};
    gp.descs = ws.descs;
    gp.kbatch = 1;
    GSCF_TRY(launch_gemm(false, true, gp, nodes * batch * 2, n1max, kmax, st));
    cur = nxt;
    lev = 1 - lev;
  }
  // ---- final ascending order ----------------------------------------------------------------------------
  if (n > 256 && batch <= 16)
    dc_rank_kernel<<<dim3(std::max(1, std::min(ceil_div(n, 4), 592 / batch)), batch), 128, 0, st>>>(ws, lev, info_dev, 1, map);
  else
    dc_rank_kernel<<<dim3(std::max(1, ceil_div(n, 128)), batch), 128, 0, st>>>(ws, lev, info_dev, 0, map);
  const int pcpb = n <= 256 ? 16 : 1;
  dc_permute_kernel<<<dim3(ceil_div(n, pcpb), batch), 128, 0, st>>>(ws, lev, bufp[cur], bufld[cur], bufs[cur], Z, ldz,
                                                                   strideZ, w, stride_w, info_dev, nv, pcpb, map);
  count_launch(2);
  GSCF_CHECK_LAUNCH();
  return GSCF_B200_OK;
}

}  // namespace gscfb
