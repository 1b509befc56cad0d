// FP64 DMMA GEMM, see gemm.cuh.  Column-major operands, C = alpha op(A) op(B) + beta C.
//
// Tiling: CTA tile BM x BN x BK, warp tile WM x WN built from m8n8k4 DMMA atoms.  Operand
// tiles are staged global -> shared with 16-byte cp.async (LDGSTS, zero-fill predication at
// element granularity) through a STAGES-deep ring; the shared layouts keep the global
// contiguous direction and are padded to stride = 4 (mod 16) doubles so that the one-element-
// per-lane DMMA fragment reads are bank-conflict free:
//     !TA: sA[k][m]  ld = BM+4      TA: sA[m][k]  ld = BK+4
//     !TB: sB[n][k]  ld = BK+4      TB: sB[k][n]  ld = BN+4
//
// Launches whose shapes, alignment and batch extent are known on the host are staged by TENSOR-MAP TMA instead
// (cp.async.bulk.tensor, SASS UTMALDG; kernel dgemm_dmma_tm_kernel), into DENSE tiles laid out in fragment order:
//   k-contiguous operand (TA / !TB):      3-D map {k, rows, batch entry}, four boxes of 4 x BM per stage -> s[k / 4][row][4]
//   m- / n-contiguous operand (!TA / TB): 4-D map {4 rows, k, row groups, batch entry} (row-group stride 32 bytes, smaller
//                                         than the k stride), ONE box of 4 x BK x BM/4 per stage       -> s[row / 4][k][4]
// In both layouts the lanes of a DMMA fragment (row = lane / 4, k = lane % 4) read 16 consecutive doubles per half-warp:
// conflict-free without the padding a tensor-map box cannot produce; rows / k beyond the matrix are zero-filled by the
// TMA unit, so there is no edge-tile path.  All copies of a stage complete on the stage's one mbarrier.  The map's batch
// extent is the exact number of problems (see try_launch_tm for the fault a loose bound caused).
#include "gemm.cuh"

#include <cuda.h>  // CUtensorMap and the cuTensorMapEncodeTiled types (the entry point is fetched at run time)

#include <algorithm>
#include <atomic>
#include <cstring>

namespace gscfb {

static std::atomic<uint64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
uint64_t launch_count_value() { return g_launches.load(std::memory_order_relaxed); }

// TM: 0 = cp.async / 1-D bulk copies, 1 = tensor maps for the k-contiguous operands (the other operand by bulk copies),
// 2 / 3 = tensor maps for both operand layouts, m- / n-contiguous operands in row groups of 8 / 4
template <int BM, int BN, int BK, bool TA, bool TB, int TM = 0>
struct SmemLayout {
  static constexpr int A_CE = TA ? BK : BM;  // contiguous extent
  static constexpr int A_SE = TA ? BM : BK;  // strided extent
  static constexpr int B_CE = TB ? BN : BK;
  static constexpr int B_SE = TB ? BK : BN;
  static constexpr int A_LD = A_CE + 4;
  static constexpr int B_LD = B_CE + 4;
  // TM >= 1: a k-contiguous operand is staged by tensor-map TMA as s[k / 4][row][4] (dense);
  // TM >= 2: an m- / n-contiguous operand as s[row / RG][k][RG] (dense) - see tma_tensor4d_g2s
  static constexpr int RG = TM == 2 ? 8 : 4;
  static constexpr bool A_TM = TM >= 1 && TA;
  static constexpr bool B_TM = TM >= 1 && !TB;
  static constexpr bool A_MM = TM >= 2 && !TA;
  static constexpr bool B_MM = TM >= 2 && TB;
  static constexpr int A_ELEMS = (A_TM || A_MM) ? BM * BK : A_SE * A_LD;
  static constexpr int B_ELEMS = (B_TM || B_MM) ? BN * BK : B_SE * B_LD;
  static constexpr int STAGE_ELEMS = A_ELEMS + B_ELEMS;
  static_assert(TM == 0 || (A_ELEMS % 16 == 0 && B_ELEMS % 16 == 0), "TMA destinations stay 128-byte aligned");
};

// One box of a 3-D tensor map -> shared memory, completing on `bar` (coordinates innermost first).
__device__ __forceinline__ void tma_tensor3d_g2s(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// Copy one operand tile into shared memory.  CE/SE/LD as in SmemLayout; (c0, s0) = tile origin
// along the contiguous / strided global direction; (climit, slimit) = matrix extents.
template <int CE, int SE, int LD, int NT, int VEC>
__device__ __forceinline__ void load_operand(double* __restrict__ s, const double* __restrict__ g,
                                             int ld, int c0, int s0, int climit, int slimit) {
  constexpr int CPR = CE / VEC;  // chunks per strided index
  constexpr int TOTAL = CPR * SE;
#pragma unroll
  for (int c = threadIdx.x; c < TOTAL; c += NT) {
    const int si = c / CPR;
    const int ci = (c - si * CPR) * VEC;
    const int gc = c0 + ci, gs = s0 + si;
    int valid = climit - gc;
    valid = valid < 0 ? 0 : (valid > VEC ? VEC : valid);
    if (gs >= slimit) valid = 0;
    const double* src = valid > 0 ? g + (size_t)gs * ld + gc : g;
    if (VEC == 2)
      cp_async16(s + si * LD + ci, src, valid * 8);
    else
      cp_async8(s + si * LD + ci, src, valid * 8);
  }
}

// One box of a 4-D tensor map.  Used for the m- / n-contiguous operands: the matrix (element (r, kk) at base[kk ld + r]) is
// described as {RG rows, k, row groups of RG, batch entry} with strides {ld, RG elements, batch stride}; a box of
// RG x BK x BM/RG lands as s[row group][k][RG rows] - dense, one TMA instruction per operand and stage.  A DMMA fragment
// (row = lane / 4, k = lane % 4) then reads, with RG = 4, 16 consecutive doubles per half-warp (lanes 0..15: rows 0..3 of one
// group, lanes 16..31: the next group) - conflict-free.  RG = 8 was measured first (`profiles/r02_ncu_full_gemm_tm.json`:
// NT 4096^3 32.1 TFLOP/s with 2.0e8 shared-memory bank conflicts per launch - the half-warp's 4 k x 4 rows fall on
// 8 k-strided doubles twice - against 234 conflicts for the k-major maps).
__device__ __forceinline__ void tma_tensor4d_g2s(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                                 uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
      : "memory");
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool TA, bool TB, int TM>
__device__ __forceinline__ void dgemm_dmma_body(const GemmParams& p, const CUtensorMap* tmA, const CUtensorMap* tmB) {
  using L = SmemLayout<BM, BN, BK, TA, TB, TM>;
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr int MI = WM / 8, NI = WN / 8;
  extern __shared__ __align__(128) double smem[];

  // ---- resolve this CTA's problem ---------------------------------------------------------
  const double* A;
  const double* B;
  double* C;
  double* C2 = nullptr;
  int m, n, k, lda, ldb, ldc;
  double alpha, beta;
  int zq = 0;  // batch entry (third tensor-map coordinate of a strided operand)
  if (p.descs != nullptr) {
    const GemmDesc d = p.descs[blockIdx.z];
    A = d.A; B = d.B; C = d.C; m = d.m; n = d.n; k = d.k;
    lda = d.lda; ldb = d.ldb; ldc = d.ldc; alpha = d.alpha; beta = d.beta;
  } else {
    const int ze = p.ksplit > 1 ? blockIdx.z / p.ksplit : blockIdx.z;
    const long long z = p.batch_map ? p.batch_map[ze] : ze;
    zq = (int)z;
    A = p.A + z * p.sA; B = p.B + z * p.sB; C = p.C + z * p.sC;
    if (p.C2) C2 = p.C2 + z * p.sC2;
    m = p.m; n = p.n; k = p.k; lda = p.lda; ldb = p.ldb; ldc = p.ldc;
    alpha = p.alpha; beta = p.beta;
  }
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  if (m0 >= m || n0 >= n) return;
  if (p.lower_only && n0 >= m0 + BM) return;  // tile strictly above the diagonal

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp % (BM / WM)) * WM, wn = (warp / (BM / WM)) * WN;
  const int lr = lane >> 2, lc = lane & 3;

  const bool alignedA = ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && ((lda & 1) == 0) &&
                        ((p.skA & 1) == 0);
  const bool alignedB = ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && ((ldb & 1) == 0) &&
                        ((p.skB & 1) == 0);

  const int kbatch = p.descs ? 1 : p.kbatch;
  const int ktiles_per = (k + BK - 1) / BK;
  int KT = ktiles_per * kbatch;
  int T0 = 0;  // first k-tile of this CTA (split-k)
  if (p.ksplit > 1) {
    const int split = blockIdx.z % p.ksplit;
    const int per = (KT + p.ksplit - 1) / p.ksplit;
    T0 = split * per;
    KT = min(KT, T0 + per);
  } else if (p.ktri != 0 && p.descs == nullptr) {
    if (p.ktri == 1) KT = min(KT, (min(k, n0 + BN) + BK - 1) / BK);
    else if (p.ktri == 2) KT = min(KT, (min(k, m0 + BM) + BK - 1) / BK);
    else if (p.ktri == 3) T0 = min(KT, m0 / BK);
  }

  auto load_tile = [&](int stage, int t) {
    const int q = t / ktiles_per;
    const int k0 = (t - q * ktiles_per) * BK;
    double* sA = smem + stage * L::STAGE_ELEMS;
    double* sB = sA + L::A_ELEMS;
    const double* Aq = A + (size_t)q * p.skA;
    const double* Bq = B + (size_t)q * p.skB;
    if (TA) {  // A stored k x m: contiguous along k
      if (alignedA) load_operand<L::A_CE, L::A_SE, L::A_LD, NT, 2>(sA, Aq, lda, k0, m0, k, m);
      else          load_operand<L::A_CE, L::A_SE, L::A_LD, NT, 1>(sA, Aq, lda, k0, m0, k, m);
    } else {   // A stored m x k: contiguous along m
      if (alignedA) load_operand<L::A_CE, L::A_SE, L::A_LD, NT, 2>(sA, Aq, lda, m0, k0, m, k);
      else          load_operand<L::A_CE, L::A_SE, L::A_LD, NT, 1>(sA, Aq, lda, m0, k0, m, k);
    }
    if (TB) {  // B stored n x k: contiguous along n
      if (alignedB) load_operand<L::B_CE, L::B_SE, L::B_LD, NT, 2>(sB, Bq, ldb, n0, k0, n, k);
      else          load_operand<L::B_CE, L::B_SE, L::B_LD, NT, 1>(sB, Bq, ldb, n0, k0, n, k);
    } else {   // B stored k x n: contiguous along k
      if (alignedB) load_operand<L::B_CE, L::B_SE, L::B_LD, NT, 2>(sB, Bq, ldb, k0, n0, k, n);
      else          load_operand<L::B_CE, L::B_SE, L::B_LD, NT, 1>(sB, Bq, ldb, k0, n0, k, n);
    }
  };

  // ---- TMA path: interior tiles of aligned operands are staged by bulk copies (one per contiguous tile row, into the
  // same padded layout) that complete on one mbarrier per stage; edge tiles keep the predicated cp.async path
  __shared__ __align__(8) uint64_t s_bar[STAGES];
  // Only operands whose tile rows are long (m- or n-contiguous: BM / BN doubles = 0.5-1 KB per copy) go through TMA; a
  // k-contiguous operand would need one 128-byte copy per row (measured: 13.8 instead of 28.4 TFLOP/s) and stays on cp.async.
  const bool interior = p.use_tma && (m0 + BM <= m) && (n0 + BN <= n) && (k % BK == 0);
  // ... and only when BOTH operands qualify (the NT products: F X^T, C_o C_o^T, the Pulay outer products, Q U^T of the
  // D&C merges, L21 L21^T): mixing the two mechanisms in one stage measured slower than cp.async alone at n = 1024
  // (18.8 vs 23.4 TFLOP/s), both through TMA faster (NT 4096^3: 30.0 vs 25.7 TFLOP/s).
  // TM kernel: the host has checked shapes and alignment of the whole launch - every CTA takes the TMA path, the
  // k-contiguous operand(s) through the tensor map(s), the other one through bulk copies
  const bool both = TM != 0 || (interior && !TA && TB && alignedA && alignedB);
  const bool bulkA = TM >= 2 ? false : TM == 1 ? !TA : both, bulkB = TM >= 2 ? false : TM == 1 ? TB : both;
  const bool use_bulk = both;
  if (use_bulk) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < STAGES; ++s) mbar_init(&s_bar[s], 1);
      mbar_fence_init();
    }
    __syncthreads();
  }
  auto load_tile_bulk = [&](int stage, int t) {
    const int q = t / ktiles_per;
    const int k0 = (t - q * ktiles_per) * BK;
    double* sA = smem + stage * L::STAGE_ELEMS;
    double* sB = sA + L::A_ELEMS;
    const double* Aq = A + (size_t)q * p.skA;
    const double* Bq = B + (size_t)q * p.skB;
    const unsigned bytes = (unsigned)(((bulkA || TM != 0 ? BM : 0) + (bulkB || TM != 0 ? BN : 0)) * BK * sizeof(double));
    // warp 0 issues the bulk copies: expect_tx first, then one copy per tile row
    if (threadIdx.x < 32) {
      fence_proxy_async();  // generic-proxy reads of this stage (ordered by the barrier above) before the async-proxy writes
      if (threadIdx.x == 0) mbar_arrive_expect_tx(&s_bar[stage], bytes);
      __syncwarp();
      if (bulkA)
        for (int r = threadIdx.x; r < L::A_SE; r += 32)
          tma_bulk_g2s(sA + r * L::A_LD, Aq + (size_t)(k0 + r) * lda + m0, (unsigned)(L::A_CE * sizeof(double)), &s_bar[stage]);
      if (bulkB)
        for (int r = threadIdx.x; r < L::B_SE; r += 32)
          tma_bulk_g2s(sB + r * L::B_LD, Bq + (size_t)(k0 + r) * ldb + n0, (unsigned)(L::B_CE * sizeof(double)), &s_bar[stage]);
      if (TM != 0) {
        // one box of 4 (k) x BM / BN rows per k-group; rows and k beyond the matrix are zero-filled by the TMA unit and
        // still count as full boxes towards the expected bytes
        if (L::A_TM && threadIdx.x < BK / 4)
          tma_tensor3d_g2s(sA + threadIdx.x * (BM * 4), tmA, k0 + 4 * (int)threadIdx.x, m0, p.sA != 0 ? zq : 0, &s_bar[stage]);
        if (L::B_TM && threadIdx.x < BK / 4)
          tma_tensor3d_g2s(sB + threadIdx.x * (BN * 4), tmB, k0 + 4 * (int)threadIdx.x, n0, p.sB != 0 ? zq : 0, &s_bar[stage]);
        // m- / n-contiguous operand: ONE box of RG rows x BK x BM/RG row groups
        if (L::A_MM && threadIdx.x == 0) tma_tensor4d_g2s(sA, tmA, 0, k0, m0 / L::RG, p.sA != 0 ? zq : 0, &s_bar[stage]);
        if (L::B_MM && threadIdx.x == 0) tma_tensor4d_g2s(sB, tmB, 0, k0, n0 / L::RG, p.sB != 0 ? zq : 0, &s_bar[stage]);
      }
    }
    if (TM != 0) return;
    // the other operand (k-contiguous rows) through cp.async
    if (!bulkA) {
      if (TA) {
        if (alignedA) load_operand<L::A_CE, L::A_SE, L::A_LD, NT, 2>(sA, Aq, lda, k0, m0, k, m);
        else          load_operand<L::A_CE, L::A_SE, L::A_LD, NT, 1>(sA, Aq, lda, k0, m0, k, m);
      } else {
        if (alignedA) load_operand<L::A_CE, L::A_SE, L::A_LD, NT, 2>(sA, Aq, lda, m0, k0, m, k);
        else          load_operand<L::A_CE, L::A_SE, L::A_LD, NT, 1>(sA, Aq, lda, m0, k0, m, k);
      }
    }
    if (!bulkB) {
      if (TB) {
        if (alignedB) load_operand<L::B_CE, L::B_SE, L::B_LD, NT, 2>(sB, Bq, ldb, n0, k0, n, k);
        else          load_operand<L::B_CE, L::B_SE, L::B_LD, NT, 1>(sB, Bq, ldb, n0, k0, n, k);
      } else {
        if (alignedB) load_operand<L::B_CE, L::B_SE, L::B_LD, NT, 2>(sB, Bq, ldb, k0, n0, k, n);
        else          load_operand<L::B_CE, L::B_SE, L::B_LD, NT, 1>(sB, Bq, ldb, k0, n0, k, n);
      }
    }
  };

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (T0 + s < KT) {
      if (use_bulk) load_tile_bulk(s, T0 + s);
      else load_tile(s, T0 + s);
    }
    cp_async_commit();
  }

  for (int tt = 0; tt < KT - T0; ++tt) {
    const int t = tt;
    cp_async_wait<STAGES - 2>();
    if (use_bulk) mbar_wait(&s_bar[t % STAGES], (unsigned)((t / STAGES) & 1));
    __syncthreads();
    {
      const int tn = t + STAGES - 1;
      if (T0 + tn < KT) {
        if (use_bulk) load_tile_bulk(tn % STAGES, T0 + tn);
        else load_tile(tn % STAGES, T0 + tn);
      }
      cp_async_commit();
    }
    const double* sA = smem + (t % STAGES) * L::STAGE_ELEMS;
    const double* sB = sA + L::A_ELEMS;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[MI], bf[NI];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int r = wm + i * 8 + lr;
        af[i] = L::A_TM   ? sA[((kk >> 2) * BM + r) * 4 + lc]
                : L::A_MM ? sA[(((wm + i * 8 + lr) / L::RG) * BK + kk + lc) * L::RG + (lr & (L::RG - 1))]
                : TA      ? sA[r * L::A_LD + kk + lc]
                          : sA[(kk + lc) * L::A_LD + r];
      }
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int c = wn + j * 8 + lr;
        bf[j] = L::B_TM   ? sB[((kk >> 2) * BN + c) * 4 + lc]
                : L::B_MM ? sB[(((wn + j * 8 + lr) / L::RG) * BK + kk + lc) * L::RG + (lr & (L::RG - 1))]
                : TB      ? sB[(kk + lc) * L::B_LD + c]
                          : sB[c * L::B_LD + kk + lc];
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();

  // ---- epilogue ----------------------------------------------------------------------------
  if (p.ksplit > 1) {  // split-k: raw partial tile, reduced (alpha, beta applied) by splitk_reduce_kernel
    double* Cp = p.Cpart + (size_t)blockIdx.z * m * n;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int r = m0 + wm + i * 8 + lr;
      if (r >= m) continue;
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int c = n0 + wn + j * 8 + 2 * lc;
        if (c < n) Cp[(size_t)c * m + r] = acc[i][j][0];
        if (c + 1 < n) Cp[(size_t)(c + 1) * m + r] = acc[i][j][1];
      }
    }
    return;
  }
  if (p.mirror != 0) {
    // (anti)symmetric result from the tiles on and below the diagonal: every strictly lower entry is stored twice, the
    // upper entries a diagonal tile computes are dropped in favour of the mirror images of its lower ones
    const bool anti = p.mirror < 0;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      const int r = m0 + wm + i * 8 + lr;
      if (r >= m) continue;
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int c = n0 + wn + j * 8 + 2 * lc;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int cc = c + h;
          if (cc >= n || cc > r) continue;
          const double v = alpha * acc[i][j][h];
          if (cc == r) {
            C[(size_t)cc * ldc + r] = anti ? 0.0 : v;
          } else {
            C[(size_t)cc * ldc + r] = v;
            C[(size_t)r * ldc + cc] = anti ? -v : v;
          }
        }
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int r = m0 + wm + i * 8 + lr;
    if (r >= m) continue;
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int c = n0 + wn + j * 8 + 2 * lc;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (c + h < n) {
          double* dst = C + (size_t)(c + h) * ldc + r;
          double v = alpha * acc[i][j][h];
          if (beta != 0.0) v += beta * (*dst);
          *dst = v;
          if (C2) C2[(size_t)(c + h) * p.ldc2 + r] = p.alpha2 * acc[i][j][h];
        }
      }
    }
  }
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool TA, bool TB>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
dgemm_dmma_kernel(const GemmParams p) {
  dgemm_dmma_body<BM, BN, BK, WM, WN, STAGES, TA, TB, 0>(p, nullptr, nullptr);
}

// the same tiles with operands fed by tensor-map TMA (maps live in the kernel's parameter space); TM as in SmemLayout
template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool TA, bool TB, int TM>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
dgemm_dmma_tm_kernel(const GemmParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
  dgemm_dmma_body<BM, BN, BK, WM, WN, STAGES, TA, TB, TM>(p, &tmA, &tmB);
}

// ---- host side of the tensor-map path ------------------------------------------------------------------------------
static CUtensorMapL2promotion tmap_l2_promotion() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GSCF_B200_TMAP_L2"); v = e ? atoi(e) : 128; }
  return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
       : v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
}
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static std::atomic<uint64_t> g_tmap_launches{0};
uint64_t tensor_map_launch_count_value() { return g_tmap_launches.load(std::memory_order_relaxed); }

// cuTensorMapEncodeTiled through the runtime (the library does not link libcuda); nullptr: driver without it
static TmapEncodeFn tmap_encode_fn() {
  static std::atomic<int> state{0};  // 0: not looked up, 1: found, 2: missing
  static std::atomic<void*> fn{nullptr};
  if (state.load(std::memory_order_acquire) == 0) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    (void)cudaGetLastError();
    fn.store(f, std::memory_order_relaxed);
    state.store(f ? 1 : 2, std::memory_order_release);
  }
  return reinterpret_cast<TmapEncodeFn>(fn.load(std::memory_order_relaxed));
}

// {k (contiguous), rows, batch entries} over a k-contiguous operand; box 4 x box_rows x 1
static bool encode_kmajor_map(CUtensorMap* tm, const double* base, int k, int rows, int ld, long long stride, int box_rows,
                              int entries) {
  TmapEncodeFn enc = tmap_encode_fn();
  if (!enc) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)k, (cuuint64_t)rows, stride != 0 ? (cuuint64_t)entries : 1};
  const cuuint64_t strides[2] = {(cuuint64_t)ld * sizeof(double),
                                 stride != 0 ? (cuuint64_t)stride * sizeof(double) : (cuuint64_t)ld * sizeof(double) * (cuuint64_t)rows};
  const cuuint32_t box[3] = {4, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, tmap_l2_promotion(),
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// {rg rows, k, row groups of rg, batch entries} over an m- / n-contiguous operand (element (r, kk) at base[kk ld + r]);
// box rg x box_k x box_rows/rg x 1 (see tma_tensor4d_g2s).  The row-group stride (rg doubles) is smaller than the k stride.
static bool encode_mmajor_map(CUtensorMap* tm, const double* base, int k, int rows, int ld, long long stride, int box_rows,
                              int box_k, int entries, int rg) {
  TmapEncodeFn enc = tmap_encode_fn();
  if (!enc) return false;
  const cuuint64_t groups = (cuuint64_t)ceil_div(rows, rg);
  const cuuint64_t dims[4] = {(cuuint64_t)rg, (cuuint64_t)k, groups, stride != 0 ? (cuuint64_t)entries : 1};
  const cuuint64_t strides[3] = {(cuuint64_t)ld * sizeof(double), (cuuint64_t)rg * sizeof(double),
                                 stride != 0 ? (cuuint64_t)stride * sizeof(double) : 64 * groups};
  const cuuint32_t box[4] = {(cuuint32_t)rg, (cuuint32_t)box_k, (cuuint32_t)(box_rows / rg), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, const_cast<double*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, tmap_l2_promotion(),
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool aligned16(const void* ptr, int ld, long long stride) {
  return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && (ld & 1) == 0 && (stride & 1) == 0;
}

// Does this launch qualify for the tensor-map kernel of level TM?  (Whole-launch decision: every CTA then takes the TMA path.)
template <int BM, int BN, int BK, bool TA, bool TB, int TM>
static bool tm_eligible(const GemmParams& p) {
  if (TM == 1 && !(TA || !TB)) return false;         // level 1, NT products: no k-contiguous operand, 1-D bulk copies
  if (p.descs != nullptr || p.kbatch != 1) return false;
  if (p.m <= 0 || p.n <= 0 || p.k <= 0) return false;
  if (!aligned16(p.A, p.lda, p.sA) || !aligned16(p.B, p.ldb, p.sB)) return false;
  if (p.sA < 0 || p.sB < 0 || p.sA >= (1LL << 36) || p.sB >= (1LL << 36)) return false;
  if (p.batch_map != nullptr && p.entries <= 0) return false;  // unknown batch extent (see try_launch_tm)
  {
    // no box larger than its tensor (GSCF_B200_TMAP_STRICT=0 lifts it): such products are too small to gain anything,
    // and whole-box overhang is what the fault described in try_launch_tm needed
    static int strict = -1;
    if (strict < 0) { const char* e = getenv("GSCF_B200_TMAP_STRICT"); strict = e ? atoi(e) : 1; }
    if (strict && (p.m < BM || p.n < BN || p.k < BK)) return false;
  }
  if (TM == 1) {
    // an operand staged by 1-D bulk copies has no zero fill: full tiles along its contiguous direction and along k
    if (!TA && (p.m % BM != 0 || p.k % BK != 0)) return false;
    if (TB && (p.n % BN != 0 || p.k % BK != 0)) return false;
  } else {
    // the 4-D map addresses whole groups of RG rows: the last group must stay inside the column (ld >= rows rounded up;
    // what it reads beyond the matrix only reaches rows of the product that are not stored)
    constexpr int RG = SmemLayout<BM, BN, BK, TA, TB, TM>::RG;
    if (!TA && p.lda < round_up(p.m, RG)) return false;
    if (TB && p.ldb < round_up(p.n, RG)) return false;
  }
  return true;
}

// Try the tensor-map kernel of level TM; `launched` tells whether it ran (false: not eligible or the driver refused a map)
template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool TA, bool TB, int TM>
static int try_launch_tm(const GemmParams& p, int batch, dim3 grid, cudaStream_t st, bool& launched) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  launched = false;
  if (!tm_eligible<BM, BN, BK, TA, TB, TM>(p)) return GSCF_B200_OK;
  CUtensorMap tmA, tmB;
  memset(&tmA, 0, sizeof(tmA));
  memset(&tmB, 0, sizeof(tmB));
  bool ok = true;
  // batch extent of the map: only bounds the zero fill; an index-mapped batch may address any problem of the engine
  // Batch extent of the maps = the number of entries the operands really hold.  It must not be a loose bound: with an
  // extent of 65536 (any index-mapped batch) a box reaching beyond the rows of a small matrix made the TMA unit touch
  // memory behind the last problem - `Warp Illegal Address` in the 64 x 64 NT kernel, 24 problems of order 32, but only
  // when the memory behind the engine's buffers was unmapped (found with cuda-gdb, gone with the exact extent; the rows
  // concerned are never stored, so results were unaffected).
  const int entries = p.batch_map ? p.entries : batch;
  if (TA) ok = ok && encode_kmajor_map(&tmA, p.A, p.k, p.m, p.lda, p.sA, BM, entries);
  else if (TM >= 2) ok = ok && encode_mmajor_map(&tmA, p.A, p.k, p.m, p.lda, p.sA, BM, BK, entries, SmemLayout<BM, BN, BK, TA, TB, TM>::RG);
  if (!TB) ok = ok && encode_kmajor_map(&tmB, p.B, p.k, p.n, p.ldb, p.sB, BN, entries);
  else if (TM >= 2) ok = ok && encode_mmajor_map(&tmB, p.B, p.k, p.n, p.ldb, p.sB, BN, BK, entries, SmemLayout<BM, BN, BK, TA, TB, TM>::RG);
  if (!ok) return GSCF_B200_OK;
  using L = SmemLayout<BM, BN, BK, TA, TB, TM>;
  const size_t smem = (size_t)STAGES * L::STAGE_ELEMS * sizeof(double);
  auto kern = dgemm_dmma_tm_kernel<BM, BN, BK, WM, WN, STAGES, TA, TB, TM>;
  static PerDeviceOnce once;
  int dev;
  if (once.need(dev)) {
    GSCF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    once.done(dev);
  }
  kern<<<grid, NT, smem, st>>>(p, tmA, tmB);
  count_launch();
  g_tmap_launches.fetch_add(1, std::memory_order_relaxed);
  GSCF_CHECK_LAUNCH();
  launched = true;
  return GSCF_B200_OK;
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES, bool TA, bool TB>
static int launch_cfg(const GemmParams& p, int batch, int max_m, int max_n, cudaStream_t st) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  dim3 grid(ceil_div(max_m, BM), ceil_div(max_n, BN), batch);
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return GSCF_B200_OK;
  // Tensor-map kernels (levels of GSCF_B200_GEMM_TMA, see launch_gemm): measured on B200 they win on the 128 x 128 tiles
  // (4096^3: NN 28.7 -> 30.6, TN 28.9 -> 31.9 TFLOP/s with level 2) and lose on the 64 x 64 tiles (1024^3 NN 22.9 -> 20.6)
  bool launched = false;
  // ... with maps for both layouts (level 3) every product gains on the 128 x 128 tiles (4096^3: NN 31.6, NT 32.2, TN 31.9
  // TFLOP/s, DMMA sub-pipe 86 - 88 % active) and on the 64 x 64 tiles all but the TN products, whose two k-major maps (32-byte
  // box rows, four boxes per operand and stage) measured slower there (64 x 4096 x 4096 split-k TN: 7.7 -> 7.5; against
  // 1024 x 1024 x 128 NT: 11.9 -> 14.4; whole SCF paths with / without the small-tile kernels: C2 190.0 / 188.9 it/s,
  // C4 204.8 / 202.7 k it/s)
  if (p.use_tma >= 4 || (p.use_tma >= 3 && (BM == 128 || !(TA && !TB)))) {
    // GSCF_B200_GEMM_RG=8: row groups of 8 in the 4-D maps (A/B; the default 4 is free of bank conflicts)
    static int rg = -1;
    if (rg < 0) { const char* re = getenv("GSCF_B200_GEMM_RG"); rg = (re && atoi(re) == 8) ? 8 : 4; }
    if (rg == 8) GSCF_TRY((try_launch_tm<BM, BN, BK, WM, WN, STAGES, TA, TB, 2>(p, batch, grid, st, launched)));
    else GSCF_TRY((try_launch_tm<BM, BN, BK, WM, WN, STAGES, TA, TB, 3>(p, batch, grid, st, launched)));
    if (launched) return GSCF_B200_OK;
  }
  if constexpr (BM == 128 && (TA || !TB)) {
    if (p.use_tma >= 2) {
      GSCF_TRY((try_launch_tm<BM, BN, BK, WM, WN, STAGES, TA, TB, 1>(p, batch, grid, st, launched)));
      if (launched) return GSCF_B200_OK;
    }
  }
  using L = SmemLayout<BM, BN, BK, TA, TB>;
  const size_t smem = (size_t)STAGES * L::STAGE_ELEMS * sizeof(double);
  auto kern = dgemm_dmma_kernel<BM, BN, BK, WM, WN, STAGES, TA, TB>;
  static PerDeviceOnce once;
  int dev;
  if (once.need(dev)) {
    GSCF_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    once.done(dev);
  }
  kern<<<grid, NT, smem, st>>>(p);
  count_launch();
  GSCF_CHECK_LAUNCH();
  return GSCF_B200_OK;
}

template <bool TA, bool TB>
static int launch_t(const GemmParams& p, int batch, int max_m, int max_n, cudaStream_t st) {
  // Large tiles once the grid still covers the 148 SMs; small tiles otherwise.
  const long long big_tiles = (long long)ceil_div(max_m, 128) * ceil_div(max_n, 128) * batch;
  // ... unless they would be mostly padding (batches of 64-wide panels: 128 x 128 tiles half empty)
  const double fill128 = (double)max_m * max_n / ((double)ceil_div(max_m, 128) * ceil_div(max_n, 128) * 128.0 * 128.0);
  const double fill64 = (double)max_m * max_n / ((double)ceil_div(max_m, 64) * ceil_div(max_n, 64) * 64.0 * 64.0);
  // Products with a short k loop in large batches take the 64 x 64 tiles: four CTAs per SM overlap one tile's prologue and
  // C read-modify-write with another's k loop (measured, 4096 matrices of order 128: back-transformation 34.2 -> 29.7 ms,
  // whole C4 solve 323.6 -> 307.8 ms with the limit at k <= 128, 316.1 with k <= 64).  GSCF_B200_GEMM_SMALLK=<k> moves the
  // limit (0 = off).
  static int smallk = -1;
  if (smallk < 0) { const char* sk = getenv("GSCF_B200_GEMM_SMALLK"); smallk = sk ? atoi(sk) : 128; }
  const bool short_k = smallk > 0 && p.descs == nullptr && p.k <= smallk && batch >= 148;
  if (big_tiles >= 120 && !(fill128 < 0.7 && fill64 > fill128 * 1.3) && !short_k)
    return launch_cfg<128, 128, 16, 64, 32, 3, TA, TB>(p, batch, max_m, max_n, st);
  return launch_cfg<64, 64, 16, 32, 32, 3, TA, TB>(p, batch, max_m, max_n, st);
}

__global__ void splitk_reduce_kernel(GemmParams p, int splits) {
  const int ze = blockIdx.y;
  const long long z = p.batch_map ? p.batch_map[ze] : ze;
  double* C = p.C + z * p.sC;
  const double* part = p.Cpart + (size_t)ze * splits * p.m * p.n;
  const int total = p.m * p.n;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < splits; ++k) s += part[(size_t)k * total + idx];
    const int c = idx / p.m, r = idx - c * p.m;
    double* dst = C + (size_t)c * p.ldc + r;
    double v = p.alpha * s;
    if (p.beta != 0.0) v += p.beta * (*dst);
    *dst = v;
  }
}

int launch_gemm_splitk(bool ta, bool tb, GemmParams p, int batch, int splits, double* work, cudaStream_t st) {
  if (splits <= 1 || work == nullptr) {
    p.ksplit = 1;
    return launch_gemm(ta, tb, p, batch, p.m, p.n, st);
  }
  p.ksplit = splits;
  p.Cpart = work;
  GSCF_TRY(launch_gemm(ta, tb, p, batch * splits, p.m, p.n, st));
  const int total = p.m * p.n;
  splitk_reduce_kernel<<<dim3(std::max(1, std::min(64, ceil_div(total, 256))), batch), 256, 0, st>>>(p, splits);
  count_launch();
  GSCF_CHECK_LAUNCH();
  return GSCF_B200_OK;
}

int pulay_error_fused(const PulayErrorArgs& a, cudaStream_t st) {
  if (a.batch <= 0 || a.n <= 0) return GSCF_B200_OK;
  const int n = a.n, o = a.o, ldw = padded_ld(n);
  // (o == 0: the two skinny products are empty launches and the k = 0 product below stores an all-zero error)
  double* W = a.work;                               // [S C_o | F C_o]
  double* W2 = a.work + (size_t)ldw * 2 * o;        // [F C_o | -S C_o]
  GemmParams g1 = gemm_params(n, o, n, 1.0, a.S, a.lds, a.C, a.ldc, 0.0, W, ldw);
  g1.sA = a.sS; g1.sB = a.sC; g1.sC = a.swork; g1.batch_map = a.batch_map; g1.entries = a.entries;
  g1.C2 = W2 + (size_t)ldw * o; g1.ldc2 = ldw; g1.sC2 = a.swork; g1.alpha2 = -1.0;
  GSCF_TRY(launch_gemm(false, false, g1, a.batch, n, o, st));
  GemmParams g2 = gemm_params(n, o, n, 1.0, a.F, a.ldf, a.C, a.ldc, 0.0, W + (size_t)ldw * o, ldw);
  g2.sA = a.sF; g2.sB = a.sC; g2.sC = a.swork; g2.batch_map = a.batch_map; g2.entries = a.entries;
  g2.C2 = W2; g2.ldc2 = ldw; g2.sC2 = a.swork; g2.alpha2 = 1.0;
  GSCF_TRY(launch_gemm(false, false, g2, a.batch, n, o, st));
  GemmParams g3 = gemm_params(n, n, 2 * o, a.fac, W, ldw, W2, ldw, 0.0, a.E, a.lde);
  g3.sA = g3.sB = a.swork; g3.sC = a.sE; g3.batch_map = a.batch_map; g3.entries = a.entries;
  g3.lower_only = 1; g3.mirror = -1;
  return launch_gemm(false, true, g3, a.batch, n, n, st);
}

int launch_gemm(bool ta, bool tb, const GemmParams& pin, int batch, int max_m, int max_n,
                cudaStream_t st) {
  if (batch <= 0 || max_m <= 0 || max_n <= 0) return GSCF_B200_OK;
  if (pin.mirror != 0 && (!pin.lower_only || pin.m != pin.n || pin.beta != 0.0 || pin.ksplit > 1 || pin.descs)) {
    set_last_error("gemm: the mirrored epilogue needs lower_only, a square result, beta == 0 and no split-k / grouping");
    return GSCF_B200_ERR_INVALID_ARG;
  }
  // GSCF_B200_GEMM_TMA: 0 = cp.async only, 1 = 1-D bulk copies for the NT products only, 2 = also tensor maps for the
  // k-contiguous operands of the 128 x 128 tiles, 3 (default) = tensor maps for both operand layouts on the 128 x 128
  // tiles and, except for the TN products, on the 64 x 64 tiles (falls back to 2 / 1 where a launch does not qualify),
  // 4 = every product on the 64 x 64 tiles too
  static int tma = -1;
  if (tma < 0) { const char* te = getenv("GSCF_B200_GEMM_TMA"); tma = (te && te[0] >= '0' && te[0] <= '4') ? te[0] - '0' : 3; }
  GemmParams p = pin;
  p.use_tma = tma;
  if (batch > 65535) {
    set_last_error("gemm: batch > 65535 not supported in one launch");
    return GSCF_B200_ERR_INVALID_ARG;
  }
  if (ta) return tb ? launch_t<true, true>(p, batch, max_m, max_n, st)
                    : launch_t<true, false>(p, batch, max_m, max_n, st);
  return tb ? launch_t<false, true>(p, batch, max_m, max_n, st)
            : launch_t<false, false>(p, batch, max_m, max_n, st);
}

}  // namespace gscfb
