// Synthetic density-fitted Fock builder: a device-resident FocklikeMatrix_i plugin
// (src/gscf/FocklikeMatrix_i.hh:53-117, ScfProblemMatrix_i.hh:32-44) used by bench.py and the
// tests.  gscf itself contains no integrals; this is the generator of SURVEY.md 8(d),
// restated on the CPU in oracle/synthetic.py.  Timed separately from the SCF path.
//
//   restricted:   gamma_Q = 2 <B_Q C_o, C_o>,  J = sum_Q gamma_Q B_Q,  K = sum_Q (B_Q C_o)(B_Q C_o)^T
//                 F = H + J - K,  E1e = 2 <H, D>,  E2e = <J - K, D>
//   unrestricted: gamma_Q = <B_Q C_a, C_a> + <B_Q C_b, C_b>,  F^s = H + J - K^s,
//                 E1e = <H, Da+Db>,  E2e = 1/2 [<J, Da+Db> - <Ka, Da> - <Kb, Db>]
//
// HBM layout: the aux tensors of one problem are one tall column-major matrix
// Btall[(Q*ld + r), c] (leading dimension n_aux*ld), so that Y = Btall * C_o is a single DMMA
// GEMM and K = sum_Q Y_Q Y_Q^T is one GEMM whose k loop runs over (Q, occupied orbital).
#include "common.cuh"
#include "gemm.cuh"

#include <vector>

using namespace gscfb;

struct gscf_b200_dfock {
  int n = 0, ld = 0, nblk = 1, occ[2] = {0, 0}, omax = 1, naux = 0, P = 0;
  long long ldt = 0;          // n_aux * ld
  long long mstride = 0;      // ld * n
  double* S = nullptr;        // [P] ld x n
  double* H = nullptr;        // [P] ld x n
  double* Bt = nullptr;       // [P] ldt x n
  double* Y = nullptr;        // [P][nblk] ldt x omax
  double* K = nullptr;        // [P][nblk] ld x n
  double* gamma = nullptr;    // [P][naux]
  double* epart = nullptr;    // [P][kEBlocks][2]
};

namespace {

constexpr int kEBlocks = 296;  // energy partials per problem (2 CTAs per SM)

__device__ __forceinline__ double splitmix_u(unsigned long long seed, unsigned long long k) {
  unsigned long long z = seed + (k + 1ULL) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z ^= z >> 31;
  return (double)(z >> 11) * 0x1.0p-52 - 1.0;
}

__device__ __forceinline__ double symrand(unsigned long long seed, int i, int j) {
  const unsigned long long hi = i > j ? i : j, lo = i > j ? j : i;
  return splitmix_u(seed, hi * (hi + 1ULL) / 2ULL + lo);
}

__global__ void generate_kernel(double* __restrict__ S, double* __restrict__ H, double* __restrict__ Bt,
                                int n, int ld, int naux, int n_alpha, const long long* __restrict__ seeds,
                                double g) {
  const int p = blockIdx.z;
  const int which = blockIdx.y;  // 0: S, 1: H, 2+Q: B_Q
  const unsigned long long base = (unsigned long long)(1000LL * seeds[p]);
  const long long mstride = (long long)ld * n;
  const long long ldt = (long long)naux * ld;
  const double sS = 0.25 / sqrt((double)n), sH = 0.5 / sqrt((double)n);
  const double sB = g / sqrt((double)n * (double)naux);
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (long long)n * n;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx / n), r = (int)(idx - (long long)c * n);
    if (which == 0) {
      S[p * mstride + (size_t)c * ld + r] = (r == c) ? 1.0 : 0.0 + sS * symrand(base + 1ULL, r, c);
    } else if (which == 1) {
      double v = sH * symrand(base + 2ULL, r, c);
      if (r == c) v += ((2.0 * r) / n - 1.0) + (r >= n_alpha ? 0.5 : 0.0);
      H[p * mstride + (size_t)c * ld + r] = v;
    } else {
      const int Q = which - 2;
      Bt[p * ldt * n + (size_t)c * ldt + (size_t)Q * ld + r] = sB * symrand(base + 10ULL + Q, r, c);
    }
  }
}

// gamma[p][Q] = fac * sum_b <Y_Q^b, C_o^b>
__global__ void gamma_kernel(const double* __restrict__ Y, long long y_pstride, long long y_bstride,
                             long long ldt, const double* __restrict__ C, long long c_pstride,
                             long long c_bstride, int ldc, int n, int ld, int nblk, int occ0, int occ1,
                             double fac, const int* __restrict__ act, double* __restrict__ gamma, int naux) {
  __shared__ double red[33];
  const int Q = blockIdx.x;
  const long long p = act[blockIdx.y];
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) {
    const int o = b == 0 ? occ0 : occ1;
    const double* y = Y + p * y_pstride + b * y_bstride + (size_t)Q * ld;
    const double* c = C + p * c_pstride + b * c_bstride;
    for (int idx = threadIdx.x; idx < n * o; idx += blockDim.x) {
      const int a = idx / n, i = idx - a * n;
      s = fma(y[(size_t)a * ldt + i], c[(size_t)a * ldc + i], s);
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) gamma[p * naux + Q] = fac * s;
}

// F^b = H + J - K^b with J = sum_Q gamma_Q B_Q streamed once; energy partials.
template <int NBLK>
__global__ void __launch_bounds__(256)
assemble_kernel(const double* __restrict__ H, const double* __restrict__ Bt, const double* __restrict__ K,
                const double* __restrict__ gamma, const double* __restrict__ D, long long d_pstride,
                long long d_bstride, double* __restrict__ F, long long f_pstride, long long f_bstride,
                int n, int ld, int naux, const int* __restrict__ act, double* __restrict__ epart) {
  __shared__ double red[33];
  extern __shared__ double sg[];  // gamma of this problem
  const long long p = act[blockIdx.y];
  const long long mstride = (long long)ld * n, ldt = (long long)naux * ld;
  for (int q = threadIdx.x; q < naux; q += blockDim.x) sg[q] = gamma[p * naux + q];
  __syncthreads();
  const double* h = H + p * mstride;
  const double* bt = Bt + p * ldt * n;
  double e1 = 0.0, e2 = 0.0;
  const int pairs = n >> 1;  // n even fast path handled by caller through ld (n odd: scalar tail)
  const long long total = (long long)(pairs + (n & 1)) * n;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total;
       w += (long long)gridDim.x * blockDim.x) {
    const int per_col = pairs + (n & 1);
    const int c = (int)(w / per_col);
    const int r = (int)(w - (long long)c * per_col) * 2;
    const bool two = r + 1 < n;
    double j0 = 0.0, j1 = 0.0;
    const double* bcol = bt + (size_t)c * ldt + r;
    if (two) {
#pragma unroll 4
      for (int q = 0; q < naux; ++q) {
        const double2 v = *reinterpret_cast<const double2*>(bcol + (size_t)q * ld);
        j0 = fma(sg[q], v.x, j0);
        j1 = fma(sg[q], v.y, j1);
      }
    } else {
      for (int q = 0; q < naux; ++q) j0 = fma(sg[q], bcol[(size_t)q * ld], j0);
    }
    const size_t off = (size_t)c * ld + r;
    const double h0 = h[off], h1 = two ? h[off + 1] : 0.0;
    double dt0 = 0.0, dt1 = 0.0;
#pragma unroll
    for (int b = 0; b < NBLK; ++b) {
      const double* k = K + (p * NBLK + b) * mstride;
      const double* d = D + p * d_pstride + b * d_bstride;
      double* f = F + p * f_pstride + b * f_bstride;
      const double k0 = k[off], k1 = two ? k[off + 1] : 0.0;
      const double d0 = d[off], d1 = two ? d[off + 1] : 0.0;
      f[off] = h0 + j0 - k0;
      if (two) f[off + 1] = h1 + j1 - k1;
      dt0 += d0;
      dt1 += d1;
      if (NBLK == 1) {
        e2 = fma(j0 - k0, d0, e2);
        e2 = fma(j1 - k1, d1, e2);
      } else {
        e2 = fma(-0.5 * k0, d0, e2);
        e2 = fma(-0.5 * k1, d1, e2);
      }
    }
    if (NBLK == 1) {
      e1 = fma(2.0 * h0, dt0, e1);
      e1 = fma(2.0 * h1, dt1, e1);
    } else {
      e1 = fma(h0, dt0, e1);
      e1 = fma(h1, dt1, e1);
      e2 = fma(0.5 * j0, dt0, e2);
      e2 = fma(0.5 * j1, dt1, e2);
    }
  }
  e1 = block_sum(e1, red);
  e2 = block_sum(e2, red);
  if (threadIdx.x == 0) {
    epart[((size_t)p * kEBlocks + blockIdx.x) * 2 + 0] = e1;
    epart[((size_t)p * kEBlocks + blockIdx.x) * 2 + 1] = e2;
  }
}

__global__ void energy_final_kernel(const double* __restrict__ epart, int nblocks, const int* __restrict__ act,
                                    double* __restrict__ energies) {
  const long long p = act[blockIdx.x];
  if (threadIdx.x < 2) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += epart[((size_t)p * kEBlocks + b) * 2 + threadIdx.x];
    energies[p * 2 + threadIdx.x] = s;
  }
}

int dfock_build(void* ctx, const gscf_b200_fock_args* a) {
  gscf_b200_dfock* f = (gscf_b200_dfock*)ctx;
  if (!f || !a || a->n_bas != f->n || a->n_blocks != f->nblk || a->ld != f->ld) {
    set_last_error("dfock: argument mismatch with the plugin configuration");
    return GSCF_B200_ERR_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)a->stream;
  const int n = f->n, ld = f->ld, nact = a->n_active;
  const long long y_bstride = f->ldt * f->omax, y_pstride = y_bstride * f->nblk;
  for (int b = 0; b < f->nblk; ++b) {
    const int o = f->occ[b];
    // Y^b = Btall * C_o^b
    GemmParams g = gemm_params((int)f->ldt, o, n, 1.0, f->Bt, (int)f->ldt, a->C + b * a->block_stride, a->ld,
                               0.0, f->Y + b * y_bstride, (int)f->ldt);
    g.sA = f->ldt * n; g.sB = a->problem_stride; g.sC = y_pstride; g.batch_map = a->active_dev;
    GSCF_TRY(launch_gemm(false, false, g, nact, (int)f->ldt, o, st));
    // K^b = sum_Q Y_Q Y_Q^T
    GemmParams k = gemm_params(n, n, o, 1.0, f->Y + b * y_bstride, (int)f->ldt, f->Y + b * y_bstride, (int)f->ldt,
                               0.0, f->K + b * f->mstride, ld);
    k.sA = k.sB = y_pstride; k.sC = f->mstride * f->nblk; k.kbatch = f->naux; k.skA = k.skB = ld;
    k.batch_map = a->active_dev;
    GSCF_TRY(launch_gemm(false, true, k, nact, n, n, st));
  }
  gamma_kernel<<<dim3(f->naux, nact), 256, 0, st>>>(f->Y, y_pstride, y_bstride, f->ldt, a->C, a->problem_stride,
                                                    a->block_stride, a->ld, n, ld, f->nblk, f->occ[0], f->occ[1],
                                                    f->nblk == 1 ? 2.0 : 1.0, a->active_dev, f->gamma, f->naux);
  count_launch();
  GSCF_CHECK_LAUNCH();
  long long items = (long long)((n + 1) / 2) * n;
  int blocks = (int)std::min<long long>(kEBlocks, ceil_div_ll(items, 256 * 4));
  if (blocks < 1) blocks = 1;
  const size_t smem = (size_t)f->naux * sizeof(double);
  if (f->nblk == 1)
    assemble_kernel<1><<<dim3(blocks, nact), 256, smem, st>>>(f->H, f->Bt, f->K, f->gamma, a->D, a->problem_stride,
                                                              a->block_stride, a->F_out, a->F_problem_stride,
                                                              a->block_stride, n, ld, f->naux, a->active_dev, f->epart);
  else
    assemble_kernel<2><<<dim3(blocks, nact), 256, smem, st>>>(f->H, f->Bt, f->K, f->gamma, a->D, a->problem_stride,
                                                              a->block_stride, a->F_out, a->F_problem_stride,
                                                              a->block_stride, n, ld, f->naux, a->active_dev, f->epart);
  count_launch();
  GSCF_CHECK_LAUNCH();
  energy_final_kernel<<<nact, 32, 0, st>>>(f->epart, blocks, a->active_dev, a->energies_out);
  count_launch();
  GSCF_CHECK_LAUNCH();
  return GSCF_B200_OK;
}

}  // namespace

extern "C" {

int gscf_b200_padded_ld(int n) { return padded_ld(n); }

int gscf_b200_dfock_create(int n_bas, int n_alpha, int n_beta, int n_blocks, int n_aux, int n_problems,
                           const long long* seeds_host, double g, gscf_b200_dfock** out) {
  if (!out || n_bas < 1 || n_aux < 1 || n_problems < 1 || (n_blocks != 1 && n_blocks != 2) || !seeds_host ||
      n_alpha < 0 || n_beta < 0 || n_alpha > n_bas || n_beta > n_bas)
    return GSCF_B200_ERR_INVALID_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_last_error("no CUDA device: gscf_b200 has no CPU fallback");
    return GSCF_B200_ERR_CUDA;
  }
  gscf_b200_dfock* f = new gscf_b200_dfock();
  f->n = n_bas; f->ld = padded_ld(n_bas); f->nblk = n_blocks;
  f->occ[0] = n_alpha; f->occ[1] = n_blocks == 2 ? n_beta : n_alpha;
  f->omax = std::max(1, std::max(f->occ[0], f->occ[1]));
  f->naux = n_aux; f->P = n_problems;
  f->ldt = (long long)n_aux * f->ld;
  f->mstride = (long long)f->ld * n_bas;
  const size_t P = n_problems;
  long long* seeds_dev = nullptr;
  bool ok = cudaMalloc((void**)&f->S, P * f->mstride * sizeof(double)) == cudaSuccess &&
            cudaMalloc((void**)&f->H, P * f->mstride * sizeof(double)) == cudaSuccess &&
            cudaMalloc((void**)&f->Bt, P * f->ldt * n_bas * sizeof(double)) == cudaSuccess &&
            cudaMalloc((void**)&f->Y, P * f->nblk * f->ldt * f->omax * sizeof(double)) == cudaSuccess &&
            cudaMalloc((void**)&f->K, P * f->nblk * f->mstride * sizeof(double)) == cudaSuccess &&
            cudaMalloc((void**)&f->gamma, P * n_aux * sizeof(double)) == cudaSuccess &&
            cudaMalloc((void**)&f->epart, P * kEBlocks * 2 * sizeof(double)) == cudaSuccess &&
            cudaMalloc((void**)&seeds_dev, P * sizeof(long long)) == cudaSuccess;
  if (!ok) {
    set_last_error("dfock_create: cudaMalloc failed");
    gscf_b200_dfock_destroy(f);
    cudaFree(seeds_dev);
    return GSCF_B200_ERR_CUDA;
  }
  cudaMemset(f->S, 0, P * f->mstride * sizeof(double));
  cudaMemset(f->H, 0, P * f->mstride * sizeof(double));
  cudaMemset(f->Bt, 0, P * f->ldt * n_bas * sizeof(double));
  cudaMemset(f->Y, 0, P * f->nblk * f->ldt * f->omax * sizeof(double));
  cudaMemset(f->K, 0, P * f->nblk * f->mstride * sizeof(double));
  cudaMemcpy(seeds_dev, seeds_host, P * sizeof(long long), cudaMemcpyHostToDevice);
  int gx = (int)std::min<long long>(148 * 4, ceil_div_ll((long long)n_bas * n_bas, 256));
  // grid.z limited to 65535: chunk the problems
  for (int p0 = 0; p0 < n_problems; p0 += 32768) {
    const int cnt = std::min(32768, n_problems - p0);
    generate_kernel<<<dim3(gx, 2 + n_aux, cnt), 256>>>(f->S + (size_t)p0 * f->mstride, f->H + (size_t)p0 * f->mstride,
                                                       f->Bt + (size_t)p0 * f->ldt * n_bas, n_bas, f->ld, n_aux, n_alpha,
                                                       seeds_dev + p0, g);
    count_launch();
  }
  cudaError_t err = cudaDeviceSynchronize();
  cudaFree(seeds_dev);
  if (err != cudaSuccess) {
    set_last_error(std::string("dfock_create: ") + cudaGetErrorString(err));
    gscf_b200_dfock_destroy(f);
    return GSCF_B200_ERR_CUDA;
  }
  *out = f;
  return GSCF_B200_OK;
}

void gscf_b200_dfock_destroy(gscf_b200_dfock* f) {
  if (!f) return;
  cudaFree(f->S); cudaFree(f->H); cudaFree(f->Bt); cudaFree(f->Y); cudaFree(f->K);
  cudaFree(f->gamma); cudaFree(f->epart);
  delete f;
}

gscf_b200_fock_device_fn gscf_b200_dfock_callback(void) { return &dfock_build; }
const double* gscf_b200_dfock_overlap(gscf_b200_dfock* f) { return f ? f->S : nullptr; }
const double* gscf_b200_dfock_core(gscf_b200_dfock* f) { return f ? f->H : nullptr; }

}  // extern "C"
