// C-ABI wrappers for the building blocks + library bookkeeping (include/gscf_b200.h section 1).
#include "common.cuh"
#include "eig.cuh"
#include "gemm.cuh"
#include "smallsolve.hpp"
#include "stream.cuh"

#include <cstring>
#include <string>

namespace gscfb {

static thread_local std::string g_last_error;
void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* last_error_cstr() { return g_last_error.c_str(); }
uint64_t launch_count_value();
uint64_t tensor_map_launch_count_value();

static int have_device() {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_last_error("no CUDA device: gscf_b200 has no CPU fallback");
    return GSCF_B200_ERR_CUDA;
  }
  return GSCF_B200_OK;
}

static SlotList make_slots(const int* host, int count) {
  SlotList s{};
  s.n = count;
  for (int i = 0; i < count && i < kMaxHistory; ++i) s.slot[i] = host[i];
  return s;
}

// ---- peak micro-benchmarks ---------------------------------------------------------------------
__global__ void dmma_peak_kernel(double* out, int iters) {
  double c[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

__global__ void dfma_peak_kernel(double* out, int iters) {
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}

__global__ void copy_peak_kernel(double2* __restrict__ dst, const double2* __restrict__ src, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

}  // namespace gscfb

using namespace gscfb;

extern "C" {

const char* gscf_b200_last_error(void) { return last_error_cstr(); }
const char* gscf_b200_version_string(void) { return "gscf-b200 0.1.0 (sm_100a)"; }
uint64_t gscf_b200_launch_count(void) { return launch_count_value(); }
uint64_t gscf_b200_tensor_map_launch_count(void) { return tensor_map_launch_count_value(); }

int gscf_b200_dgemm(int transa, int transb, int m, int n, int k, double alpha, const double* A, int lda,
                    long long strideA, const double* B, int ldb, long long strideB, double beta, double* C,
                    int ldc, long long strideC, int batch, void* stream) {
  GSCF_TRY(have_device());
  if (m < 0 || n < 0 || k < 0 || batch < 0) return GSCF_B200_ERR_INVALID_ARG;
  return gemm_batched(transa != 0, transb != 0, m, n, k, alpha, A, lda, strideA, B, ldb, strideB, beta, C, ldc,
                      strideC, batch, nullptr, (cudaStream_t)stream);
}

size_t gscf_b200_syev_worksize(int n, int batch, int method) { return syev_worksize(n, batch, method); }

int gscf_b200_syev(int method, int n, double* A, int lda, long long strideA, double* w, long long stride_w,
                   double* Z, int ldz, long long strideZ, int batch, void* work, size_t work_bytes, int* info_dev,
                   void* stream) {
  GSCF_TRY(have_device());
  const int m = resolve_eig_method(method, n);
  if (m < 0) return GSCF_B200_ERR_INVALID_ARG;
  if (m == GSCF_B200_EIG_TRIDIAG_DC)  // the large path works on full symmetric storage
    GSCF_TRY(symmetrise_from_lower(A, lda, strideA, n, batch, nullptr, (cudaStream_t)stream));
  return syev_batched(m, n, A, lda, strideA, w, stride_w, Z, ldz, strideZ, batch, nullptr, work, work_bytes,
                      info_dev, (cudaStream_t)stream);
}

int gscf_b200_syevx(int method, int n, int n_eigenpairs, double* A, int lda, long long strideA, double* w,
                    long long stride_w, double* Z, int ldz, long long strideZ, int batch, void* work, size_t work_bytes,
                    int* info_dev, void* stream) {
  GSCF_TRY(have_device());
  if (n_eigenpairs < 1 || n_eigenpairs > n) {
    set_last_error("syevx: n_eigenpairs must be in [1, n]");
    return GSCF_B200_ERR_INVALID_ARG;
  }
  const int m = resolve_eig_method(method, n);
  if (m < 0) return GSCF_B200_ERR_INVALID_ARG;
  if (m == GSCF_B200_EIG_TRIDIAG_DC)
    GSCF_TRY(symmetrise_from_lower(A, lda, strideA, n, batch, nullptr, (cudaStream_t)stream));
  return syev_batched(m, n, A, lda, strideA, w, stride_w, Z, ldz, strideZ, batch, nullptr, work, work_bytes,
                      info_dev, (cudaStream_t)stream, n_eigenpairs);
}

size_t gscf_b200_frob_multidot_worksize(int count, int rows, int cols, int batch) {
  return multidot_worksize(count, rows, cols, batch);
}

int gscf_b200_frob_multidot(const double* e_new, const double* ring, long long slot_stride, const int* slots_host,
                            int count, int rows, int cols, int ld, long long problem_stride_new,
                            long long problem_stride_ring, int batch, double* out_dev, int out_stride, void* work,
                            size_t work_bytes, void* stream) {
  GSCF_TRY(have_device());
  if (count < 0 || count > kMaxHistory || !slots_host) return GSCF_B200_ERR_INVALID_ARG;
  if (work_bytes < multidot_worksize(count, rows, cols, batch)) {
    set_last_error("frob_multidot: workspace too small");
    return GSCF_B200_ERR_INVALID_ARG;
  }
  return multidot(e_new, problem_stride_new, ring, slot_stride, problem_stride_ring, make_slots(slots_host, count),
                  rows, cols, ld, batch, nullptr, out_dev, out_stride, (double*)work, (cudaStream_t)stream);
}

int gscf_b200_axpby_n(double* out, long long problem_stride_out, const double* ring, long long slot_stride,
                      long long problem_stride_ring, const int* slots_host, int count, const double* coeff_dev,
                      int coeff_stride, int rows, int cols, int ld, int batch, void* stream) {
  GSCF_TRY(have_device());
  if (count < 0 || count > kMaxHistory || !slots_host) return GSCF_B200_ERR_INVALID_ARG;
  return axpby_n(out, problem_stride_out, ring, slot_stride, problem_stride_ring, make_slots(slots_host, count),
                 coeff_dev, coeff_stride, rows, cols, ld, batch, nullptr, (cudaStream_t)stream);
}

int gscf_b200_diis_solve(const double* gram_dev, int gram_ld, long long gram_stride, const int* order_host, int k,
                         double* coeff_dev, int coeff_stride, int* fail_dev, int batch, void* stream) {
  GSCF_TRY(have_device());
  if (k < 0 || k > kMaxHistory || !order_host) return GSCF_B200_ERR_INVALID_ARG;
  return diis_solve(gram_dev, gram_ld, gram_stride, make_slots(order_host, k), coeff_dev, coeff_stride, fail_dev,
                    batch, nullptr, (cudaStream_t)stream);
}

int gscf_b200_density(int n, int n_occ, const double* C, int ldc, long long strideC, double* D, int ldd,
                      long long strideD, int batch, void* stream) {
  GSCF_TRY(have_device());
  return gemm_batched(false, true, n, n, n_occ, 1.0, C, ldc, strideC, C, ldc, strideC, 0.0, D, ldd, strideD, batch,
                      nullptr, (cudaStream_t)stream);
}

size_t gscf_b200_pulay_error_worksize(int n, int n_occ, int batch) {
  return pulay_error_work_doubles(n, n_occ) * (size_t)batch * sizeof(double);
}

int gscf_b200_pulay_error(int n, int n_occ, double fac, const double* S, int lds, long long strideS,
                          const double* F, int ldf, long long strideF, const double* C, int ldc, long long strideC,
                          double* E, int lde, long long strideE, double* work, long long stride_work, int batch,
                          void* stream) {
  GSCF_TRY(have_device());
  if (n < 1 || n_occ < 0 || !S || !F || !C || !E || !work || batch < 0) return GSCF_B200_ERR_INVALID_ARG;
  if (batch > 1 && stride_work < (long long)pulay_error_work_doubles(n, n_occ)) {
    set_last_error("pulay_error: stride_work is smaller than gscf_b200_pulay_error_worksize(n, n_occ, 1) / 8");
    return GSCF_B200_ERR_INVALID_ARG;
  }
  PulayErrorArgs a{};
  a.n = n; a.o = n_occ; a.fac = fac;
  a.S = S; a.lds = lds; a.sS = strideS;
  a.F = F; a.ldf = ldf; a.sF = strideF;
  a.C = C; a.ldc = ldc; a.sC = strideC;
  a.E = E; a.lde = lde; a.sE = strideE;
  a.work = work; a.swork = stride_work; a.batch = batch; a.batch_map = nullptr; a.entries = 0;
  return pulay_error_fused(a, (cudaStream_t)stream);
}

int gscf_b200_copy_matrix(void* dst, int ld_dst, int dst_loc, const void* src, int ld_src, int src_loc, int rows,
                          int cols) {
  GSCF_TRY(have_device());
  cudaMemcpyKind kind = cudaMemcpyDefault;
  if (dst_loc == GSCF_B200_HOST && src_loc == GSCF_B200_DEVICE) kind = cudaMemcpyDeviceToHost;
  else if (dst_loc == GSCF_B200_DEVICE && src_loc == GSCF_B200_HOST) kind = cudaMemcpyHostToDevice;
  else if (dst_loc == GSCF_B200_DEVICE && src_loc == GSCF_B200_DEVICE) kind = cudaMemcpyDeviceToDevice;
  else kind = cudaMemcpyHostToHost;
  GSCF_CUDA_TRY(cudaMemcpy2D(dst, (size_t)ld_dst * sizeof(double), src, (size_t)ld_src * sizeof(double),
                             (size_t)rows * sizeof(double), cols, kind));
  return GSCF_B200_OK;
}

double gscf_b200_compute_damping_coeff(double oda_s, double oda_c, double min_fock_prefactor) {
  return compute_damping_coeff(oda_s, oda_c, min_fock_prefactor);
}

int gscf_b200_measure_peaks(double* dmma_tflops, double* dfma_tflops, double* copy_gbs) {
  GSCF_TRY(have_device());
  cudaDeviceProp prop;
  GSCF_CUDA_TRY(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double* out = nullptr;
  GSCF_CUDA_TRY(cudaMalloc((void**)&out, 64));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float ms = 0.f;
  const int blocks = sms * 4, threads = 256, iters = 20000;
  // DMMA: 8 independent accumulators per warp; 8*4*8*2 = 512 flop per dmma884
  dmma_peak_kernel<<<blocks, threads>>>(out, 100);
  cudaDeviceSynchronize();
  double best = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    dmma_peak_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    const double flops = (double)blocks * (threads / 32) * iters * 8.0 * 512.0;
    best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  if (dmma_tflops) *dmma_tflops = best;
  dfma_peak_kernel<<<blocks, threads>>>(out, 100);
  cudaDeviceSynchronize();
  best = 0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    dfma_peak_kernel<<<blocks, threads>>>(out, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    const double flops = (double)blocks * threads * iters * 16.0 * 2.0;
    best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  if (dfma_tflops) *dfma_tflops = best;
  const size_t nbytes = (size_t)1 << 30;
  double2 *src = nullptr, *dst = nullptr;
  best = 0;
  if (cudaMalloc((void**)&src, nbytes) == cudaSuccess && cudaMalloc((void**)&dst, nbytes) == cudaSuccess) {
    cudaMemset(src, 1, nbytes);
    copy_peak_kernel<<<sms * 8, 512>>>(dst, src, nbytes / 16);
    cudaDeviceSynchronize();
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(a);
      copy_peak_kernel<<<sms * 8, 512>>>(dst, src, nbytes / 16);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      cudaEventElapsedTime(&ms, a, b);
      best = std::max(best, 2.0 * nbytes / (ms * 1e-3) / 1e9);
    }
  }
  cudaFree(src);
  cudaFree(dst);
  if (copy_gbs) *copy_gbs = best;
  cudaFree(out);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  GSCF_CUDA_TRY(cudaGetLastError());
  return GSCF_B200_OK;
}

}  // extern "C"
