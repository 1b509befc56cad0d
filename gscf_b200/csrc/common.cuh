// gscf-b200: shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/gscf_b200.h"

namespace gscfb {

// ---- error plumbing -------------------------------------------------------------------
// Every C-ABI entry point returns an int status (include/gscf_b200.h); CUDA failures are
// latched into a thread-local message readable through gscf_b200_last_error().
void set_last_error(const std::string& msg);
const char* last_error_cstr();

#define GSCF_CUDA_TRY(expr)                                                              \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::gscfb::set_last_error(std::string(#expr) + " -> " + cudaGetErrorString(_e) +     \
                              " at " + __FILE__ + ":" + std::to_string(__LINE__));      \
      return GSCF_B200_ERR_CUDA;                                                         \
    }                                                                                    \
  } while (0)

#define GSCF_TRY(expr)                      \
  do {                                      \
    int _s = (expr);                        \
    if (_s != GSCF_B200_OK) return _s;      \
  } while (0)

#define GSCF_CHECK_LAUNCH() GSCF_CUDA_TRY(cudaGetLastError())

// Launch accounting for bench.py's gpu_launches claim.
void count_launch(int n = 1);

// Optional sub-phase timing (CUDA events on `st`); a no-op unless an engine enabled it.
enum PhaseId { PH_TRIDIAG = 8, PH_STEDC = 9, PH_BACKTRANSFORM = 10, PH_SYMV_SAMPLE = 11, PH_ORTHO_GEMM = 12,
               PH_TRAILING_GEMM = 13 };
void phase_mark(int id, bool begin, cudaStream_t st);

// Kernel function attributes (opt-in shared memory, cluster sizes) are PER DEVICE: a `static bool configured` would leave
// the second GPU a process drives unconfigured.  One bit per device ordinal; safe under concurrent host threads (a race
// only repeats the idempotent cudaFuncSetAttribute).
struct PerDeviceOnce {
  std::atomic<unsigned long long> mask{0ull};
  // true: the caller still has to configure the current device (returned in dev)
  bool need(int& dev) {
    dev = 0;
    cudaGetDevice(&dev);
    return dev >= 64 || !(mask.load(std::memory_order_acquire) >> dev & 1ull);
  }
  void done(int dev) {
    if (dev < 64) mask.fetch_or(1ull << dev, std::memory_order_release);
  }
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
inline int ceil_div(int x, int m) { return (x + m - 1) / m; }
inline long long ceil_div_ll(long long x, long long m) { return (x + m - 1) / m; }

// Leading dimension used for every n x n matrix the engine owns: multiple of 16 doubles
// (128 B) so that every column starts on a full L2 line and 16-byte cp.async is legal.
inline int padded_ld(int n) { return round_up(n < 1 ? 1 : n, 16); }

// ---- device helpers ---------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum, result valid in every thread.  `red` is >= 33 doubles of shared memory.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarp = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double t = lane < nwarp ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// cp.async (LDGSTS) with zero fill: copies `src_bytes` (0..16) and zero-fills up to 16.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- TMA bulk copies (cp.async.bulk, SASS UBLKCP) completing on an mbarrier -----------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; signals complete_tx on `bar`
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// One FP64 tensor-core tile: D(8x8) += A(8x4) * B(4x8).  SASS: DMMA.8x8x4.
// Fragment ownership (lane = 0..31):  a = A[lane/4][lane%4],  b = B[lane%4][lane/4],
// c0,c1 = C[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile(
      "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

}  // namespace gscfb
