// HBM-bound DIIS kernels: fused multi-dot, m-way weighted sum, device LU.  See stream.cuh.
//
// Roofline: both streaming kernels move (count+1) * 8 n^2 algorithmic bytes per problem
//  and are sized as grid = multiple of 148 SMs x resident CTAs; every thread keeps
// 2 x (count+1) independent 16-byte loads in flight.
#include "stream.cuh"

#include "gemm.cuh"
#include "smallsolve.hpp"

namespace gscfb {

constexpr int kStreamThreads = 256;

static inline int stream_blocks(long long work_items, int batch) {
  // work item = one pair of rows in one column; aim at ~8 items per thread, cap the grid at
  // 148 SMs x 8 resident CTAs.
  long long per_block = (long long)kStreamThreads * 8;
  long long blocks = ceil_div_ll(work_items, per_block);
  long long cap = (148LL * 8 + batch - 1) / batch;
  if (cap < 1) cap = 1;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ------------------------------------------------------------------------------------------
// multidot
// ------------------------------------------------------------------------------------------
template <int COUNT_MAX>
__global__ void __launch_bounds__(kStreamThreads)
multidot_partial_kernel(const double* __restrict__ e_new, long long pstride_new,
                        const double* __restrict__ ring, long long slot_stride,
                        long long pstride_ring, SlotList slots, int rows, int cols, int ld,
                        const int* __restrict__ batch_map, double* __restrict__ partial) {
  __shared__ double red[33];
  const int z = blockIdx.y;
  const long long p = batch_map ? batch_map[z] : z;
  const double* en = e_new + p * pstride_new;
  const double* rg = ring + p * pstride_ring;
  const int pairs = (rows + 1) >> 1;
  const long long total = (long long)pairs * cols;
  const bool vec_ok = ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(en) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(rg) & 15) == 0) && ((slot_stride & 1) == 0);
  double acc[COUNT_MAX];
#pragma unroll
  for (int j = 0; j < COUNT_MAX; ++j) acc[j] = 0.0;
  if (vec_ok && rows == ld) {
    // no padding rows: the matrices are flat arrays of double2.  Every CTA streams one contiguous chunk of each
    // operand (DRAM page locality), two items per thread and pass = 2 (count + 1) independent 16-byte loads in flight
    const long long per = (total + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per, hi = min(total, lo + per);
    const double2* en2 = reinterpret_cast<const double2*>(en);
    for (long long w = lo + threadIdx.x; w < hi; w += 2 * blockDim.x) {
      const long long w1 = w + blockDim.x;
      const bool has1 = w1 < hi;
      const double2 xa = en2[w];
      const double2 xb = has1 ? en2[w1] : make_double2(0.0, 0.0);
      double2 ya[COUNT_MAX], yb[COUNT_MAX];
#pragma unroll
      for (int j = 0; j < COUNT_MAX; ++j) {
        if (j < slots.n) {
          const double2* q = reinterpret_cast<const double2*>(rg + (size_t)slots.slot[j] * slot_stride);
          ya[j] = q[w];
          yb[j] = has1 ? q[w1] : make_double2(0.0, 0.0);
        }
      }
#pragma unroll
      for (int j = 0; j < COUNT_MAX; ++j) {
        if (j < slots.n) {
          acc[j] = fma(xa.x, ya[j].x, acc[j]);
          acc[j] = fma(xa.y, ya[j].y, acc[j]);
          acc[j] = fma(xb.x, yb[j].x, acc[j]);
          acc[j] = fma(xb.y, yb[j].y, acc[j]);
        }
      }
    }
  } else
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total;
       w += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(w / pairs);
    const int r = (int)(w - (long long)c * pairs) * 2;
    const size_t off = (size_t)c * ld + r;
    double x0, x1;
    const bool two = (r + 1 < rows);
    if (vec_ok && two) {
      const double2 v = *reinterpret_cast<const double2*>(en + off);
      x0 = v.x; x1 = v.y;
    } else {
      x0 = en[off];
      x1 = two ? en[off + 1] : 0.0;
    }
#pragma unroll
    for (int j = 0; j < COUNT_MAX; ++j) {
      if (j < slots.n) {
        const double* q = rg + (size_t)slots.slot[j] * slot_stride + off;
        double y0, y1;
        if (vec_ok && two) {
          const double2 v = *reinterpret_cast<const double2*>(q);
          y0 = v.x; y1 = v.y;
        } else {
          y0 = q[0];
          y1 = two ? q[1] : 0.0;
        }
        acc[j] = fma(x0, y0, acc[j]);
        acc[j] = fma(x1, y1, acc[j]);
      }
    }
  }
  // per-warp sums into shared memory, one barrier, then thread j adds the warps in order (deterministic)
  __shared__ double swarp[kStreamThreads / 32][COUNT_MAX];
  (void)red;
#pragma unroll
  for (int j = 0; j < COUNT_MAX; ++j) {
    const double s = warp_sum(acc[j]);
    if ((threadIdx.x & 31) == 0) swarp[threadIdx.x >> 5][j] = s;
  }
  __syncthreads();
  if (threadIdx.x < slots.n) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kStreamThreads / 32; ++w) s += swarp[w][threadIdx.x];
    partial[((size_t)z * gridDim.x + blockIdx.x) * kMaxHistory + threadIdx.x] = s;
  }
}

__global__ void multidot_final_kernel(const double* __restrict__ partial, int nblocks, int count,
                                      const int* __restrict__ batch_map, double* __restrict__ out,
                                      int out_stride) {
  // one warp per dot product: lane l adds blocks l, l + 32, ... in order, then a butterfly (fixed order: deterministic)
  const int z = blockIdx.x;
  const int j = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (j >= count) return;
  const long long p = batch_map ? batch_map[z] : z;
  double s = 0.0;
  for (int b = lane; b < nblocks; b += 32) s += partial[((size_t)z * nblocks + b) * kMaxHistory + j];
  s = warp_sum(s);
  if (lane == 0) out[p * out_stride + j] = s;
}

size_t multidot_worksize(int count, int rows, int cols, int batch) {
  (void)count;
  const int blocks = stream_blocks((long long)((rows + 1) / 2) * cols, batch);
  return (size_t)blocks * batch * kMaxHistory * sizeof(double);
}

int multidot(const double* e_new, long long pstride_new, const double* ring, long long slot_stride,
             long long pstride_ring, const SlotList& slots, int rows, int cols, int ld, int batch,
             const int* batch_map, double* out_dev, int out_stride, double* work,
             cudaStream_t st) {
  if (slots.n <= 0 || batch <= 0) return GSCF_B200_OK;
  if (slots.n > kMaxHistory) return GSCF_B200_ERR_INVALID_ARG;
  const int blocks = stream_blocks((long long)((rows + 1) / 2) * cols, batch);
  dim3 grid(blocks, batch);
  if (slots.n <= 4)
    multidot_partial_kernel<4><<<grid, kStreamThreads, 0, st>>>(
        e_new, pstride_new, ring, slot_stride, pstride_ring, slots, rows, cols, ld, batch_map, work);
  else if (slots.n <= 8)
    multidot_partial_kernel<8><<<grid, kStreamThreads, 0, st>>>(
        e_new, pstride_new, ring, slot_stride, pstride_ring, slots, rows, cols, ld, batch_map, work);
  else
    multidot_partial_kernel<16><<<grid, kStreamThreads, 0, st>>>(
        e_new, pstride_new, ring, slot_stride, pstride_ring, slots, rows, cols, ld, batch_map, work);
  count_launch();
  GSCF_CHECK_LAUNCH();
  multidot_final_kernel<<<batch, 32 * slots.n, 0, st>>>(work, blocks, slots.n, batch_map, out_dev, out_stride);
  count_launch();
  GSCF_CHECK_LAUNCH();
  return GSCF_B200_OK;
}

// ------------------------------------------------------------------------------------------
// axpby_n
// ------------------------------------------------------------------------------------------
template <int COUNT_MAX>
__global__ void __launch_bounds__(kStreamThreads)
axpby_n_kernel(double* __restrict__ out, long long pstride_out, const double* __restrict__ ring,
               long long slot_stride, long long pstride_ring, SlotList slots,
               const double* __restrict__ coeff, int coeff_stride, int rows, int cols, int ld,
               const int* __restrict__ batch_map) {
  const int z = blockIdx.y;
  const long long p = batch_map ? batch_map[z] : z;
  double* o = out + p * pstride_out;
  const double* rg = ring + p * pstride_ring;
  double cf[COUNT_MAX];
#pragma unroll
  for (int j = 0; j < COUNT_MAX; ++j) cf[j] = j < slots.n ? coeff[p * coeff_stride + j] : 0.0;
  const int pairs = (rows + 1) >> 1;
  const long long total = (long long)pairs * cols;
  const bool vec_ok = ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(rg) & 15) == 0) && ((slot_stride & 1) == 0);
  if (vec_ok && rows == ld) {
    // flat double2 arrays, one contiguous chunk per CTA, two items per thread and pass (see multidot)
    const long long per = (total + gridDim.x - 1) / gridDim.x;
    const long long lo = (long long)blockIdx.x * per, hi = min(total, lo + per);
    double2* o2 = reinterpret_cast<double2*>(o);
    for (long long w = lo + threadIdx.x; w < hi; w += 2 * blockDim.x) {
      const long long w1 = w + blockDim.x;
      const bool has1 = w1 < hi;
      double2 ya[COUNT_MAX], yb[COUNT_MAX];
#pragma unroll
      for (int j = 0; j < COUNT_MAX; ++j) {
        if (j < slots.n) {
          const double2* q = reinterpret_cast<const double2*>(rg + (size_t)slots.slot[j] * slot_stride);
          ya[j] = q[w];
          yb[j] = has1 ? q[w1] : make_double2(0.0, 0.0);
        }
      }
      double2 sa = make_double2(0.0, 0.0), sb = make_double2(0.0, 0.0);
#pragma unroll
      for (int j = 0; j < COUNT_MAX; ++j) {
        if (j < slots.n) {  // same left-to-right order as the general path below
          sa.x = (j == 0) ? cf[0] * ya[0].x : fma(cf[j], ya[j].x, sa.x);
          sa.y = (j == 0) ? cf[0] * ya[0].y : fma(cf[j], ya[j].y, sa.y);
          sb.x = (j == 0) ? cf[0] * yb[0].x : fma(cf[j], yb[j].x, sb.x);
          sb.y = (j == 0) ? cf[0] * yb[0].y : fma(cf[j], yb[j].y, sb.y);
        }
      }
      o2[w] = sa;
      if (has1) o2[w1] = sb;
    }
  } else
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total;
       w += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(w / pairs);
    const int r = (int)(w - (long long)c * pairs) * 2;
    const size_t off = (size_t)c * ld + r;
    const bool two = (r + 1 < rows);
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int j = 0; j < COUNT_MAX; ++j) {
      if (j < slots.n) {
        const double* q = rg + (size_t)slots.slot[j] * slot_stride + off;
        double y0, y1;
        if (vec_ok && two) {
          const double2 v = *reinterpret_cast<const double2*>(q);
          y0 = v.x; y1 = v.y;
        } else {
          y0 = q[0];
          y1 = two ? q[1] : 0.0;
        }
        // first term is a plain product: c_0 * F_0 + c_1 * F_1 + ... in ring order, matching
        // the reference's left-to-right lazy sum (PulayDiisScf.hh:419-428)
        s0 = (j == 0) ? cf[0] * y0 : fma(cf[j], y0, s0);
        s1 = (j == 0) ? cf[0] * y1 : fma(cf[j], y1, s1);
      }
    }
    if (vec_ok && two) {
      *reinterpret_cast<double2*>(o + off) = make_double2(s0, s1);
    } else {
      o[off] = s0;
      if (two) o[off + 1] = s1;
    }
  }
}

int axpby_n(double* out, long long pstride_out, const double* ring, long long slot_stride,
            long long pstride_ring, const SlotList& slots, const double* coeff_dev,
            int coeff_stride, int rows, int cols, int ld, int batch, const int* batch_map,
            cudaStream_t st) {
  if (slots.n <= 0 || batch <= 0) return GSCF_B200_OK;
  if (slots.n > kMaxHistory) return GSCF_B200_ERR_INVALID_ARG;
  const int blocks = stream_blocks((long long)((rows + 1) / 2) * cols, batch);
  dim3 grid(blocks, batch);
  if (slots.n <= 4)
    axpby_n_kernel<4><<<grid, kStreamThreads, 0, st>>>(out, pstride_out, ring, slot_stride,
                                                       pstride_ring, slots, coeff_dev,
                                                       coeff_stride, rows, cols, ld, batch_map);
  else if (slots.n <= 8)
    axpby_n_kernel<8><<<grid, kStreamThreads, 0, st>>>(out, pstride_out, ring, slot_stride,
                                                       pstride_ring, slots, coeff_dev,
                                                       coeff_stride, rows, cols, ld, batch_map);
  else
    axpby_n_kernel<16><<<grid, kStreamThreads, 0, st>>>(out, pstride_out, ring, slot_stride,
                                                        pstride_ring, slots, coeff_dev,
                                                        coeff_stride, rows, cols, ld, batch_map);
  count_launch();
  GSCF_CHECK_LAUNCH();
  return GSCF_B200_OK;
}

// ------------------------------------------------------------------------------------------
// DIIS coefficient solve
// ------------------------------------------------------------------------------------------
__global__ void diis_solve_kernel(const double* __restrict__ gram, int gram_ld,
                                  long long gram_stride, SlotList order,
                                  double* __restrict__ coeff, int coeff_stride,
                                  int* __restrict__ fail, int batch,
                                  const int* __restrict__ batch_map) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= batch) return;
  const long long p = batch_map ? batch_map[z] : z;
  double c[kMaxDiis];
  const bool ok = diis_build_and_solve(gram + p * gram_stride, gram_ld, order.slot, order.n, c);
  for (int i = 0; i < order.n; ++i) coeff[p * coeff_stride + i] = c[i];
  if (fail) fail[p] = ok ? 0 : 1;
}

int diis_solve(const double* gram_dev, int gram_ld, long long gram_stride, const SlotList& order,
               double* coeff_dev, int coeff_stride, int* fail_dev, int batch,
               const int* batch_map, cudaStream_t st) {
  if (batch <= 0) return GSCF_B200_OK;
  if (order.n > kMaxHistory) return GSCF_B200_ERR_INVALID_ARG;
  diis_solve_kernel<<<ceil_div(batch, 64), 64, 0, st>>>(gram_dev, gram_ld, gram_stride, order,
                                                        coeff_dev, coeff_stride, fail_dev, batch,
                                                        batch_map);
  count_launch();
  GSCF_CHECK_LAUNCH();
  return GSCF_B200_OK;
}

int diis_solve_host(const double* gram, int gram_ld, const int* order, int k, double* coeff) {
  return diis_build_and_solve(gram, gram_ld, order, k, coeff) ? GSCF_B200_OK
                                                               : GSCF_B200_ERR_DIIS_STEP;
}

// ------------------------------------------------------------------------------------------
// element-wise helpers
// ------------------------------------------------------------------------------------------
__global__ void scale_copy_kernel(double* __restrict__ dst, int ldd, long long sdst,
                                  const double* __restrict__ src, int lds, long long ssrc,
                                  double a, int rows, int cols, const int* __restrict__ batch_map) {
  const long long p = batch_map ? batch_map[blockIdx.y] : blockIdx.y;
  double* d = dst + p * sdst;
  const double* s = src + p * ssrc;
  const long long total = (long long)rows * cols;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < total;
       w += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(w / rows);
    const int r = (int)(w - (long long)c * rows);
    d[(size_t)c * ldd + r] = a * s[(size_t)c * lds + r];
  }
}

int scale_copy(double* dst, int ldd, long long sdst, const double* src, int lds, long long ssrc,
               double a, int rows, int cols, int batch, const int* batch_map, cudaStream_t st) {
  if (batch <= 0 || rows <= 0 || cols <= 0) return GSCF_B200_OK;
  const int blocks = stream_blocks((long long)rows * cols / 2 + 1, batch);
  scale_copy_kernel<<<dim3(blocks, batch), kStreamThreads, 0, st>>>(dst, ldd, sdst, src, lds, ssrc,
                                                                    a, rows, cols, batch_map);
  count_launch();
  GSCF_CHECK_LAUNCH();
  return GSCF_B200_OK;
}

__global__ void scale_columns_pow_kernel(double* __restrict__ dst, int ldd, long long sdst,
                                         const double* __restrict__ src, int lds, long long ssrc,
                                         const double* __restrict__ w, long long sw, int mode,
                                         int rows, int cols) {
  const long long p = blockIdx.y;
  double* d = dst + p * sdst;
  const double* s = src + p * ssrc;
  const double* wp = w + p * sw;
  const long long total = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / rows);
    const int r = (int)(i - (long long)c * rows);
    const double x = wp[c];
    const double f = mode == 0 ? 1.0 / sqrt(x) : 1.0 / sqrt(sqrt(x));
    d[(size_t)c * ldd + r] = s[(size_t)c * lds + r] * f;
  }
}

int scale_columns_pow(double* dst, int ldd, long long sdst, const double* src, int lds,
                      long long ssrc, const double* w, long long sw, int mode, int rows, int cols,
                      int batch, cudaStream_t st) {
  if (batch <= 0 || rows <= 0 || cols <= 0) return GSCF_B200_OK;
  const int blocks = stream_blocks((long long)rows * cols / 2 + 1, batch);
  scale_columns_pow_kernel<<<dim3(blocks, batch), kStreamThreads, 0, st>>>(
      dst, ldd, sdst, src, lds, ssrc, w, sw, mode, rows, cols);
  count_launch();
  GSCF_CHECK_LAUNCH();
  return GSCF_B200_OK;
}

int fill_zero(double* dst, size_t n, cudaStream_t st) {
  GSCF_CUDA_TRY(cudaMemsetAsync(dst, 0, n * sizeof(double), st));
  return GSCF_B200_OK;
}

}  // namespace gscfb
