// HBM-bound streaming kernels of the DIIS / damping steps and small device-side solves.
#pragma once
#include "common.cuh"

namespace gscfb {

constexpr int kMaxHistory = 16;  // ring capacity limit (DIIS depth <= 16)

struct SlotList {
  int n;
  int slot[kMaxHistory];
};

// out[p][j] = <Enew_p, ring_p[slot j]>_F   (PulayDiisScf.hh:440-463, fused: Enew read once)
size_t multidot_worksize(int count, int rows, int cols, int batch);
int multidot(const double* e_new, long long pstride_new, const double* ring, long long slot_stride,
             long long pstride_ring, const SlotList& slots, int rows, int cols, int ld, int batch,
             const int* batch_map, double* out_dev, int out_stride, double* work,
             cudaStream_t st);

// out_p = sum_i coeff[p][i] * ring_p[slot i]   (OperatorLinearCombination.hh:26-47)
int axpby_n(double* out, long long pstride_out, const double* ring, long long slot_stride,
            long long pstride_ring, const SlotList& slots, const double* coeff_dev,
            int coeff_stride, int rows, int cols, int ld, int batch, const int* batch_map,
            cudaStream_t st);

// DIIS bordered system + LU with partial pivoting (PulayDiisScf.hh:326-393)
int diis_solve(const double* gram_dev, int gram_ld, long long gram_stride, const SlotList& order,
               double* coeff_dev, int coeff_stride, int* fail_dev, int batch,
               const int* batch_map, cudaStream_t st);

// y = a*x (+ b*y) element-wise on rows x cols (ld) matrices, batched
int scale_copy(double* dst, int ldd, long long sdst, const double* src, int lds, long long ssrc,
               double a, int rows, int cols, int batch, const int* batch_map, cudaStream_t st);
// dst columns scaled: dst[:, j] = src[:, j] * f(w[j]);  mode 0: w^-1/2, 1: w^-1/4
int scale_columns_pow(double* dst, int ldd, long long sdst, const double* src, int lds,
                      long long ssrc, const double* w, long long sw, int mode, int rows, int cols,
                      int batch, cudaStream_t st);
int fill_zero(double* dst, size_t n, cudaStream_t st);

// Host-side restatement used by the single-problem path and by tests: identical LU.
int diis_solve_host(const double* gram, int gram_ld, const int* order, int k, double* coeff);

}  // namespace gscfb
