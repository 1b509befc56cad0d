// Small dense helpers shared by host code and device kernels (compiled for both).
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define GSCF_HD __host__ __device__
#else
#define GSCF_HD
#endif

namespace gscfb {

constexpr int kMaxDiis = 16;  // must match kMaxHistory

// Build the bordered DIIS system of PulayDiisScf.hh:326-356 from a Gram table and solve it.
//   gram[a*gram_ld + b] = <e_slot_a, e_slot_b>,  order[i] = slot of the i-th oldest error.
//   B(i,j) = gram[order[i]][order[j]] (i >= j entry used for both triangles, as the reference
//   fills B(i,i-j) = B(i-j,i) from one stored overlap), border -1, corner 0, rhs = (0..0,-1).
// LU with partial pivoting, row interchanges, no scaling/regularisation (dgesv semantics,
// PulayDiisScf.hh:380-392).  Returns false when a pivot is exactly zero or the solution is not
// finite (-> ExcDiisStepFailed).  coeff receives the first k entries (PulayDiisScf.hh:389).
GSCF_HD inline bool diis_build_and_solve(const double* gram, int gram_ld, const int* order, int k,
                                         double* coeff) {
  if (k == 0) return true;
  if (k == 1) {
    coeff[0] = 1.0;
    return true;
  }
  const int nn = k + 1;
  double M[(kMaxDiis + 1) * (kMaxDiis + 1)];
  double x[kMaxDiis + 1];
  for (int i = 0; i < k; ++i) {
    for (int j = 0; j <= i; ++j) {
      // overlap computed when error i (the newer one) arrived: row = newer slot
      const double v = gram[order[i] * gram_ld + order[j]];
      M[i * nn + j] = v;
      M[j * nn + i] = v;
    }
    M[i * nn + k] = -1.0;
    M[k * nn + i] = -1.0;
    x[i] = 0.0;
  }
  M[k * nn + k] = 0.0;
  x[k] = -1.0;
  // right-looking LU, column pivot search like dgetf2 (first maximal |entry|)
  for (int c = 0; c < nn; ++c) {
    int piv = c;
    double best = fabs(M[c * nn + c]);
    for (int r = c + 1; r < nn; ++r) {
      const double v = fabs(M[r * nn + c]);
      if (v > best) {
        best = v;
        piv = r;
      }
    }
    if (!(best > 0.0)) return false;
    if (piv != c) {
      for (int j = 0; j < nn; ++j) {
        const double t = M[c * nn + j];
        M[c * nn + j] = M[piv * nn + j];
        M[piv * nn + j] = t;
      }
      const double t = x[c];
      x[c] = x[piv];
      x[piv] = t;
    }
    const double inv = 1.0 / M[c * nn + c];
    for (int r = c + 1; r < nn; ++r) {
      const double l = M[r * nn + c] * inv;
      M[r * nn + c] = l;
      for (int j = c + 1; j < nn; ++j) M[r * nn + j] -= l * M[c * nn + j];
      x[r] -= l * x[c];
    }
  }
  for (int r = nn - 1; r >= 0; --r) {
    double s = x[r];
    for (int j = r + 1; j < nn; ++j) s -= M[r * nn + j] * x[j];
    x[r] = s / M[r * nn + r];
  }
  bool ok = true;
  for (int i = 0; i < k; ++i) {
    coeff[i] = x[i];
    if (!isfinite(x[i])) ok = false;
  }
  return ok;
}

// TruncatedOptDampScf.hh:370-394, exact control flow (evaluated on the host from
// device-reduced scalars so that the == 1 tests downstream are exact).
GSCF_HD inline double compute_damping_coeff(double oda_s, double oda_c, double min_fock_prefactor) {
  const double ignore_threshold = 1e-14;
  if (fabs(oda_s) < ignore_threshold || fabs(oda_c) < ignore_threshold) return 1.0;
  if (oda_s >= 0) return 1.0;
  const double ndc = (2. * oda_c <= -oda_s) ? 1.0 : -oda_s / (2 * oda_c);
  if (ndc < min_fock_prefactor) return min_fock_prefactor;
  if (ndc > 1 - min_fock_prefactor) return 1.0;
  return ndc;
}

}  // namespace gscfb
