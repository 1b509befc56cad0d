// Numerical core of the divide & conquer tridiagonal eigensolver, shared by the device kernels
// (stedc.cu) and the host unit test (tests/cpp/dc_host_test.cc): deflation and the secular
// equation root finder of the rank-one update  diag(d) + rho z z^T  (Cuppen / Gu-Eisenstat; the
// algorithm family behind LAPACK dstedc, which is what the reference reaches through
// lazyten -> armadillo for ScfBase::update_eigenpairs, ScfBase.hh:291-339).
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define GSCF_HD __host__ __device__
#else
#define GSCF_HD
#endif

namespace gscfb {

constexpr double kDcEps = 1.1102230246251565e-16;  // 2^-53, LAPACK dlamch('E')

// Root i (0-based) of  g(lam) = 1 + rho * sum_j z_j^2 / (d_j - lam),  d strictly ascending,
// rho > 0, all z_j != 0.  Root i lies in (d_i, d_{i+1}) (the last one in (d_{K-1}, d_{K-1} +
// rho |z|^2]).  The iteration runs in tau = lam - d_origin with the origin at the nearer pole so
// that delta_j = (d_j - d_origin) - tau = d_j - lam is obtained to high *relative* accuracy (the
// property the Gu-Eisenstat orthogonality argument needs).  Middle-way rational interpolation
// (two-pole model of the left/right partial sums) safeguarded by bisection on a bracket.
// delta[j * dstride] receives d_j - lam.  Returns lam.
GSCF_HD inline double secular_root(int K, int i, const double* d, const double* z, double rho,
                                   double* delta, long long dstride, int* iters_out) {
  if (iters_out) *iters_out = 0;
  if (K == 1) {
    const double tau = rho * z[0] * z[0];
    delta[0] = -tau;
    return d[0] + tau;
  }
  const bool last = (i == K - 1);
  int io;
  double lo, hi, tau;
  if (!last) {
    const double del = d[i + 1] - d[i];
    const double mid = 0.5 * del;
    double g = 1.0;
    for (int j = 0; j < K; ++j) g += rho * z[j] * z[j] / ((d[j] - d[i]) - mid);
    if (g >= 0.0) {
      io = i; lo = 0.0; hi = mid; tau = mid;
    } else {
      io = i + 1; lo = -mid; hi = 0.0; tau = -mid;
    }
  } else {
    double zz = 0.0;
    for (int j = 0; j < K; ++j) zz += z[j] * z[j];
    io = K - 1; lo = 0.0; hi = rho * zz; tau = 0.5 * hi;
    // g(hi) >= 0 holds in exact arithmetic; widen a little against rounding
    hi = hi * (1.0 + 8.0 * kDcEps) + 1e-300;
  }
  const double dorg = d[io];
  const int isplit = last ? K - 1 : i;  // poles 0..isplit feed psi, the rest phi
  int it = 0;
  for (; it < 100; ++it) {
    double psi = 0.0, dpsi = 0.0, phi = 0.0, dphi = 0.0;
    for (int j = 0; j <= isplit; ++j) {
      const double x = (d[j] - dorg) - tau;
      const double t = z[j] / x;
      psi += rho * z[j] * t;
      dpsi += rho * t * t;
    }
    for (int j = isplit + 1; j < K; ++j) {
      const double x = (d[j] - dorg) - tau;
      const double t = z[j] / x;
      phi += rho * z[j] * t;
      dphi += rho * t * t;
    }
    const double g = 1.0 + psi + phi;
    const double erretm = 8.0 * (phi - psi) + 2.0 + fabs(tau) * (dpsi + dphi);
    if (fabs(g) <= kDcEps * erretm) break;
    if (g < 0.0) { if (tau > lo) lo = tau; } else { if (tau < hi) hi = tau; }
    // middle-way step
    double t_new;
    const double D1 = (d[isplit] - dorg) - tau;  // < 0
    if (!last) {
      const double D2 = (d[isplit + 1] - dorg) - tau;  // > 0
      const double a = dpsi * D1 * D1, A = dphi * D2 * D2;
      const double c = 1.0 + (psi - dpsi * D1) + (phi - dphi * D2);
      const double qa = c, qb = c * (D1 + D2) + a + A, qc = D1 * D2 * g;
      double eta;
      if (qa == 0.0) {
        eta = qc / qb;
      } else {
        double disc = qb * qb - 4.0 * qa * qc;
        if (disc < 0.0) disc = 0.0;
        const double q = 0.5 * (qb + copysign(sqrt(disc), qb));
        const double e2 = qc / q, e1 = q / qa;
        const double t2 = tau + e2, t1 = tau + e1;
        if (t2 > lo && t2 < hi) eta = e2;
        else if (t1 > lo && t1 < hi) eta = e1;
        else eta = NAN;
      }
      t_new = tau + eta;
    } else {
      const double a = dpsi * D1 * D1;
      const double c = 1.0 + (psi - dpsi * D1);
      t_new = (c > 0.0) ? tau + (D1 + a / c) : NAN;
    }
    if (!(t_new > lo && t_new < hi)) t_new = 0.5 * (lo + hi);
    if (t_new == tau || !(hi - lo > 0.0)) break;
    // bracket exhausted to rounding
    if (hi - lo <= 2.0 * kDcEps * fmax(fabs(lo), fabs(hi))) { tau = t_new; break; }
    tau = t_new;
  }
  if (iters_out) *iters_out = it;
  for (int j = 0; j < K; ++j) delta[(long long)j * dstride] = (d[j] - dorg) - tau;
  return dorg + tau;
}

#if defined(__CUDACC__)
// Warp-parallel variant of secular_root: the 32 lanes split the K poles of every function evaluation
// (butterfly sums, so all lanes hold identical scalars and take identical branches).  Same iteration,
// same stopping rule.  All lanes of the warp must call it; the result is valid in every lane.
__device__ inline double secular_root_warp(int K, int i, const double* __restrict__ d, const double* __restrict__ z,
                                           double rho, double* __restrict__ delta, long long dstride) {
  const int lane = threadIdx.x & 31;
  auto wsum = [](double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
  if (K == 1) {
    const double tau = rho * z[0] * z[0];
    if (lane == 0) delta[0] = -tau;
    return d[0] + tau;
  }
  const bool last = (i == K - 1);
  int io;
  double lo, hi, tau;
  if (!last) {
    const double di = d[i];
    const double del = d[i + 1] - di;
    const double mid = 0.5 * del;
    double g = 0.0;
    for (int j = lane; j < K; j += 32) g += rho * z[j] * z[j] / ((d[j] - di) - mid);
    g = 1.0 + wsum(g);
    if (g >= 0.0) { io = i; lo = 0.0; hi = mid; tau = mid; }
    else { io = i + 1; lo = -mid; hi = 0.0; tau = -mid; }
  } else {
    double zz = 0.0;
    for (int j = lane; j < K; j += 32) zz += z[j] * z[j];
    zz = wsum(zz);
    io = K - 1; lo = 0.0; hi = rho * zz; tau = 0.5 * hi;
    hi = hi * (1.0 + 8.0 * kDcEps) + 1e-300;
  }
  const double dorg = d[io];
  const int isplit = last ? K - 1 : i;
  const double dsp = d[isplit] - dorg;
  const double dsp1 = last ? 0.0 : d[isplit + 1] - dorg;
  for (int it = 0; it < 100; ++it) {
    double psi = 0.0, dpsi = 0.0, phi = 0.0, dphi = 0.0;
    for (int j = lane; j < K; j += 32) {
      const double x = (d[j] - dorg) - tau;
      const double t = z[j] / x;
      const double a = rho * z[j] * t, b = rho * t * t;
      if (j <= isplit) { psi += a; dpsi += b; } else { phi += a; dphi += b; }
    }
    psi = wsum(psi); dpsi = wsum(dpsi); phi = wsum(phi); dphi = wsum(dphi);
    const double g = 1.0 + psi + phi;
    const double erretm = 8.0 * (phi - psi) + 2.0 + fabs(tau) * (dpsi + dphi);
    if (fabs(g) <= kDcEps * erretm) break;
    if (g < 0.0) { if (tau > lo) lo = tau; } else { if (tau < hi) hi = tau; }
    double t_new;
    const double D1 = dsp - tau;
    if (!last) {
      const double D2 = dsp1 - tau;
      const double a = dpsi * D1 * D1, A = dphi * D2 * D2;
      const double c = 1.0 + (psi - dpsi * D1) + (phi - dphi * D2);
      const double qa = c, qb = c * (D1 + D2) + a + A, qc = D1 * D2 * g;
      double eta;
      if (qa == 0.0) {
        eta = qc / qb;
      } else {
        double disc = qb * qb - 4.0 * qa * qc;
        if (disc < 0.0) disc = 0.0;
        const double qq = 0.5 * (qb + copysign(sqrt(disc), qb));
        const double e2 = qc / qq, e1 = qq / qa;
        const double t2 = tau + e2, t1 = tau + e1;
        if (t2 > lo && t2 < hi) eta = e2;
        else if (t1 > lo && t1 < hi) eta = e1;
        else eta = NAN;
      }
      t_new = tau + eta;
    } else {
      const double a = dpsi * D1 * D1;
      const double c = 1.0 + (psi - dpsi * D1);
      t_new = (c > 0.0) ? tau + (D1 + a / c) : NAN;
    }
    if (!(t_new > lo && t_new < hi)) t_new = 0.5 * (lo + hi);
    if (t_new == tau || !(hi - lo > 0.0)) break;
    if (hi - lo <= 2.0 * kDcEps * fmax(fabs(lo), fabs(hi))) { tau = t_new; break; }
    tau = t_new;
  }
  for (int j = lane; j < K; j += 32) delta[(long long)j * dstride] = (d[j] - dorg) - tau;
  return dorg + tau;
}
#endif

// Deflation of one merge (the logic of LAPACK dlaed2 restated).  d, z: combined eigenvalues and
// (normalised) z vector of the two children, k entries; idx: permutation sorting d ascending;
// ctype[j] in {1 (column lives in the top rows), 3 (bottom rows)} on input, may become 2 (mixed).
// Outputs: K poles (dlamda ascending, w = z components, ndcol = their column), the deflated
// columns dcol (k-K of them, their eigenvalues stay in d[]), and the Givens rotations to apply to
// the eigenvector columns:  x = col rot_a, y = col rot_b:  x' = c x + s y,  y' = c y - s x.
GSCF_HD inline void dc_deflate(int k, double* d, double* z, double rho, const int* idx, int* ctype,
                               int* K_out, double* dlamda, double* w, int* ndcol, int* dcol,
                               int* nrot_out, int* rot_a, int* rot_b, double* rot_c, double* rot_s) {
  double dmax = 0.0, zmax = 0.0;
  for (int j = 0; j < k; ++j) {
    dmax = fmax(dmax, fabs(d[j]));
    zmax = fmax(zmax, fabs(z[j]));
  }
  const double tol = 8.0 * kDcEps * fmax(dmax, zmax);
  int K = 0, nd = 0, nrot = 0;
  if (rho * zmax <= tol) {
    for (int j = 0; j < k; ++j) dcol[nd++] = idx[j];
    *K_out = 0;
    *nrot_out = 0;
    return;
  }
  int pj = -1;
  for (int jj = 0; jj < k; ++jj) {
    const int nj = idx[jj];
    if (rho * fabs(z[nj]) <= tol) {
      dcol[nd++] = nj;
      continue;
    }
    if (pj < 0) {
      pj = nj;
      continue;
    }
    double s = z[pj], c = z[nj];
    const double tau = hypot(c, s);
    const double t = d[nj] - d[pj];
    c /= tau;
    s = -s / tau;
    if (fabs(t * c * s) <= tol) {
      z[nj] = tau;
      z[pj] = 0.0;
      rot_a[nrot] = pj; rot_b[nrot] = nj; rot_c[nrot] = c; rot_s[nrot] = s;
      ++nrot;
      if (ctype[pj] != ctype[nj]) ctype[nj] = 2;
      const double tt = d[pj] * c * c + d[nj] * s * s;
      d[nj] = d[pj] * s * s + d[nj] * c * c;
      d[pj] = tt;
      dcol[nd++] = pj;
      pj = nj;
    } else {
      dlamda[K] = d[pj]; w[K] = z[pj]; ndcol[K] = pj;
      ++K;
      pj = nj;
    }
  }
  if (pj >= 0) {
    dlamda[K] = d[pj]; w[K] = z[pj]; ndcol[K] = pj;
    ++K;
  }
  *K_out = K;
  *nrot_out = nrot;
}

}  // namespace gscfb
