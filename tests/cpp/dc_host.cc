// Host harness around gscf_b200/csrc/dc_core.hpp: a serial divide & conquer tridiagonal
// eigensolver built from the *same* deflation / secular-equation code the device kernels use.
// tests/test_dc_core_cpu.py compiles this with g++ and checks it against LAPACK on the CPU, so
// the numerically delicate part of the large-n eigensolver is validated without a GPU.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

#include "../../gscf_b200/csrc/dc_core.hpp"

using namespace gscfb;

namespace {

struct Ctx {
  int n;
  std::vector<double> d, e;  // e = original off-diagonal
  std::vector<double> Q;     // n x n column-major, block diagonal while solving
  std::vector<double> lam;
  int leaf;
  long long secular_iters = 0, secular_calls = 0;
  int max_iters = 0;
};

// cyclic Jacobi on a small dense symmetric matrix (leaf solver of the host harness)
void leaf_solve(Ctx& c, int lo, int hi) {
  const int m = hi - lo;
  std::vector<double> A(m * m, 0.0), V(m * m, 0.0);
  for (int i = 0; i < m; ++i) {
    A[i * m + i] = c.d[lo + i];
    V[i * m + i] = 1.0;
    if (i + 1 < m) A[(i + 1) * m + i] = A[i * m + i + 1] = c.e[lo + i];
  }
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0;
    for (int p = 0; p < m; ++p)
      for (int q = p + 1; q < m; ++q) off += A[p * m + q] * A[p * m + q];
    if (off < 1e-300) break;
    for (int p = 0; p < m; ++p)
      for (int q = p + 1; q < m; ++q) {
        const double apq = A[p * m + q];
        if (std::fabs(apq) < 1e-300) continue;
        const double th = (A[q * m + q] - A[p * m + p]) / (2 * apq);
        const double t = std::copysign(1.0, th) / (std::fabs(th) + std::sqrt(th * th + 1));
        const double cs = 1 / std::sqrt(t * t + 1), sn = t * cs;
        for (int k = 0; k < m; ++k) {
          const double akp = A[k * m + p], akq = A[k * m + q];
          A[k * m + p] = cs * akp - sn * akq;
          A[k * m + q] = sn * akp + cs * akq;
        }
        for (int k = 0; k < m; ++k) {
          const double apk = A[p * m + k], aqk = A[q * m + k];
          A[p * m + k] = cs * apk - sn * aqk;
          A[q * m + k] = sn * apk + cs * aqk;
        }
        for (int k = 0; k < m; ++k) {
          const double vkp = V[k + p * m], vkq = V[k + q * m];
          V[k + p * m] = cs * vkp - sn * vkq;
          V[k + q * m] = sn * vkp + cs * vkq;
        }
      }
  }
  for (int j = 0; j < m; ++j) {
    c.lam[lo + j] = A[j * m + j];
    for (int i = 0; i < m; ++i) c.Q[(size_t)(lo + j) * c.n + lo + i] = V[i + j * m];
  }
}

void merge(Ctx& c, int lo, int mid, int hi) {
  const int n = c.n, n1 = mid - lo, k = hi - lo;
  const double ecoup = c.e[mid - 1];
  double rho = std::fabs(ecoup);
  const double sgn = ecoup < 0 ? -1.0 : 1.0;
  std::vector<double> d(k), z(k);
  std::vector<int> idx(k), ctype(k);
  for (int j = 0; j < k; ++j) {
    d[j] = c.lam[lo + j];
    z[j] = j < n1 ? c.Q[(size_t)(lo + j) * n + (mid - 1)] : sgn * c.Q[(size_t)(lo + j) * n + mid];
    ctype[j] = j < n1 ? 1 : 3;
  }
  // normalise z (|z|^2 = 2), rho <- 2 rho
  for (auto& v : z) v /= std::sqrt(2.0);
  rho *= 2.0;
  std::iota(idx.begin(), idx.end(), 0);
  std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return d[a] < d[b]; });
  int K = 0, nrot = 0;
  std::vector<double> dl(k), w(k), rc(k), rs(k);
  std::vector<int> ndcol(k), dcol(k), ra(k), rb(k);
  dc_deflate(k, d.data(), z.data(), rho, idx.data(), ctype.data(), &K, dl.data(), w.data(), ndcol.data(),
             dcol.data(), &nrot, ra.data(), rb.data(), rc.data(), rs.data());
  // rotations on the eigenvector columns
  for (int r = 0; r < nrot; ++r) {
    double* x = &c.Q[(size_t)(lo + ra[r]) * n + lo];
    double* y = &c.Q[(size_t)(lo + rb[r]) * n + lo];
    for (int i = 0; i < k; ++i) {
      const double xv = x[i], yv = y[i];
      x[i] = rc[r] * xv + rs[r] * yv;
      y[i] = rc[r] * yv - rs[r] * xv;
    }
  }
  std::vector<double> S((size_t)K * K), lamn(K);
  for (int i = 0; i < K; ++i) {
    int its = 0;
    lamn[i] = secular_root(K, i, dl.data(), w.data(), rho, &S[(size_t)i * K], 1, &its);
    c.secular_iters += its;
    c.secular_calls++;
    c.max_iters = std::max(c.max_iters, its);
  }
  // Gu-Eisenstat: zhat_j^2 = prod_i delta(j,i) / prod_{i != j} (d_j - d_i)   (up to rho)
  std::vector<double> zh(K);
  for (int j = 0; j < K; ++j) {
    double p = S[(size_t)j * K + j];
    for (int i = 0; i < K; ++i)
      if (i != j) p *= S[(size_t)i * K + j] / (dl[j] - dl[i]);
    zh[j] = std::copysign(std::sqrt(-p), w[j]);
  }
  std::vector<double> U((size_t)K * K);
  for (int i = 0; i < K; ++i) {
    double nrm = 0;
    for (int j = 0; j < K; ++j) {
      U[(size_t)i * K + j] = zh[j] / S[(size_t)i * K + j];
      nrm += U[(size_t)i * K + j] * U[(size_t)i * K + j];
    }
    nrm = std::sqrt(nrm);
    for (int j = 0; j < K; ++j) U[(size_t)i * K + j] /= nrm;
  }
  // new vectors = Q[:, ndcol] * U ; deflated copied
  std::vector<double> Qn((size_t)k * k, 0.0), ln(k);
  for (int i = 0; i < K; ++i) {
    ln[i] = lamn[i];
    for (int j = 0; j < K; ++j) {
      const double u = U[(size_t)i * K + j];
      const double* q = &c.Q[(size_t)(lo + ndcol[j]) * n + lo];
      for (int r = 0; r < k; ++r) Qn[(size_t)i * k + r] += q[r] * u;
    }
  }
  for (int t = 0; t < k - K; ++t) {
    ln[K + t] = d[dcol[t]];
    const double* q = &c.Q[(size_t)(lo + dcol[t]) * n + lo];
    for (int r = 0; r < k; ++r) Qn[(size_t)(K + t) * k + r] = q[r];
  }
  for (int j = 0; j < k; ++j) {
    c.lam[lo + j] = ln[j];
    for (int r = 0; r < k; ++r) c.Q[(size_t)(lo + j) * n + lo + r] = Qn[(size_t)j * k + r];
  }
}

void solve(Ctx& c, int lo, int hi) {
  if (hi - lo <= c.leaf) {
    leaf_solve(c, lo, hi);
    return;
  }
  const int mid = lo + (hi - lo) / 2;
  const double rho = std::fabs(c.e[mid - 1]);
  c.d[mid - 1] -= rho;
  c.d[mid] -= rho;
  solve(c, lo, mid);
  solve(c, mid, hi);
  merge(c, lo, mid, hi);
}

}  // namespace

extern "C" int dc_host_solve(int n, const double* d, const double* e, int leaf, double* w, double* Z,
                             double* stats /* [avg secular iterations, max] */) {
  Ctx c;
  c.n = n;
  c.d.assign(d, d + n);
  c.e.assign(e, e + std::max(0, n - 1));
  c.e.push_back(0.0);
  c.Q.assign((size_t)n * n, 0.0);
  c.lam.assign(n, 0.0);
  c.leaf = leaf;
  // the leaves must see the coupling-free blocks: zero e at the split points is implicit because
  // leaf_solve only reads e inside [lo, hi)
  solve(c, 0, n);
  std::vector<int> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return c.lam[a] < c.lam[b]; });
  for (int j = 0; j < n; ++j) {
    w[j] = c.lam[order[j]];
    for (int i = 0; i < n; ++i) Z[(size_t)j * n + i] = c.Q[(size_t)order[j] * n + i];
  }
  if (stats) {
    stats[0] = c.secular_calls ? (double)c.secular_iters / c.secular_calls : 0.0;
    stats[1] = c.max_iters;
  }
  return 0;
}
