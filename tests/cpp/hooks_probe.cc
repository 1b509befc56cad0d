// CPU-only probe of the host layer's compile-time / run-time plumbing (no engine is created):
//  * ScfBase::b200_hooks_overridden(): hook trampolines are registered only for solver types that override a hook,
//  * OperatorLinearCombination for ordinary and block-diagonal (alpha / beta) operators (OperatorLinearCombination.hh:26-84),
//  * PulayDiisScf::diis_linear_system_matrix against the layout of PulayDiisScf.hh:326-356,
//  * the default residual error (ScfBase.hh:268-289) before and after a second eigensolution exists.
// Prints "PROBE OK" and exits 0 when every check holds.
#include <cmath>
#include <cstdio>

#include <gscf/PlainScf.hh>
#include <gscf/PulayDiisScf.hh>
#include <gscf/TruncatedOptDampScf.hh>

typedef lazyten::SmallMatrix<double> matrix_type;
typedef matrix_type::vector_type vector_type;

// a stored n x n block usable as a lazy matrix
struct DenseBlock : public lazyten::LazyMatrix_i<matrix_type> {
  matrix_type m;
  explicit DenseBlock(size_t n = 0) : m(n, n) {}
  size_t n_rows() const override { return m.n_rows(); }
  size_t n_cols() const override { return m.n_cols(); }
  double operator()(size_t r, size_t c) const override { return m(r, c); }
  void apply(const lazyten::MultiVector<const lazyten::MutableMemoryVector_i<double>>& x,
             lazyten::MultiVector<lazyten::MutableMemoryVector_i<double>>& y,
             const lazyten::Transposed mode = lazyten::Transposed::None, const double c_this = 1,
             const double c_y = 0) const override {
    m.apply(x, y, mode, c_this, c_y);
  }
  void update(const krims::GenMap&) override {}
  lazy_matrix_expression_ptr_type clone() const override { return lazy_matrix_expression_ptr_type(new DenseBlock(*this)); }
};

struct RestrictedFock final : public DenseBlock, public gscf::FocklikeMatrix_i<double> {
  typedef double scalar_type;
  size_t nocc;
  RestrictedFock(size_t n, size_t o) : DenseBlock(n), nocc(o) {}
  const std::string& scf_update_key() const override { static const std::string k = "evec_coefficients"; return k; }
  double energy_1e_terms() const override { return 0; }
  double energy_2e_terms() const override { return 0; }
  krims::Range<size_t> indices_orbspace(gscf::OrbitalSpace) const override { return {0, nocc}; }
};

struct UnrestrictedFock final : public lazyten::BlockDiagonalMatrix<DenseBlock, 2>, public gscf::FocklikeMatrix_i<double> {
  typedef lazyten::BlockDiagonalMatrix<DenseBlock, 2> bd_type;
  typedef double scalar_type;
  size_t na, nb;
  UnrestrictedFock(size_t n, size_t na_, size_t nb_) : bd_type({DenseBlock(n), DenseBlock(n)}), na(na_), nb(nb_) {}
  const std::string& scf_update_key() const override { static const std::string k = "evec_coefficients"; return k; }
  void update(const krims::GenMap&) {}
  double energy_1e_terms() const override { return 0; }
  double energy_2e_terms() const override { return 0; }
  krims::Range<size_t> indices_orbspace(gscf::OrbitalSpace osp) const override {
    const size_t n = diag_blocks()[0].n_rows();
    return osp == gscf::OrbitalSpace::OCC_ALPHA ? krims::Range<size_t>{0, na} : krims::Range<size_t>{n, n + nb};
  }
};

typedef gscf::PulayDiisScf<gscf::PulayDiisScfState<RestrictedFock, matrix_type>> Diis;
typedef gscf::PlainScf<gscf::PlainScfState<RestrictedFock, matrix_type>> Plain;
typedef gscf::TruncatedOptDampScf<gscf::TruncatedOptDampScfState<RestrictedFock, matrix_type>> Toda;

template <typename S>
struct Probe : public S {
  bool hooks() const { return this->b200_hooks_overridden(); }
  matrix_type residual(const typename S::state_type& s) const { return this->calculate_error(s); }
};
struct ErrOnly : public Probe<Diis> {
  matrix_type calculate_error(const state_type&) const override { return matrix_type(2, 2); }
};
struct WithAfter : public Probe<Plain> {
  void after_iteration_step(state_type&) const override {}
};
struct WithDiag : public Probe<Diis> {
  void on_new_diagmat(state_type&) const override {}
};
struct WithEig : public Probe<Toda> {
  void on_update_eigenpairs(state_type&) const override {}
};
struct WithTodaDiag : public Probe<Toda> {
  void on_new_diagmat(state_type&) const override {}
};

static int g_fail = 0;
#define CHECK(c) do { if (!(c)) { std::printf("CHECK failed line %d: %s\n", __LINE__, #c); ++g_fail; } } while (0)

int main() {
  // ---- hook detection ----
#if defined(__GNUC__) && !defined(__clang__)
  CHECK(!Probe<Plain>().hooks());
  CHECK(!Probe<Diis>().hooks());
  CHECK(!Probe<Toda>().hooks());
  CHECK(!ErrOnly().hooks());  // calculate_error is not a hook: no per-iteration mirror needed
  CHECK(WithAfter().hooks());
  CHECK(WithDiag().hooks());
  CHECK(WithEig().hooks());
  CHECK(WithTodaDiag().hooks());
#endif
  // ---- linear combinations ----
  static_assert(!lazyten::IsBlockDiagonalMatrix<RestrictedFock>::value, "");
  static_assert(lazyten::IsBlockDiagonalMatrix<UnrestrictedFock>::value, "");
  RestrictedFock f1(3, 1), f2(3, 1);
  UnrestrictedFock u1(3, 2, 1), u2(3, 2, 1);
  for (size_t i = 0; i < 3; ++i)
    for (size_t j = 0; j < 3; ++j) {
      f1.m(i, j) = 1.0 + i + j; f2.m(i, j) = 10.0 * (i + 1) * (j + 1);
      u1.diag_blocks()[0].m(i, j) = 1.0 + i * j; u1.diag_blocks()[1].m(i, j) = 2.0 + i + j;
      u2.diag_blocks()[0].m(i, j) = -1.0 * (i + j); u2.diag_blocks()[1].m(i, j) = 0.5 * (i + 1) * (j + 1);
    }
  gscf::OperatorLinearCombination<RestrictedFock> lc(f1, 0.25);
  lc.push_term(f2, 0.75);
  CHECK(lc.n_terms() == 2 && lc.n_rows() == 3);
  CHECK(std::fabs(lc(1, 2) - (0.25 * f1.m(1, 2) + 0.75 * f2.m(1, 2))) < 1e-15);
  gscf::OperatorLinearCombination<UnrestrictedFock> ulc(u1, 0.4);
  ulc.push_term(u2, 0.6);
  CHECK(ulc.n_terms() == 2 && ulc.n_rows() == 6 && ulc.diag_blocks().size() == 2);
  CHECK(std::fabs(ulc(1, 2) - (0.4 * u1.diag_blocks()[0].m(1, 2) + 0.6 * u2.diag_blocks()[0].m(1, 2))) < 1e-15);
  CHECK(std::fabs(ulc(4, 5) - (0.4 * u1.diag_blocks()[1].m(1, 2) + 0.6 * u2.diag_blocks()[1].m(1, 2))) < 1e-15);
  CHECK(ulc(1, 4) == 0.0 && ulc(5, 0) == 0.0);  // off-diagonal blocks
  CHECK(gscf::b200::Blocks<UnrestrictedFock>::count == 2 && gscf::b200::Blocks<UnrestrictedFock>::order(u1) == 3);
  CHECK(gscf::b200::Blocks<RestrictedFock>::count == 1 && gscf::b200::Blocks<RestrictedFock>::order(f1) == 3);
  // ---- DIIS system layout ----
  matrix_type S(3, 3);
  for (size_t i = 0; i < 3; ++i) S(i, i) = 1;
  Diis::state_type st(f1, S);
  st.resize_buffers(4);
  st.error_overlaps.push_back({1.0});            // <0|0>
  st.error_overlaps.push_back({4.0, 2.0});       // <1|1>, <1|0>
  st.error_overlaps.push_back({9.0, 6.0, 3.0});  // <2|2>, <2|1>, <2|0>
  const matrix_type B = Diis().diis_linear_system_matrix(st);
  CHECK(B.n_rows() == 4 && B.n_cols() == 4);
  CHECK(B(0, 0) == 1.0 && B(1, 1) == 4.0 && B(2, 2) == 9.0 && B(3, 3) == 0.0);
  CHECK(B(1, 0) == 2.0 && B(0, 1) == 2.0 && B(2, 1) == 6.0 && B(1, 2) == 6.0 && B(2, 0) == 3.0 && B(0, 2) == 3.0);
  for (size_t i = 0; i < 3; ++i) CHECK(B(i, 3) == -1.0 && B(3, i) == -1.0);
  // ---- default residual error ----
  Probe<Plain> pp;
  Plain::state_type ps(f1, S);
  CHECK(std::isnan(ps.last_error_norm));
  matrix_type e0 = pp.residual(ps);
  CHECK(e0.n_rows() == 1 && e0.n_cols() == 1 && std::isnan(e0(0, 0)));
  Plain::state_type::esoln_type a, b;
  a.evectors_ptr = std::make_shared<lazyten::MultiVector<vector_type>>(3, 2);
  b.evectors_ptr = std::make_shared<lazyten::MultiVector<vector_type>>(3, 2);
  (*a.evectors_ptr)[0](1) = 1.0; (*b.evectors_ptr)[0](1) = 0.75; (*b.evectors_ptr)[1](2) = -2.0;
  ps.push_new_eigensolution(a, {0, 0});
  ps.push_new_eigensolution(b, {0, 0});
  matrix_type e1 = pp.residual(ps);
  CHECK(e1.n_rows() == 3 && e1.n_cols() == 2 && e1(1, 0) == -0.25 && e1(2, 1) == -2.0 && e1(0, 0) == 0.0);
  // ---- the "eigensolver" sub-map reaches the inner eigensolver choice (ScfBase.hh:104) ----
  {
    krims::GenMap sub{{"method", std::string("tridiag_dc")}};
    Diis dsub({{"eigensolver", sub}, {"diis_n_prev_steps", size_t(6)}});
    CHECK(dsub.b200_eig_method == GSCF_B200_EIG_TRIDIAG_DC && dsub.n_prev_steps == 6);
    krims::GenMap sub2{{"method", std::string("lapack")}, {"tolerance", 1e-9}};
    Plain psub({{"eigensolver", sub2}});
    CHECK(psub.b200_eig_method == GSCF_B200_EIG_AUTO);
    krims::GenMap back;
    dsub.get_control_params(back);
    CHECK(back.submap("eigensolver").at("method", std::string("")) == "tridiag_dc");
  }
  std::printf(g_fail == 0 ? "PROBE OK\n" : "%d PROBE CHECKS FAILED\n", g_fail);
  return g_fail == 0 ? 0 : 1;
}
