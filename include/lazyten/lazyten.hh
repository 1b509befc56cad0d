// Minimal lazyten surface used by gscf - part of the gscf-b200 C++ host
// layer.  Stored matrices / vectors are plain host containers; the SCF arithmetic itself runs on
// the GPU behind include/gscf_b200.h.  The free functions below (dot, outer_prod_sum, ...) exist
// so that user plugins written against lazyten (e.g. gscfmock::FockMatrix, pulay_error) compile;
// they are host reference loops, not the product path.
#pragma once
#include <array>
#include <cmath>
#include <cstdio>
#include <initializer_list>
#include <limits>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "../krims/krims.hh"

namespace lazyten {

template <typename T>
struct Constants {
  static constexpr T all = std::numeric_limits<T>::max();
  static constexpr T invalid = std::numeric_limits<T>::has_quiet_NaN ? std::numeric_limits<T>::quiet_NaN()
                                                                     : std::numeric_limits<T>::max();
  static constexpr T zero = T(0);
  static constexpr T one = T(1);
  static constexpr T default_tolerance = T(1e-14);
};
template <typename T>
constexpr T Constants<T>::all;
template <typename T>
constexpr T Constants<T>::invalid;
template <typename T>
constexpr T Constants<T>::zero;
template <typename T>
constexpr T Constants<T>::one;
template <typename T>
constexpr T Constants<T>::default_tolerance;

enum class Transposed { None, Trans, ConjTrans };
enum class OperatorProperties { None, RealSymmetric, Hermitian };

// ---- vectors ----------------------------------------------------------------------------------------
template <typename Scalar>
class SmallVector {
 public:
  typedef Scalar scalar_type;
  typedef Scalar real_type;
  typedef size_t size_type;
  typedef typename std::vector<Scalar>::iterator iterator;
  typedef typename std::vector<Scalar>::const_iterator const_iterator;
  SmallVector() {}
  explicit SmallVector(size_t n, bool zero = true) : m_d(n, Scalar(0)) { (void)zero; }
  SmallVector(std::initializer_list<Scalar> il) : m_d(il) {}
  template <typename It, typename = decltype(*std::declval<It>())>
  SmallVector(It b, It e) : m_d(b, e) {}
  size_t size() const { return m_d.size(); }
  size_t n_elem() const { return m_d.size(); }
  Scalar& operator()(size_t i) { return m_d[i]; }
  const Scalar& operator()(size_t i) const { return m_d[i]; }
  Scalar& operator[](size_t i) { return m_d[i]; }
  const Scalar& operator[](size_t i) const { return m_d[i]; }
  iterator begin() { return m_d.begin(); }
  iterator end() { return m_d.end(); }
  const_iterator begin() const { return m_d.begin(); }
  const_iterator end() const { return m_d.end(); }
  Scalar* memptr() { return m_d.data(); }
  const Scalar* memptr() const { return m_d.data(); }
  SmallVector& operator*=(Scalar s) {
    for (auto& v : m_d) v *= s;
    return *this;
  }

 private:
  std::vector<Scalar> m_d;
};
template <typename Scalar>
using MutableMemoryVector_i = SmallVector<Scalar>;
template <typename... Ts>
using mat_vec_apply_enabled_t = int;

template <typename Vector>
class MultiVector {
 public:
  typedef typename std::remove_const<Vector>::type vector_type;
  typedef typename vector_type::scalar_type scalar_type;
  MultiVector() {}
  MultiVector(size_t n_elem, size_t n_vectors, bool zero = true) {
    (void)zero;
    for (size_t j = 0; j < n_vectors; ++j) m_v.push_back(std::make_shared<vector_type>(n_elem));
  }
  // n_elem x n_vectors table, row i lists component i of every vector (lazyten initialiser order)
  MultiVector(std::initializer_list<std::initializer_list<scalar_type>> rows) {
    const size_t ne = rows.size(), nv = ne ? rows.begin()->size() : 0;
    for (size_t j = 0; j < nv; ++j) m_v.push_back(std::make_shared<vector_type>(ne));
    size_t i = 0;
    for (const auto& r : rows) {
      size_t j = 0;
      for (const auto& x : r) (*std::const_pointer_cast<vector_type>(m_v[j++]))(i) = x;
      ++i;
    }
  }
  template <typename V2, typename = krims::enable_if_t<std::is_same<typename std::remove_const<V2>::type,
                                                                    vector_type>::value>>
  MultiVector(const MultiVector<V2>& o) {
    for (size_t j = 0; j < o.n_vectors(); ++j) m_v.push_back(o.ptr(j));
  }
  size_t n_vectors() const { return m_v.size(); }
  size_t n_elem() const { return m_v.empty() ? 0 : m_v[0]->size(); }
  Vector& operator[](size_t j) { return *std::const_pointer_cast<vector_type>(m_v[j]); }
  const Vector& operator[](size_t j) const { return *m_v[j]; }
  std::shared_ptr<const vector_type> ptr(size_t j) const { return m_v[j]; }
  void push_back(std::shared_ptr<const vector_type> v) { m_v.push_back(std::move(v)); }
  MultiVector<const vector_type> subview(const krims::Range<size_t>& r) const {
    MultiVector<const vector_type> out;
    for (size_t j = r.lower_bound(); j < r.upper_bound(); ++j) out.push_back(m_v[j]);
    return out;
  }

 private:
  std::vector<std::shared_ptr<const vector_type>> m_v;
};

template <typename Vector>
MultiVector<Vector> as_multivector(const Vector& v) {
  MultiVector<Vector> mv;
  mv.push_back(std::make_shared<Vector>(v));
  return mv;
}
template <typename Vector, typename Container>
MultiVector<Vector> make_as_multivector(const Container& c) {
  MultiVector<Vector> mv;
  mv.push_back(std::make_shared<Vector>(std::begin(c), std::end(c)));
  return mv;
}
template <typename V1, typename V2>
void adjust_phase(V1& v, const V2& ref) {
  double d = 0;
  for (size_t i = 0; i < ref.size(); ++i) d += v(i) * ref(i);
  if (d < 0)
    for (size_t i = 0; i < ref.size(); ++i) const_cast<typename std::remove_const<V1>::type&>(v)(i) = -v(i);
}

// ---- stored matrix -----------------------------------------------------------------------------------
template <typename Scalar>
class SmallMatrix {
 public:
  typedef Scalar scalar_type;
  typedef Scalar real_type;
  typedef size_t size_type;
  typedef SmallVector<Scalar> vector_type;
  typedef SmallMatrix<Scalar> stored_matrix_type;
  SmallMatrix() : m_r(0), m_c(0) {}
  SmallMatrix(size_t r, size_t c, bool zero = true) : m_r(r), m_c(c), m_d(r * c, Scalar(0)) { (void)zero; }
  SmallMatrix(std::initializer_list<std::initializer_list<Scalar>> rows)
        : m_r(rows.size()), m_c(rows.size() ? rows.begin()->size() : 0), m_d(m_r * m_c) {
    size_t i = 0;
    for (const auto& r : rows) {
      size_t j = 0;
      for (const auto& x : r) (*this)(i, j++) = x;
      ++i;
    }
  }
  size_t n_rows() const { return m_r; }
  size_t n_cols() const { return m_c; }
  Scalar& operator()(size_t i, size_t j) { return m_d[j * m_r + i]; }
  const Scalar& operator()(size_t i, size_t j) const { return m_d[j * m_r + i]; }
  Scalar* memptr() { return m_d.data(); }  // column-major, ld = n_rows
  const Scalar* memptr() const { return m_d.data(); }
  void set_zero() { std::fill(m_d.begin(), m_d.end(), Scalar(0)); }
  void add_properties(OperatorProperties) {}
  bool is_hermitian(Scalar tol = Constants<Scalar>::default_tolerance) const {
    if (m_r != m_c) return false;
    for (size_t i = 0; i < m_r; ++i)
      for (size_t j = 0; j < i; ++j)
        if (std::fabs((*this)(i, j) - (*this)(j, i)) > tol * std::max(Scalar(1), std::fabs((*this)(i, j)))) return false;
    return true;
  }
  template <typename VI, typename VO>
  void apply(const MultiVector<VI>& x, MultiVector<VO>& y, Transposed mode = Transposed::None, Scalar c_this = 1,
             Scalar c_y = 0) const {
    const bool tr = mode != Transposed::None;
    for (size_t v = 0; v < x.n_vectors(); ++v)
      for (size_t i = 0; i < (tr ? m_c : m_r); ++i) {
        Scalar s = 0;
        for (size_t k = 0; k < (tr ? m_r : m_c); ++k) s += (tr ? (*this)(k, i) : (*this)(i, k)) * x[v](k);
        y[v](i) = c_this * s + (c_y == Scalar(0) ? Scalar(0) : c_y * y[v](i));
      }
  }
  SmallMatrix& operator+=(const SmallMatrix& o) {
    for (size_t k = 0; k < m_d.size(); ++k) m_d[k] += o.m_d[k];
    return *this;
  }
  SmallMatrix& operator-=(const SmallMatrix& o) {
    for (size_t k = 0; k < m_d.size(); ++k) m_d[k] -= o.m_d[k];
    return *this;
  }
  SmallMatrix& operator*=(Scalar s) {
    for (auto& v : m_d) v *= s;
    return *this;
  }

 private:
  size_t m_r, m_c;
  std::vector<Scalar> m_d;
};
template <typename S>
SmallMatrix<S> operator+(SmallMatrix<S> a, const SmallMatrix<S>& b) { return a += b; }
template <typename S>
SmallMatrix<S> operator-(SmallMatrix<S> a, const SmallMatrix<S>& b) { return a -= b; }
template <typename S>
SmallMatrix<S> operator*(S s, SmallMatrix<S> a) { return a *= s; }
template <typename S>
SmallMatrix<S> operator*(SmallMatrix<S> a, S s) { return a *= s; }
template <typename S>
SmallMatrix<S> operator*(const SmallMatrix<S>& a, const SmallMatrix<S>& b) {
  SmallMatrix<S> c(a.n_rows(), b.n_cols());
  for (size_t j = 0; j < b.n_cols(); ++j)
    for (size_t k = 0; k < a.n_cols(); ++k)
      for (size_t i = 0; i < a.n_rows(); ++i) c(i, j) += a(i, k) * b(k, j);
  return c;
}

template <typename T>
struct StoredTypeOf {
  typedef typename T::stored_matrix_type type;
};
// BlockDiagonalMatrix<Block, N> / IsBlockDiagonalMatrix / LazyMatrixSum: the surface OperatorLinearCombination.hh:26-84
// and ScfBase.hh:300-312 touch.  An unrestricted (alpha / beta) problem matrix derives from BlockDiagonalMatrix<Block, 2>.
template <typename Block, size_t N>
class BlockDiagonalMatrix {
 public:
  typedef Block block_type;
  typedef typename Block::stored_matrix_type stored_matrix_type;
  typedef typename stored_matrix_type::scalar_type scalar_type;
  typedef typename stored_matrix_type::real_type real_type;
  typedef typename stored_matrix_type::size_type size_type;
  typedef typename stored_matrix_type::vector_type vector_type;
  static constexpr size_t n_blocks = N;
  explicit BlockDiagonalMatrix(std::array<Block, N> blocks) : m_blocks(std::move(blocks)) {}
  virtual ~BlockDiagonalMatrix() = default;
  BlockDiagonalMatrix(const BlockDiagonalMatrix&) = default;
  BlockDiagonalMatrix(BlockDiagonalMatrix&&) = default;
  BlockDiagonalMatrix& operator=(const BlockDiagonalMatrix&) = default;
  BlockDiagonalMatrix& operator=(BlockDiagonalMatrix&&) = default;
  const std::array<Block, N>& diag_blocks() const { return m_blocks; }
  std::array<Block, N>& diag_blocks() { return m_blocks; }
  size_t n_rows() const {
    size_t r = 0;
    for (const auto& b : m_blocks) r += b.n_rows();
    return r;
  }
  size_t n_cols() const { return n_rows(); }
  scalar_type operator()(size_t row, size_t col) const {
    size_t off = 0;
    for (const auto& b : m_blocks) {
      const size_t nb = b.n_rows();
      if (row < off + nb) return (col >= off && col < off + nb) ? b(row - off, col - off) : scalar_type(0);
      off += nb;
    }
    return scalar_type(0);
  }
  bool is_hermitian(real_type tol = Constants<real_type>::default_tolerance) const {
    for (const auto& b : m_blocks)
      if (!b.is_hermitian(tol)) return false;
    return true;
  }

 private:
  std::array<Block, N> m_blocks;
};
template <typename T, typename = void>
struct IsBlockDiagonalMatrix : public std::false_type {};
template <typename T>
struct IsBlockDiagonalMatrix<T, krims::VoidType<typename T::block_type, decltype(T::n_blocks)>>
      : public std::is_base_of<BlockDiagonalMatrix<typename T::block_type, T::n_blocks>, T> {};

// Lazy matrix interface (what a problem matrix plugin derives from, FockMatrix.hh:91)
template <typename StoredMatrix>
class LazyMatrix_i {
 public:
  typedef StoredMatrix stored_matrix_type;
  typedef typename StoredMatrix::scalar_type scalar_type;
  typedef typename StoredMatrix::real_type real_type;
  typedef typename StoredMatrix::size_type size_type;
  typedef typename StoredMatrix::vector_type vector_type;
  typedef std::unique_ptr<LazyMatrix_i<StoredMatrix>> lazy_matrix_expression_ptr_type;
  virtual ~LazyMatrix_i() = default;
  virtual size_t n_rows() const = 0;
  virtual size_t n_cols() const = 0;
  virtual scalar_type operator()(size_t row, size_t col) const = 0;
  virtual void apply(const MultiVector<const MutableMemoryVector_i<scalar_type>>& x,
                     MultiVector<MutableMemoryVector_i<scalar_type>>& y, const Transposed mode = Transposed::None,
                     const scalar_type c_this = 1, const scalar_type c_y = 0) const = 0;
  virtual void update(const krims::GenMap&) = 0;
  virtual lazy_matrix_expression_ptr_type clone() const = 0;
  bool is_hermitian(real_type tol = Constants<real_type>::default_tolerance) const {
    for (size_t i = 0; i < n_rows(); ++i)
      for (size_t j = 0; j < i; ++j) {
        const scalar_type a = (*this)(i, j), b = (*this)(j, i);
        if (std::fabs(a - b) > tol * std::max(real_type(1), std::fabs(a))) return false;
      }
    return true;
  }
};

// sum_i c_i M_i of lazy matrices, evaluated element-wise on demand (the terms must outlive the sum)
template <typename StoredMatrix>
class LazyMatrixSum {
 public:
  typedef StoredMatrix stored_matrix_type;
  typedef typename StoredMatrix::scalar_type scalar_type;
  typedef typename StoredMatrix::real_type real_type;
  typedef typename StoredMatrix::size_type size_type;
  typedef typename StoredMatrix::vector_type vector_type;
  typedef LazyMatrix_i<StoredMatrix> term_type;
  LazyMatrixSum() {}
  LazyMatrixSum(const term_type& op, scalar_type coeff) { add_term(op, coeff); }
  void add_term(const term_type& op, scalar_type coeff) { m_terms.emplace_back(&op, coeff); }
  size_t n_terms() const { return m_terms.size(); }
  size_t n_rows() const { return m_terms.empty() ? 0 : m_terms.front().first->n_rows(); }
  size_t n_cols() const { return m_terms.empty() ? 0 : m_terms.front().first->n_cols(); }
  scalar_type operator()(size_t i, size_t j) const {
    scalar_type s = 0;
    for (const auto& t : m_terms) s += t.second * (*t.first)(i, j);
    return s;
  }
  bool is_hermitian(real_type tol = Constants<real_type>::default_tolerance) const {
    for (const auto& t : m_terms)
      if (!t.first->is_hermitian(tol)) return false;
    return true;
  }

 private:
  std::vector<std::pair<const term_type*, scalar_type>> m_terms;
};

// matrix x multivector
template <typename S, typename V>
MultiVector<SmallVector<S>> operator*(const SmallMatrix<S>& m, const MultiVector<V>& x) {
  MultiVector<SmallVector<S>> y(m.n_rows(), x.n_vectors());
  m.apply(x, y);
  return y;
}
template <typename Stored, typename V>
MultiVector<typename Stored::vector_type> operator*(const LazyMatrix_i<Stored>& m, const MultiVector<V>& x) {
  MultiVector<typename Stored::vector_type> y(m.n_rows(), x.n_vectors());
  MultiVector<const typename Stored::vector_type> xc(x);
  m.apply(xc, y);
  return y;
}
template <typename V>
MultiVector<typename MultiVector<V>::vector_type> operator*(typename MultiVector<V>::scalar_type s,
                                                            const MultiVector<V>& x) {
  MultiVector<typename MultiVector<V>::vector_type> y(x.n_elem(), x.n_vectors());
  for (size_t v = 0; v < x.n_vectors(); ++v)
    for (size_t i = 0; i < x.n_elem(); ++i) y[v](i) = s * x[v](i);
  return y;
}

// sum_k a_k b_k^T
template <typename VA, typename VB>
SmallMatrix<typename MultiVector<VA>::scalar_type> outer_prod_sum(const MultiVector<VA>& a, const MultiVector<VB>& b) {
  SmallMatrix<typename MultiVector<VA>::scalar_type> r(a.n_elem(), b.n_elem());
  for (size_t k = 0; k < a.n_vectors(); ++k)
    for (size_t j = 0; j < b.n_elem(); ++j)
      for (size_t i = 0; i < a.n_elem(); ++i) r(i, j) += a[k](i) * b[k](j);
  return r;
}
template <typename S>
S trace(const SmallMatrix<S>& m) {
  S t = 0;
  for (size_t i = 0; i < m.n_rows(); ++i) t += m(i, i);
  return t;
}
template <typename S>
S dot(const SmallMatrix<S>& a, const SmallMatrix<S>& b) {
  S t = 0;
  for (size_t j = 0; j < a.n_cols(); ++j)
    for (size_t i = 0; i < a.n_rows(); ++i) t += a(i, j) * b(i, j);
  return t;
}
template <typename S>
S norm_frobenius(const SmallMatrix<S>& a) {
  return std::sqrt(dot(a, a));
}

// ---- eigensolutions ---------------------------------------------------------------------------------
template <typename Scalar, typename Vector>
struct Eigensolution {
  typedef Scalar evalue_type;
  typedef Vector evector_type;
  std::shared_ptr<std::vector<Scalar>> evalues_ptr;
  std::shared_ptr<MultiVector<Vector>> evectors_ptr;
  Eigensolution()
        : evalues_ptr(std::make_shared<std::vector<Scalar>>()), evectors_ptr(std::make_shared<MultiVector<Vector>>()) {}
  std::vector<Scalar>& evalues() { return *evalues_ptr; }
  const std::vector<Scalar>& evalues() const { return *evalues_ptr; }
  MultiVector<Vector>& evectors() { return *evectors_ptr; }
  const MultiVector<Vector>& evectors() const { return *evectors_ptr; }
};
template <bool Hermitian, typename Matrix>
using EigensolutionTypeFor =
      Eigensolution<typename Matrix::scalar_type, typename Matrix::stored_matrix_type::vector_type>;

struct EigensystemSolverKeys {
  static const std::string& tolerance_key() {
    static const std::string k = "tolerance";
    return k;
  }
};

// ---- solver bases -------------------------------------------------------------------------------------
class SolverException : public krims::ExceptionBase {
 public:
  explicit SolverException(const std::string& m) : krims::ExceptionBase(m), m_extra(m) {}
  const std::string& extra() const { return m_extra; }

 private:
  std::string m_extra;
};
#define DefSolverException1(name, type1, arg1, msg)                                   \
  class name : public ::lazyten::SolverException {                                   \
   public:                                                                            \
    explicit name(const type1& arg1) : ::lazyten::SolverException(make(arg1)) {}     \
                                                                                      \
   private:                                                                           \
    static std::string make(const type1& arg1) {                                     \
      std::stringstream ss;                                                           \
      ss msg;                                                                         \
      return ss.str();                                                                \
    }                                                                                 \
  }
DefSolverException1(ExcMaximumNumberOfIterationsReached, size_t, max_iter,
                    << "The iterative solver reached the maximum number of iterations (" << max_iter << ")");
#define solver_assert(cond, state, exc) \
  do {                                  \
    if (!(cond)) {                      \
      auto _e = exc;                    \
      (state).fail(_e.what());          \
      throw _e;                         \
    }                                   \
  } while (0)

class SolverStateBase {
 public:
  virtual ~SolverStateBase() = default;
  bool is_failed() const { return m_failed; }
  const std::string& fail_reason() const { return m_reason; }
  void fail(const std::string& reason) {
    m_failed = true;
    m_reason = reason;
  }
  void clear_failed() {
    m_failed = false;
    m_reason.clear();
  }

 private:
  bool m_failed = false;
  std::string m_reason;
};
template <typename Base>
class IterativeStateWrapper : public Base {
 public:
  IterativeStateWrapper() : Base(), m_n_iter(0) {}
  explicit IterativeStateWrapper(Base b) : Base(std::move(b)), m_n_iter(0) {}
  size_t n_iter() const { return m_n_iter; }
  void increase_iteration_count() { ++m_n_iter; }
  void set_iteration_count(size_t n) { m_n_iter = n; }
  virtual size_t n_mtx_applies() const = 0;

 private:
  size_t m_n_iter;
};

struct IterativeWrapperKeys {
  static const std::string max_iter;
};
#ifndef GSCF_B200_NO_KEY_DEFINITIONS
inline const std::string& iterative_wrapper_max_iter_key() {
  static const std::string k = "max_iter";
  return k;
}
#endif

template <typename State>
class SolverBase {
 public:
  typedef State state_type;
  virtual ~SolverBase() = default;
  virtual void solve_state(state_type& state) const = 0;
  void update_control_params(const krims::GenMap&) {}
  void get_control_params(krims::GenMap&) const {}
};
template <typename Base>
class IterativeWrapper : public Base {
 public:
  typedef typename Base::state_type state_type;
  size_t max_iter = 100;
  void update_control_params(const krims::GenMap& map) {
    Base::update_control_params(map);
    max_iter = map.at("max_iter", max_iter);
  }
  void get_control_params(krims::GenMap& map) const {
    Base::get_control_params(map);
    map.update("max_iter", max_iter);
  }
  virtual bool is_converged(const state_type&) const { return false; }

 protected:
  virtual void before_iteration_step(state_type&) const {}
  virtual void after_iteration_step(state_type&) const {}
  // is_converged first, then the max_iter assert (first iteration is 1)
  bool convergence_reached(state_type& s) const {
    if (is_converged(s)) return true;
    solver_assert(s.n_iter() < max_iter, s, ExcMaximumNumberOfIterationsReached(max_iter));
    return false;
  }
  void start_iteration_step(state_type& s) const {
    s.increase_iteration_count();
    before_iteration_step(s);
  }
  void end_iteration_step(state_type& s) const { after_iteration_step(s); }
};

// rescue::failed_eigenproblem (ScfBase.hh:334): post-mortem dump of an eigenproblem the inner solver could not solve -
// the matrix blocks and the overlap, column-major, in Mathematica list syntax like lazyten's io writer.
namespace rescue {
inline void failed_eigenproblem(const std::string& dir, const double* A, size_t n_blocks, const double* S, size_t n) {
  const std::string path = dir + "/failed_eigenproblem.m";
  std::FILE* f = std::fopen(path.c_str(), "w");
  if (!f) return;
  auto dump = [&](const char* name, const double* M) {
    std::fprintf(f, "%s = {", name);
    for (size_t i = 0; i < n; ++i) {
      std::fprintf(f, "%s{", i ? "," : "");
      for (size_t j = 0; j < n; ++j) std::fprintf(f, "%s%.17g", j ? "," : "", M[j * n + i]);
      std::fprintf(f, "}");
    }
    std::fprintf(f, "};\n");
  };
  for (size_t b = 0; b < n_blocks; ++b) dump(b == 0 ? "A" : "Abeta", A + b * n * n);
  dump("S", S);
  std::fclose(f);
}
}  // namespace rescue

}  // namespace lazyten
