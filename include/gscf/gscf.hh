// gscf-b200 C++ host layer: gscf's solver classes, state structs, plugin interfaces and control
// keys with the reference's names and semantics (src/gscf/*.hh of molsturm/gscf), header-only,
// C++17.  All SCF arithmetic is executed by libgscf_b200.so through the C ABI (gscf_b200.h):
// solve_state() creates a device engine, registers trampolines for the problem-matrix plugin,
// the error function and the hooks, and runs the whole loop device-resident.
//
// What is identical to the reference: class / template names, public data members of the state
// structs, virtual hooks, GenMap keys and defaults, iteration order, exceptions.  What differs:
// the protected "building block" members (update_eigenpairs, ...) are not individually callable -
// the loop lives behind gscf_b200_solve_{plain,diis,toda}.  Unrestricted problems are block-diagonal
// (alpha / beta) problem matrices (lazyten::BlockDiagonalMatrix<Block, 2>, ScfBase.hh:300-312,
// OperatorLinearCombination.hh:49-84): the eigensolution then holds 2 n vectors of 2 n elements, alpha first.
// Hook / mirror trampolines are registered only when the solver type overrides a hook; otherwise the
// state members (rings, coefficients, eigensolutions) are materialised once, after the loop.
// B200 extension: ScfBase::solve_ensemble runs P independent problems in one engine (lock step).
#pragma once
#include <cstdlib>
#include <cstring>
#include <exception>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "../gscf_b200.h"
#include "../krims/krims.hh"
#include "../lazyten/lazyten.hh"

// GCC's bound-member-function extension: the address a virtual call on *this would reach, against the address of the
// named class's own implementation.  Used to find out whether a hook is overridden (ScfBase::b200_hooks_overridden).
#if defined(__GNUC__) && !defined(__clang__)
#pragma GCC diagnostic ignored "-Wpmf-conversions"
#define GSCF_B200_OVERRIDDEN_(Class, fn) \
  ((b200_hook_fp)(this->*(&Class::fn)) != (b200_hook_fp)(&Class::fn))
#endif

namespace gscf {

// ---- plugin interfaces (ScfProblemMatrix_i.hh:32-44, FocklikeMatrix_i.hh:28-131) --------------------
class ScfProblemMatrix_i {
 public:
  virtual const std::string& scf_update_key() const = 0;
  ScfProblemMatrix_i() = default;
  virtual ~ScfProblemMatrix_i() = default;
  ScfProblemMatrix_i(const ScfProblemMatrix_i&) = default;
  ScfProblemMatrix_i(ScfProblemMatrix_i&&) = default;
  ScfProblemMatrix_i& operator=(ScfProblemMatrix_i&&) = default;
  ScfProblemMatrix_i& operator=(const ScfProblemMatrix_i&) = default;
};

enum class OrbitalSpace { OCC_ALPHA, OCC_BETA, VIRT_ALPHA, VIRT_BETA };

template <typename Scalar>
class FocklikeMatrix_i : public ScfProblemMatrix_i {
 public:
  typedef Scalar scalar_type;
  virtual scalar_type energy_1e_terms() const = 0;
  virtual scalar_type energy_2e_terms() const = 0;
  virtual krims::Range<size_t> indices_orbspace(OrbitalSpace osp) const = 0;
  FocklikeMatrix_i() = default;
  virtual ~FocklikeMatrix_i() = default;
  FocklikeMatrix_i(const FocklikeMatrix_i&) = default;
  FocklikeMatrix_i(FocklikeMatrix_i&&) = default;
  FocklikeMatrix_i& operator=(FocklikeMatrix_i&&) = default;
  FocklikeMatrix_i& operator=(const FocklikeMatrix_i&) = default;
};
template <typename T, typename = void>
struct IsFocklikeMatrix : public std::false_type {};
template <typename T>
struct IsFocklikeMatrix<T, krims::VoidType<typename T::scalar_type>>
      : public std::is_base_of<FocklikeMatrix_i<typename T::scalar_type>, T> {};

// ---- keys (ScfBaseKeys.cc:24-26, PulayDiisScfKeys.cc:24, TruncatedOptDampScfKeys.cc:24-25) -----------
struct ScfBaseKeys : public lazyten::IterativeWrapperKeys {
  inline static const std::string max_iter = "max_iter";
  inline static const std::string n_eigenpairs = "n_eigenpairs";
  inline static const std::string eigensolver_params = "eigensolver";
  inline static const std::string max_error_norm = "max_error_norm";
};
typedef ScfBaseKeys PlainScfKeys;
struct PulayDiisScfKeys : public ScfBaseKeys {
  inline static const std::string n_prev_steps = "diis_n_prev_steps";
};
struct TruncatedOptDampScfKeys : public ScfBaseKeys {
  inline static const std::string n_prev_steps = "toda_n_prev_steps";
  inline static const std::string min_fock_prefactor = "toda_min_fock_prefactor";
};

namespace version {
inline std::string version_string() { return "0.1.0-b200"; }
}  // namespace version

DefSolverException1(ExcInnerEigensolverFailed, std::string, details,
                    << "The SCF procedure failed to converge, since an inner eigensolver failed: " << details);
DefSolverException1(ExcDiisStepFailed, std::string, details,
                    << "The SCF DIIS convergence accelerator step has failed: " << details);

// ---- B200-specific extension points -----------------------------------------------------------------------
namespace b200 {
/** A problem matrix that can build F(C) on the device derives from this in addition to
 *  FocklikeMatrix_i: the engine then never copies C / F through the host. */
class DeviceFockBuilder_i {
 public:
  virtual ~DeviceFockBuilder_i() = default;
  virtual gscf_b200_fock_device_fn device_fock_fn() const = 0;
  virtual void* device_fock_ctx() const = 0;
  virtual void set_energies(double /*e1*/, double /*e2*/) {}
};
struct Error : public krims::ExceptionBase {
  int status;
  Error(int s, const std::string& m) : krims::ExceptionBase("gscf_b200 status " + std::to_string(s) + ": " + m), status(s) {}
};
}  // namespace b200

// ---- state (ScfStateBase.hh:29-305) ------------------------------------------------------------------------
struct EigenproblemStatistics {
  size_t n_iter() const { return m_n_iter; }
  size_t n_mtx_applies() const { return m_n_mtx_applies; }
  EigenproblemStatistics() {}
  EigenproblemStatistics(size_t n_iter, size_t n_mtx_applies) : m_n_iter(n_iter), m_n_mtx_applies(n_mtx_applies) {}

 private:
  size_t m_n_iter = 0;
  size_t m_n_mtx_applies = 0;
};

template <typename ProblemMatrix, typename OverlapMatrix, typename DiagonalisedMatrix = ProblemMatrix>
class ScfStateBase : public lazyten::IterativeStateWrapper<lazyten::SolverStateBase> {
  static_assert(std::is_base_of<ScfProblemMatrix_i, ProblemMatrix>::value,
                "The ProblemMatrix needs to be derived off ScfProblemMatrix_i for SCF algorithms to work.");

 public:
  typedef ProblemMatrix probmat_type;
  typedef OverlapMatrix overlap_type;
  typedef DiagonalisedMatrix diagmat_type;
  typedef typename probmat_type::scalar_type scalar_type;
  typedef typename probmat_type::real_type real_type;
  typedef typename probmat_type::size_type size_type;
  typedef typename probmat_type::stored_matrix_type matrix_type;
  typedef typename matrix_type::vector_type vector_type;
  typedef lazyten::Eigensolution<scalar_type, vector_type> esoln_type;

  const OverlapMatrix& overlap_matrix() const { return *m_overlap_matrix_ptr; }
  const probmat_type& problem_matrix() const { return *problem_matrix_ptr; }
  probmat_type& problem_matrix() { return *problem_matrix_ptr; }
  const diagmat_type& diagonalised_matrix() const { return *diagonalised_matrix_ptr; }
  diagmat_type& diagonalised_matrix() { return *diagonalised_matrix_ptr; }
  const esoln_type& eigensolution() const { return m_eigensolution; }
  const esoln_type& previous_eigensolution() const { return m_prev_eigensolution; }
  const EigenproblemStatistics& eigenproblem_stats() const { return m_eprob_stats; }

  void push_new_eigensolution(esoln_type new_eigensolution, EigenproblemStatistics new_stats) {
    m_prev_eigensolution = std::move(m_eigensolution);
    m_eigensolution = std::move(new_eigensolution);
    m_eprob_stats = new_stats;
  }

  size_t n_mtx_applies() const override { return m_n_mtx_applies; }
  size_t& n_mtx_applies() { return m_n_mtx_applies; }

  // not yet evaluated: compares false against any threshold (ScfStateBase.hh:171)
  real_type last_error_norm = lazyten::Constants<real_type>::invalid;

  /** The state takes the problem matrix over (it is updated in place or copied on write later) and only subscribes
   *  to the overlap, which has to outlive it (ScfStateBase.hh:168-184). */
  ScfStateBase(probmat_type prob_mat, const OverlapMatrix& overlap_mat)
        : problem_matrix_ptr(new probmat_type(std::move(prob_mat))),
          m_overlap_matrix_ptr(krims::make_subscription(overlap_mat, "ScfState")) {
#ifndef NDEBUG
    const real_type herm_tol = 100 * lazyten::Constants<real_type>::default_tolerance;
    if (!problem_matrix_ptr->is_hermitian(herm_tol)) throw krims::ExcInvalidState("problem matrix not Hermitian");
#endif
  }

  template <typename DiagMat>
  void obtain_guess_from(const ScfStateBase<probmat_type, overlap_type, DiagMat>& other) {
    m_prev_eigensolution = other.previous_eigensolution();
    obtain_guess_from(other.eigensolution());
  }
  void obtain_guess_from(esoln_type other_soln) {
    m_eigensolution = std::move(other_soln);
    assert_dbg(m_eigensolution.evectors_ptr != nullptr && m_eigensolution.evectors().n_vectors() > 0,
               krims::ExcInvalidState("Guess solution does not contain valid eigenvectors"));
    const std::string key = problem_matrix_ptr->scf_update_key();
    std::shared_ptr<const lazyten::MultiVector<vector_type>> cptr = m_eigensolution.evectors_ptr;
    problem_matrix_ptr->update({{key, cptr}});  // ScfStateBase.hh:301-302
  }

  std::shared_ptr<probmat_type> problem_matrix_ptr;
  std::shared_ptr<diagmat_type> diagonalised_matrix_ptr;  // nullptr outside an iteration

 private:
  EigenproblemStatistics m_eprob_stats;
  size_t m_n_mtx_applies = 0;
  krims::SubscriptionPointer<const overlap_type> m_overlap_matrix_ptr;
  esoln_type m_eigensolution;
  esoln_type m_prev_eigensolution;
};

template <typename T, typename = void>
struct IsScfState : public std::false_type {};
template <typename T>
struct IsScfState<T, krims::VoidType<typename T::probmat_type, typename T::overlap_type, typename T::diagmat_type>>
      : public std::is_base_of<ScfStateBase<typename T::probmat_type, typename T::overlap_type, typename T::diagmat_type>,
                               T> {};

// ---- OperatorLinearCombination (OperatorLinearCombination.hh:26-84) ----------------------------------------
// sum_i c_i F_i as a lazy sum; what the device materialises with axpby_n.  Primary template: ordinary operators.
template <typename Operator, bool BlockDiagonal = lazyten::IsBlockDiagonalMatrix<Operator>::value>
struct OperatorLinearCombination : public lazyten::LazyMatrixSum<typename Operator::stored_matrix_type> {
  typedef Operator operator_type;
  typedef lazyten::LazyMatrixSum<typename Operator::stored_matrix_type> base_type;
  typedef typename operator_type::scalar_type scalar_type;
  typedef typename operator_type::real_type real_type;
  typedef typename operator_type::size_type size_type;
  typedef typename operator_type::stored_matrix_type stored_matrix_type;
  explicit OperatorLinearCombination(const operator_type& op, scalar_type coeff = 1) : base_type(op, coeff) {}
  void push_term(const operator_type& op, scalar_type coeff = 1) { this->add_term(op, coeff); }
};

// Block-diagonal operators with exactly two (alpha, beta) blocks: one lazy sum per block, every term contributes its
// blocks with the SAME coefficient (OperatorLinearCombination.hh:49-84) - one DIIS / damping coefficient set for both spins.
template <typename Operator>
struct OperatorLinearCombination<Operator, true>
      : public lazyten::BlockDiagonalMatrix<lazyten::LazyMatrixSum<typename Operator::stored_matrix_type>,
                                            Operator::n_blocks> {
  static_assert(Operator::n_blocks == 2, "OperatorLinearCombination of block-diagonal operators assumes two blocks");
  typedef Operator operator_type;
  typedef lazyten::LazyMatrixSum<typename Operator::stored_matrix_type> block_type;
  typedef lazyten::BlockDiagonalMatrix<block_type, Operator::n_blocks> base_type;
  typedef typename operator_type::scalar_type scalar_type;
  typedef typename operator_type::real_type real_type;
  typedef typename operator_type::size_type size_type;
  typedef typename operator_type::stored_matrix_type stored_matrix_type;
  explicit OperatorLinearCombination(const operator_type& op, scalar_type coeff = 1)
        : base_type(std::array<block_type, 2>{block_type(op.diag_blocks()[0], coeff),
                                              block_type(op.diag_blocks()[1], coeff)}) {}
  void push_term(const operator_type& op, scalar_type coeff = 1) {
    for (size_t b = 0; b < Operator::n_blocks; ++b) this->diag_blocks()[b].add_term(op.diag_blocks()[b], coeff);
  }
  size_t n_terms() const { return this->diag_blocks()[0].n_terms(); }
};

namespace b200 {
// uniform block access to ordinary (one block) and block-diagonal (alpha / beta) problem matrices
template <typename PM, bool BD = lazyten::IsBlockDiagonalMatrix<PM>::value>
struct Blocks {
  static constexpr size_t count = 1;
  static size_t order(const PM& m) { return m.n_rows(); }
  static double at(const PM& m, size_t, size_t i, size_t j) { return m(i, j); }
};
template <typename PM>
struct Blocks<PM, true> {
  static constexpr size_t count = PM::n_blocks;
  static size_t order(const PM& m) { return m.diag_blocks()[0].n_rows(); }
  static double at(const PM& m, size_t b, size_t i, size_t j) { return m.diag_blocks()[b](i, j); }
};
}  // namespace b200

// ---- ScfBase (ScfBase.hh:48-370) ------------------------------------------------------------------------------
template <typename State>
class ScfBase : public lazyten::IterativeWrapper<lazyten::SolverBase<State>> {
  static_assert(IsScfState<State>::value, "State needs to be a type derived from ScfStateBase");

 public:
  typedef lazyten::IterativeWrapper<lazyten::SolverBase<State>> base_type;
  typedef typename base_type::state_type state_type;
  typedef typename state_type::probmat_type probmat_type;
  typedef typename state_type::overlap_type overlap_type;
  typedef typename state_type::diagmat_type diagmat_type;
  typedef typename state_type::scalar_type scalar_type;
  typedef typename state_type::real_type real_type;
  typedef typename state_type::size_type size_type;
  typedef typename state_type::matrix_type matrix_type;
  typedef typename state_type::vector_type vector_type;

  size_type n_eigenpairs = lazyten::Constants<size_type>::all;
  krims::GenMap eigensolver_params;
  real_type max_error_norm = 5e-7;
  bool user_eigensolver_tolerance = false;
  /** B200 extension ("b200/eig_method"): gscf_b200_eig_method */
  int b200_eig_method = GSCF_B200_EIG_AUTO;

  void update_control_params(const krims::GenMap& map) {
    base_type::update_control_params(map);
    n_eigenpairs = map.at(ScfBaseKeys::n_eigenpairs, n_eigenpairs);
    max_error_norm = map.at(ScfBaseKeys::max_error_norm, max_error_norm);
    eigensolver_params = map.submap(ScfBaseKeys::eigensolver_params);
    user_eigensolver_tolerance = map.exists("tolerance");  // top-level lookup, ScfBase.hh:105
    b200_eig_method = map.at("b200/eig_method", b200_eig_method);
    // The "eigensolver" sub-map goes to the inner eigensolver (ScfBase.hh:104, :324).  lazyten's key "method" picks among
    // solvers that return the same eigenpairs; here it picks among this library's dense solvers: the values lazyten
    // knows ("auto", "lapack", "arpack") map to the automatic choice, plus "jacobi" and "tridiag_dc".  Its "tolerance"
    // has no meaning for a direct solver (inner_solver_tolerance is machine epsilon anyway, ScfBase.hh:251).
    if (eigensolver_params.exists("method")) {
      const std::string method = eigensolver_params.at("method", std::string("auto"));
      if (method == "jacobi") b200_eig_method = GSCF_B200_EIG_JACOBI;
      else if (method == "tridiag_dc" || method == "dc") b200_eig_method = GSCF_B200_EIG_TRIDIAG_DC;
      else b200_eig_method = GSCF_B200_EIG_AUTO;
    }
  }
  void get_control_params(krims::GenMap& map) const {
    base_type::get_control_params(map);
    map.update(ScfBaseKeys::n_eigenpairs, n_eigenpairs);
    map.update(ScfBaseKeys::max_error_norm, max_error_norm);
    map.update(ScfBaseKeys::eigensolver_params, eigensolver_params);
  }

  bool is_converged(const state_type& s) const override { return s.last_error_norm < max_error_norm; }

  virtual state_type solve(probmat_type probmat_bb, const overlap_type& overlap_bb) const {
    state_type state{std::move(probmat_bb), overlap_bb};
    this->solve_state(state);
    return state;
  }
  template <typename GuessState,
            typename = krims::enable_if_t<std::is_base_of<ScfStateBase<probmat_type, overlap_type, diagmat_type>,
                                                          krims::remove_reference_t<GuessState>>::value ||
                                          IsScfState<krims::remove_reference_t<GuessState>>::value>>
  state_type solve_with_guess(probmat_type probmat_bb, const overlap_type& overlap_bb, GuessState&& guess_state) const {
    state_type state{std::move(probmat_bb), overlap_bb};
    state.obtain_guess_from(guess_state.eigensolution());
    this->solve_state(state);
    return state;
  }
  template <typename GuessState, typename = krims::enable_if_t<IsScfState<krims::remove_reference_t<GuessState>>::value>>
  state_type solve_with_guess(GuessState&& guess_state) const {
    state_type state{guess_state.problem_matrix(), guess_state.overlap_matrix()};
    state.obtain_guess_from(guess_state.eigensolution());
    this->solve_state(state);
    return state;
  }

  /** B200 extension (no counterpart in the reference): P independent problems of the same shape in ONE engine, driven
   *  in lock step (a member leaves the active set when it converges or fails).  overlaps: one per problem or a single
   *  shared one.  No hooks are dispatched; a failed member gets its fail flag set (state.is_failed()) instead of an
   *  exception, so one bad problem does not abort the ensemble. */
  std::vector<state_type> solve_ensemble(std::vector<probmat_type> probmats,
                                         const std::vector<const overlap_type*>& overlaps) const {
    assert_throw(!probmats.empty() && (overlaps.size() == 1 || overlaps.size() == probmats.size()),
                 krims::ExcInvalidState("solve_ensemble: need one overlap per problem or a single shared one"));
    std::vector<state_type> states;
    states.reserve(probmats.size());
    for (size_t p = 0; p < probmats.size(); ++p)
      states.emplace_back(std::move(probmats[p]), *overlaps[overlaps.size() == 1 ? 0 : p]);
    std::vector<state_type*> ptrs;
    for (auto& st : states) ptrs.push_back(&st);
    run_engine(ptrs, /*ensemble=*/true);
    return states;
  }

 protected:
  virtual void on_update_eigenpairs(state_type&) const {}
  virtual void on_update_problem_matrix(state_type&) const {}
  /** DIIS / TODA override this (PulayDiisScf.hh:202, TruncatedOptDampScf.hh:200). */
  virtual void on_new_diagmat_dispatch(state_type&) const {}

  /** Default error (ScfBase.hh:268-289): how far the eigenvectors moved since the previous iteration, column by
   *  column - sign dependent, every real user overrides it.  Before a second eigensolution exists the error is a
   *  1 x 1 "invalid" matrix, i.e. not converged.  Solvers wrapped in b200::DevicePulayError use the device kernel. */
  virtual matrix_type calculate_error(const state_type& s) const {
    const auto& before = s.previous_eigensolution().evectors();
    const auto& now = s.eigensolution().evectors();
    if (before.n_vectors() == 0) {
      matrix_type not_yet(1, 1, false);
      not_yet(0, 0) = lazyten::Constants<scalar_type>::invalid;
      return not_yet;
    }
    const size_type n_vec = std::min(now.n_vectors(), before.n_vectors());
    matrix_type diff(now.n_elem(), n_vec, false);
    for (size_type v = 0; v < n_vec; ++v) {
      const auto& cn = now[v];
      const auto& cb = before[v];
      for (size_type r = 0; r < now.n_elem(); ++r) diff(r, v) = cn(r) - cb(r);
    }
    return diff;
  }
  virtual bool b200_device_error() const { return false; }

  real_type inner_solver_tolerance(state_type&) const { return std::numeric_limits<real_type>::epsilon(); }

  // ---- engine driver shared by the three algorithms ---------------------------------------------------
  enum class Method { Plain, Diis, Toda };
  struct Driver;
  virtual Method b200_method() const = 0;
  virtual int b200_history() const { return 2; }
  virtual void b200_fill_params(gscf_b200_params&) const {}
  virtual void b200_after_fock(state_type&) const {}
  /** per-iteration mirroring of algorithm state (only when a hook is overridden) */
  virtual void b200_mirror_iteration(state_type&, gscf_b200_engine*, int /*event*/) const {}
  /** one-off materialisation of algorithm state after the loop (problem p of the engine) */
  virtual void b200_mirror_final(state_type&, gscf_b200_engine*, int /*p*/, int /*P*/) const {}
  /** Does the dynamic type override any hook?  Decides whether the hook / mirror trampolines are registered at all
   *  (hooks force a device -> host materialisation only if they exist). */
  virtual bool b200_hooks_overridden() const {
#if defined(__GNUC__) && !defined(__clang__)
    return GSCF_B200_OVERRIDDEN_(ScfBase, before_iteration_step) ||
           GSCF_B200_OVERRIDDEN_(ScfBase, after_iteration_step) ||
           GSCF_B200_OVERRIDDEN_(ScfBase, on_update_eigenpairs) ||
           GSCF_B200_OVERRIDDEN_(ScfBase, on_update_problem_matrix);
#else
    return true;  // no portable way to ask: always register
#endif
  }
  void run_engine(std::vector<state_type*>& states, bool ensemble) const;
  void run_engine(state_type& state) const {
    std::vector<state_type*> one{&state};
    run_engine(one, false);
  }
  typedef void (*b200_hook_fp)(const void*, state_type&);
};

template <typename State>
struct ScfBase<State>::Driver {
  typedef b200::Blocks<probmat_type> blocks;
  const ScfBase<State>* solver;
  std::vector<state_type*> states;
  gscf_b200_engine* eng = nullptr;
  size_t n = 0;     // order of one block
  size_t nblk = 1;  // 1 restricted, 2 alpha / beta
  std::vector<int> mirrored_iter;
  std::exception_ptr error = nullptr;

  static void check(int st) {
    if (st != GSCF_B200_OK) throw b200::Error(st, gscf_b200_last_error());
  }
  // the engine owns the iteration counters; without a registered hook nobody else keeps the state's count current
  void sync_iteration_count(state_type& s, int p) {
    std::vector<int> its(states.size(), 0);
    if (gscf_b200_get_n_iter(eng, its.data()) == GSCF_B200_OK) s.set_iteration_count((size_t)its[p]);
  }
  // C: [nblk] n x n column-major, w: [nblk][n]  ->  the state's eigensolution.  Unrestricted: nblk * n vectors of
  // nblk * n elements, alpha block first, zero outside the own block (how lazyten hands out block-diagonal eigenpairs,
  // cf. the index ranges of TruncatedOptDampScf.hh:300-345).
  void mirror_eigensolution(state_type& s, int p, const double* C, const double* w, int it) {
    if (mirrored_iter[p] == it) return;
    typename state_type::esoln_type es;
    const int kq = gscf_b200_get_n_eigenpairs(eng);
    const size_t k = (kq > 0 && (size_t)kq < n) ? (size_t)kq : n;  // n_eigenpairs (ScfBase.hh:87), per block
    auto mv = std::make_shared<lazyten::MultiVector<vector_type>>(n * nblk, k * nblk);
    es.evalues().resize(k * nblk);
    for (size_t b = 0; b < nblk; ++b)
      for (size_t j = 0; j < k; ++j) {
        es.evalues()[b * k + j] = w[b * n + j];
        const double* col = C + (b * n + j) * n;
        for (size_t i = 0; i < n; ++i) (*mv)[b * k + j](b * n + i) = col[i];
      }
    es.evectors_ptr = mv;
    s.push_new_eigensolution(std::move(es), {0, 0});
    mirrored_iter[p] = it;
  }
  // host Fock builder: ScfBase::update_problem_matrix (ScfBase.hh:341-370) incl. copy-on-write
  static int fock_tramp(void* ctx, int p, int nb, int nblocks, const double* C, const double* w, double* F, double* e1,
                        double* e2) {
    Driver* d = static_cast<Driver*>(ctx);
    try {
      state_type& s = *d->states[p];
      d->sync_iteration_count(s, p);
      d->mirror_eigensolution(s, p, C, w, (int)s.n_iter());
      if (s.problem_matrix_ptr.use_count() > 1)
        s.problem_matrix_ptr = std::make_shared<probmat_type>(*s.problem_matrix_ptr);
      const std::string key = s.problem_matrix().scf_update_key();
      std::shared_ptr<const lazyten::MultiVector<vector_type>> cptr = s.eigensolution().evectors_ptr;
      s.problem_matrix().update({{key, cptr}});
      d->solver->b200_after_fock(s);
      const probmat_type& pm = s.problem_matrix();
      for (int b = 0; b < nblocks; ++b)
        for (int j = 0; j < nb; ++j)
          for (int i = 0; i < nb; ++i) F[((size_t)b * nb + j) * nb + i] = blocks::at(pm, b, i, j);
      *e1 = pm.energy_1e_terms();
      *e2 = pm.energy_2e_terms();
      return 0;
    } catch (...) {
      d->error = std::current_exception();
      return 1;
    }
  }
  // calculate_error override (ScfBase.hh:217) -> [nblk] n x n: the diagonal blocks of an (nblk n)^2 error, or the first
  // columns of a tall one (default residual error); a 1 x 1 matrix (first iteration of the residual error) is passed
  // on as its single value
  static int error_tramp(void* ctx, int p, int nb, int nblocks, const double*, const double* C, const double* w,
                         double* E) {
    Driver* d = static_cast<Driver*>(ctx);
    try {
      state_type& s = *d->states[p];
      d->sync_iteration_count(s, p);
      d->mirror_eigensolution(s, p, C, w, (int)s.n_iter());
      const matrix_type e = d->solver->calculate_error(s);
      const size_t N = (size_t)nb * nblocks;
      for (size_t k = 0; k < (size_t)nblocks * nb * nb; ++k) E[k] = 0.0;
      if (e.n_rows() != N) {
        E[0] = e.n_rows() ? e(0, 0) : lazyten::Constants<scalar_type>::invalid;
        return 0;
      }
      for (int b = 0; b < nblocks; ++b)
        for (int j = 0; j < nb; ++j) {
          const size_t col = (size_t)b * nb + j;
          if (col >= e.n_cols()) continue;
          for (int i = 0; i < nb; ++i) E[((size_t)b * nb + j) * nb + i] = e((size_t)b * nb + i, col);
        }
      return 0;
    } catch (...) {
      d->error = std::current_exception();
      return 1;
    }
  }
  void mirror_from_engine(state_type& s, int p, int it) {
    if (mirrored_iter[p] == it) return;
    const size_t P = states.size();
    std::vector<double> C(P * nblk * n * n), w(P * nblk * n);
    check(gscf_b200_get_coeff(eng, C.data(), (int)n));
    check(gscf_b200_get_orben(eng, w.data()));
    mirror_eigensolution(s, p, C.data() + (size_t)p * nblk * n * n, w.data() + (size_t)p * nblk * n, it);
  }
  static int hook_tramp(void* ctx, int ev, int it) {  // single problem only
    Driver* d = static_cast<Driver*>(ctx);
    try {
      state_type& s = *d->states[0];
      s.set_iteration_count((size_t)it);
      switch (ev) {
        case GSCF_B200_EV_BEFORE_ITERATION:
          d->solver->before_iteration_step(s);
          break;
        case GSCF_B200_EV_NEW_DIAGMAT:
          d->solver->b200_mirror_iteration(s, d->eng, ev);
          d->solver->on_new_diagmat_dispatch(s);
          break;
        case GSCF_B200_EV_UPDATE_EIGENPAIRS:
          d->mirror_from_engine(s, 0, it);
          d->solver->b200_mirror_iteration(s, d->eng, ev);
          d->solver->on_update_eigenpairs(s);
          s.diagonalised_matrix_ptr.reset();
          break;
        case GSCF_B200_EV_UPDATE_PROBLEM_MATRIX:
          d->solver->on_update_problem_matrix(s);
          break;
        case GSCF_B200_EV_AFTER_ITERATION: {
          double err = 0;
          check(gscf_b200_get_error_norm(d->eng, &err));
          s.last_error_norm = err;
          d->solver->b200_mirror_iteration(s, d->eng, ev);
          d->solver->after_iteration_step(s);
          break;
        }
      }
      return 0;
    } catch (...) {
      d->error = std::current_exception();
      return 1;
    }
  }
};

template <typename State>
void ScfBase<State>::run_engine(std::vector<state_type*>& states, bool ensemble) const {
  typedef b200::Blocks<probmat_type> blocks;
  const size_t P = states.size();
  for (state_type* sp : states) assert_throw(!sp->is_failed(), krims::ExcInvalidState("Cannot solve a failed state"));
  const probmat_type& pm0 = states[0]->problem_matrix();
  const size_t nblk = blocks::count;
  const size_t n = blocks::order(pm0);
  const krims::Range<size_t> occ_a = pm0.indices_orbspace(OrbitalSpace::OCC_ALPHA);
  const krims::Range<size_t> occ_b = pm0.indices_orbspace(OrbitalSpace::OCC_BETA);
  // restricted: one block, identical ranges.  Unrestricted (ScfBase.hh:300-312): a two-block problem matrix; the
  // beta range indexes the second half of the eigensolution.  Anything else has no representation in the engine.
  assert_implemented(nblk == 2 || occ_a == occ_b);
  assert_implemented(nblk == 1 || n_eigenpairs == lazyten::Constants<size_type>::all || n_eigenpairs >= n * nblk);

  Driver drv{this, states};
  drv.n = n;
  drv.nblk = nblk;
  drv.mirrored_iter.assign(P, -1);
  gscf_b200_config cfg{};
  cfg.n_bas = (int)n; cfg.n_problems = (int)P; cfg.n_blocks = (int)nblk;
  cfg.n_alpha = (int)occ_a.length(); cfg.n_beta = (int)occ_b.length();
  cfg.max_history = std::max(2, b200_history());
  cfg.eig_method = b200_eig_method; cfg.stream = nullptr;
  bool shared = true;
  for (size_t p = 1; p < P; ++p) shared = shared && (&states[p]->overlap_matrix() == &states[0]->overlap_matrix());
  cfg.shared_overlap = (P > 1 && shared) ? 1 : 0;
  Driver::check(gscf_b200_engine_create(&cfg, &drv.eng));
  struct Guard {
    gscf_b200_engine* e;
    ~Guard() { gscf_b200_engine_destroy(e); }
  } guard{drv.eng};

  {  // overlap: the (first) n x n diagonal block of every state's overlap
    const size_t nS = cfg.shared_overlap ? 1 : P;
    std::vector<double> buf(nS * n * n);
    for (size_t p = 0; p < nS; ++p) {
      const overlap_type& S = states[p]->overlap_matrix();
      for (size_t j = 0; j < n; ++j)
        for (size_t i = 0; i < n; ++i) buf[(p * n + j) * n + i] = S(i, j);
    }
    Driver::check(gscf_b200_set_overlap(drv.eng, buf.data(), (int)n, GSCF_B200_HOST));
  }

  auto* dev = dynamic_cast<const b200::DeviceFockBuilder_i*>(&pm0);
  if (dev) Driver::check(gscf_b200_set_fock_builder_device(drv.eng, dev->device_fock_fn(), dev->device_fock_ctx()));
  else Driver::check(gscf_b200_set_fock_builder_host(drv.eng, &Driver::fock_tramp, &drv));
  if (!b200_device_error()) Driver::check(gscf_b200_set_error_host(drv.eng, &Driver::error_tramp, &drv));
  const bool hooks = !ensemble && b200_hooks_overridden();
  if (hooks) Driver::check(gscf_b200_set_hook(drv.eng, &Driver::hook_tramp, &drv));

  {  // the problem matrices handed to solve() with their energies
    std::vector<double> buf(P * nblk * n * n), e1(P), e2(P);
    for (size_t p = 0; p < P; ++p) {
      const probmat_type& pm = states[p]->problem_matrix();
      for (size_t b = 0; b < nblk; ++b)
        for (size_t j = 0; j < n; ++j)
          for (size_t i = 0; i < n; ++i) buf[((p * nblk + b) * n + j) * n + i] = blocks::at(pm, b, i, j);
      e1[p] = pm.energy_1e_terms();
      e2[p] = pm.energy_2e_terms();
    }
    Driver::check(gscf_b200_set_problem_matrix(drv.eng, buf.data(), (int)n, GSCF_B200_HOST, e1.data(), e2.data()));
  }

  gscf_b200_params prm;
  gscf_b200_default_params(&prm);
  prm.max_iter = (int)this->max_iter;
  prm.max_error_norm = max_error_norm;
  prm.n_eigenpairs = (n_eigenpairs == lazyten::Constants<size_type>::all || n_eigenpairs >= n) ? -1 : (long long)n_eigenpairs;
  b200_fill_params(prm);
  int st = GSCF_B200_OK;
  switch (b200_method()) {
    case Method::Plain: st = gscf_b200_solve_plain(drv.eng, &prm); break;
    case Method::Diis: st = gscf_b200_solve_diis(drv.eng, &prm); break;
    case Method::Toda: st = gscf_b200_solve_toda(drv.eng, &prm); break;
  }
  // materialise the final state(s)
  std::vector<int> nit(P, 0), status(P, 0);
  std::vector<double> err(P, 0.0), ea(P, 0.0), eb(P, 0.0);
  std::vector<long long> napp(P, 0);
  gscf_b200_get_n_iter(drv.eng, nit.data());
  gscf_b200_get_status(drv.eng, status.data());
  gscf_b200_get_error_norm(drv.eng, err.data());
  gscf_b200_get_n_mtx_applies(drv.eng, napp.data());
  for (size_t p = 0; p < P; ++p) {
    states[p]->set_iteration_count((size_t)nit[p]);
    states[p]->last_error_norm = err[p];
    states[p]->n_mtx_applies() = (size_t)napp[p];
  }
  if (dev) {
    gscf_b200_get_energies(drv.eng, ea.data(), eb.data());
    const_cast<b200::DeviceFockBuilder_i*>(dev)->set_energies(ea[0], eb[0]);
  }
  if (st == GSCF_B200_ERR_CALLBACK && drv.error) {
    for (state_type* sp : states) sp->fail("callback threw");
    std::rethrow_exception(drv.error);
  }
  if (st != GSCF_B200_ERR_CALLBACK) {
    std::vector<double> C, w;
    for (size_t p = 0; p < P; ++p) {
      if (nit[p] <= 0 || drv.mirrored_iter[p] == nit[p]) continue;
      if (C.empty()) {  // one device -> host copy for the whole ensemble
        C.resize(P * nblk * n * n);
        w.resize(P * nblk * n);
        Driver::check(gscf_b200_get_coeff(drv.eng, C.data(), (int)n));
        Driver::check(gscf_b200_get_orben(drv.eng, w.data()));
      }
      drv.mirror_eigensolution(*states[p], (int)p, C.data() + p * nblk * n * n, w.data() + p * nblk * n, nit[p]);
    }
    if (!hooks)
      for (size_t p = 0; p < P; ++p) b200_mirror_final(*states[p], drv.eng, (int)p, (int)P);
  }
  for (state_type* sp : states) sp->diagonalised_matrix_ptr.reset();
  auto message = [&](int code) -> std::string { return std::string(gscf_b200_last_error()) + " (status " + std::to_string(code) + ")"; };
  if (ensemble) {
    for (size_t p = 0; p < P; ++p)
      if (status[p] != GSCF_B200_OK) states[p]->fail(message(status[p]));
    if (st != GSCF_B200_OK && st != GSCF_B200_ERR_MAX_ITER && st != GSCF_B200_ERR_DIIS_STEP && st != GSCF_B200_ERR_EIGENSOLVER)
      throw b200::Error(st, gscf_b200_last_error());
    return;
  }
  state_type& state = *states[0];
  switch (st) {
    case GSCF_B200_OK: return;
    case GSCF_B200_ERR_MAX_ITER:
      solver_assert(false, state, lazyten::ExcMaximumNumberOfIterationsReached(this->max_iter));
      break;
    case GSCF_B200_ERR_DIIS_STEP: solver_assert(false, state, ExcDiisStepFailed(gscf_b200_last_error())); break;
    case GSCF_B200_ERR_EIGENSOLVER: {
      // lazyten::rescue::failed_eigenproblem (ScfBase.hh:334): dump the eigenproblem that failed (only when the
      // environment names a directory: GSCF_B200_RESCUE_DIR)
      const std::string why = gscf_b200_last_error();
      if (const char* dir = std::getenv("GSCF_B200_RESCUE_DIR")) {
        std::vector<double> Fd(nblk * n * n), Sd(n * n);
        if (gscf_b200_get_diagmat(drv.eng, Fd.data(), (int)n) == GSCF_B200_OK) {
          const overlap_type& S = state.overlap_matrix();
          for (size_t j = 0; j < n; ++j)
            for (size_t i = 0; i < n; ++i) Sd[j * n + i] = S(i, j);
          lazyten::rescue::failed_eigenproblem(std::string(dir), Fd.data(), nblk, Sd.data(), n);
        }
      }
      solver_assert(false, state, ExcInnerEigensolverFailed(why));
      break;
    }
    default: state.fail(gscf_b200_last_error()); throw b200::Error(st, gscf_b200_last_error());
  }
}

// ---- PlainScf (PlainScf.hh:31-147) ---------------------------------------------------------------------------------
template <typename ProblemMatrix, typename OverlapMatrix>
struct PlainScfState : public ScfStateBase<ProblemMatrix, OverlapMatrix, ProblemMatrix> {
  typedef ScfStateBase<ProblemMatrix, OverlapMatrix, ProblemMatrix> base_type;
  typedef typename base_type::probmat_type probmat_type;
  typedef typename base_type::overlap_type overlap_type;
  typedef typename base_type::diagmat_type diagmat_type;
  typedef typename base_type::scalar_type scalar_type;
  typedef typename base_type::size_type size_type;
  typedef typename base_type::matrix_type matrix_type;
  typedef typename base_type::vector_type vector_type;
  PlainScfState(probmat_type probmat, const overlap_type& overlap_mat) : base_type{std::move(probmat), overlap_mat} {}
};

template <typename ScfState>
class PlainScf : public ScfBase<ScfState> {
 public:
  typedef ScfBase<ScfState> base_type;
  typedef typename base_type::state_type state_type;
  typedef typename state_type::probmat_type probmat_type;
  typedef typename state_type::overlap_type overlap_type;
  typedef typename state_type::real_type real_type;
  typedef typename state_type::vector_type vector_type;
  typedef typename state_type::matrix_type matrix_type;
  static_assert(std::is_base_of<PlainScfState<probmat_type, overlap_type>, ScfState>::value,
                "ScfState needs to be derived off PlainScfState");
  PlainScf() {}
  PlainScf(const krims::GenMap& map) : PlainScf() { update_control_params(map); }
  void update_control_params(const krims::GenMap& map) { base_type::update_control_params(map); }
  void get_control_params(krims::GenMap& map) const { base_type::get_control_params(map); }
  void solve_state(state_type& state) const override { this->run_engine(state); }

 protected:
  typename base_type::Method b200_method() const override { return base_type::Method::Plain; }
  void b200_mirror_iteration(state_type& s, gscf_b200_engine*, int ev) const override {
    if (ev == GSCF_B200_EV_UPDATE_EIGENPAIRS) s.diagonalised_matrix_ptr = s.problem_matrix_ptr;  // PlainScf.hh:129
  }
};

// ---- PulayDiisScf (PulayDiisScf.hh:38-463) -------------------------------------------------------------------------
template <typename ProblemMatrix, typename OverlapMatrix>
struct PulayDiisScfState
      : public ScfStateBase<ProblemMatrix, OverlapMatrix, OperatorLinearCombination<ProblemMatrix>> {
  typedef ScfStateBase<ProblemMatrix, OverlapMatrix, OperatorLinearCombination<ProblemMatrix>> base_type;
  typedef typename base_type::probmat_type probmat_type;
  typedef typename base_type::overlap_type overlap_type;
  typedef typename base_type::diagmat_type diagmat_type;
  typedef typename base_type::scalar_type scalar_type;
  typedef typename base_type::real_type real_type;
  typedef typename base_type::size_type size_type;
  typedef typename base_type::matrix_type matrix_type;
  typedef typename base_type::vector_type vector_type;
  typedef typename base_type::esoln_type esoln_type;

  krims::CircularBuffer<std::shared_ptr<const probmat_type>> prev_problem_matrix_ptrs;
  krims::CircularBuffer<matrix_type> errors;
  krims::CircularBuffer<std::vector<scalar_type>> error_overlaps;
  vector_type diis_coefficients;

  PulayDiisScfState(probmat_type probmat, const overlap_type& overlap_mat)
        : base_type{std::move(probmat), overlap_mat},
          prev_problem_matrix_ptrs{0},
          errors{0},
          error_overlaps{0},
          diis_coefficients(0) {}
  void resize_buffers(const size_type n_prev_steps) {
    prev_problem_matrix_ptrs.max_size(n_prev_steps);
    errors.max_size(n_prev_steps);
    error_overlaps.max_size(n_prev_steps);
  }
};

template <typename ScfState>
class PulayDiisScf : public ScfBase<ScfState> {
 public:
  typedef ScfBase<ScfState> base_type;
  typedef typename base_type::state_type state_type;
  typedef typename state_type::probmat_type probmat_type;
  typedef typename state_type::overlap_type overlap_type;
  typedef typename state_type::diagmat_type diagmat_type;
  typedef typename state_type::scalar_type scalar_type;
  typedef typename state_type::real_type real_type;
  typedef typename state_type::size_type size_type;
  typedef typename state_type::matrix_type matrix_type;
  typedef typename state_type::vector_type vector_type;
  static_assert(std::is_base_of<PulayDiisScfState<probmat_type, overlap_type>, ScfState>::value,
                "ScfState needs to be derived off PulayDiisScfState");

  PulayDiisScf() {}
  PulayDiisScf(const krims::GenMap& map) : PulayDiisScf() { update_control_params(map); }
  size_type n_prev_steps = 4;
  void update_control_params(const krims::GenMap& map) {
    base_type::update_control_params(map);
    n_prev_steps = map.at(PulayDiisScfKeys::n_prev_steps, n_prev_steps);
  }
  void get_control_params(krims::GenMap& map) const {
    base_type::get_control_params(map);
    map.update(PulayDiisScfKeys::n_prev_steps, n_prev_steps);
  }
  void solve_state(state_type& state) const override {
    state.resize_buffers(n_prev_steps);  // PulayDiisScf.hh:289
    this->run_engine(state);
  }
  /** B200 extension, see ScfBase::solve_ensemble. */
  std::vector<state_type> solve_ensemble(std::vector<probmat_type> probmats,
                                         const std::vector<const overlap_type*>& overlaps) const {
    std::vector<state_type> res = base_type::solve_ensemble(std::move(probmats), overlaps);
    return res;
  }

  /** The bordered DIIS system of the current state (PulayDiisScf.hh:326-356): entry (row, col <= row) is the overlap of
   *  error `row` with error `col`, stored newest-first in overlap vector `row` at position row - col; the border row
   *  and column are -1, the corner stays 0. */
  matrix_type diis_linear_system_matrix(const state_type& s) const {
    std::vector<const std::vector<scalar_type>*> ov;
    for (const auto& v : s.error_overlaps) ov.push_back(&v);
    const size_type k = ov.size();
    matrix_type B(k + 1, k + 1, true);
    for (size_type row = 0; row < k; ++row) {
      for (size_type col = 0; col <= row; ++col) {
        const scalar_type v = (*ov[row])[row - col];
        B(row, col) = v;
        B(col, row) = v;
      }
      B(row, k) = -1.;
      B(k, row) = -1.;
    }
    return B;
  }

 protected:
  virtual void on_new_diagmat(state_type&) const {}
  void on_new_diagmat_dispatch(state_type& s) const override { on_new_diagmat(s); }
  typename base_type::Method b200_method() const override { return base_type::Method::Diis; }
  bool b200_hooks_overridden() const override {
#if defined(__GNUC__) && !defined(__clang__)
    typedef typename base_type::b200_hook_fp b200_hook_fp;
    return base_type::b200_hooks_overridden() || GSCF_B200_OVERRIDDEN_(PulayDiisScf, on_new_diagmat);
#else
    return true;
#endif
  }
  int b200_history() const override { return (int)n_prev_steps; }
  void b200_fill_params(gscf_b200_params& p) const override { p.diis_n_prev_steps = (int)n_prev_steps; }
  void b200_after_fock(state_type& s) const override {
    if (s.prev_problem_matrix_ptrs.max_size() != n_prev_steps) s.resize_buffers(n_prev_steps);
    s.prev_problem_matrix_ptrs.push_back(s.problem_matrix_ptr);  // PulayDiisScf.hh:315
  }
  // [nblk] n x n blocks from the engine -> the state's error matrix ((nblk n)^2, block diagonal)
  static matrix_type b200_error_matrix(const double* E, size_t n, size_t nblk) {
    matrix_type e(n * nblk, n * nblk, true);
    for (size_t b = 0; b < nblk; ++b)
      for (size_t j = 0; j < n; ++j)
        for (size_t i = 0; i < n; ++i) e(b * n + i, b * n + j) = E[(b * n + j) * n + i];
    return e;
  }
  void b200_mirror_iteration(state_type& s, gscf_b200_engine* eng, int ev) const override {
    typedef b200::Blocks<probmat_type> blocks;
    if (ev == GSCF_B200_EV_NEW_DIAGMAT) {
      // diis_coefficients (oldest -> newest) of THIS iteration and the lazy linear combination (:358-433)
      const int k = (int)s.prev_problem_matrix_ptrs.size();
      std::vector<double> coeff(k, 0.0);
      if (k == 1) coeff[0] = 1.0;
      if (k >= 2) gscf_b200_get_current_diis_coefficients(eng, coeff.data(), k);
      s.diis_coefficients = vector_type(coeff.begin(), coeff.end());
      if (k == 0) {
        s.diagonalised_matrix_ptr = std::make_shared<diagmat_type>(s.problem_matrix());
      } else {
        auto it = std::begin(s.prev_problem_matrix_ptrs);
        s.diagonalised_matrix_ptr = std::make_shared<diagmat_type>(**it, coeff[0]);
        size_t i = 1;
        for (++it; it != std::end(s.prev_problem_matrix_ptrs); ++it, ++i)
          s.diagonalised_matrix_ptr->push_term(**it, coeff[i]);
      }
    } else if (ev == GSCF_B200_EV_AFTER_ITERATION) {
      // errors.push_back / append_new_overlaps (:318-320, :440-463)
      const size_t n = blocks::order(s.problem_matrix()), nblk = blocks::count;
      std::vector<double> E(nblk * n * n);
      gscf_b200_get_error(eng, E.data(), (int)n);
      s.errors.push_back(b200_error_matrix(E.data(), n, nblk));
      const int k = (int)s.errors.size();
      std::vector<double> ov(16, 0.0);
      int len = 0;
      if (gscf_b200_get_error_overlaps(eng, 0, k - 1, ov.data(), &len) == GSCF_B200_OK)
        s.error_overlaps.push_back(std::vector<scalar_type>(ov.begin(), ov.begin() + len));
    }
  }
  // no hook overridden: the rings and the last coefficients are fetched once, after the loop
  void b200_mirror_final(state_type& s, gscf_b200_engine* eng, int p, int P) const override {
    typedef b200::Blocks<probmat_type> blocks;
    const size_t n = blocks::order(s.problem_matrix()), nblk = blocks::count;
    if (s.errors.max_size() != n_prev_steps) s.resize_buffers(n_prev_steps);
    const int k = gscf_b200_get_history_size(eng);
    std::vector<double> c((size_t)P * 16, 0.0);
    int kc = 0;
    if (gscf_b200_get_diis_coefficients(eng, c.data(), &kc) == GSCF_B200_OK && kc > 0)
      s.diis_coefficients = vector_type(c.begin() + (size_t)p * 16, c.begin() + (size_t)p * 16 + kc);
    std::vector<double> E((size_t)P * nblk * n * n);
    for (int i = 0; i < k; ++i) {
      if (gscf_b200_get_error_entry(eng, i, E.data(), (int)n) != GSCF_B200_OK) break;
      s.errors.push_back(b200_error_matrix(E.data() + (size_t)p * nblk * n * n, n, nblk));
      std::vector<double> ov(16, 0.0);
      int len = 0;
      if (gscf_b200_get_error_overlaps(eng, p, i, ov.data(), &len) == GSCF_B200_OK)
        s.error_overlaps.push_back(std::vector<scalar_type>(ov.begin(), ov.begin() + len));
    }
  }
};

// ---- TruncatedOptDampScf (TruncatedOptDampScf.hh:39-409) -----------------------------------------------------------
template <typename ProblemMatrix, typename OverlapMatrix>
struct TruncatedOptDampScfState
      : public ScfStateBase<ProblemMatrix, OverlapMatrix, OperatorLinearCombination<ProblemMatrix>> {
  static_assert(IsFocklikeMatrix<ProblemMatrix>::value,
                "The ProblemMatrix has to be a fock-like matrix for this SCF algorithm.");
  typedef ScfStateBase<ProblemMatrix, OverlapMatrix, OperatorLinearCombination<ProblemMatrix>> base_type;
  typedef typename base_type::probmat_type probmat_type;
  typedef typename base_type::overlap_type overlap_type;
  typedef typename base_type::diagmat_type diagmat_type;
  typedef typename base_type::scalar_type scalar_type;
  typedef typename base_type::real_type real_type;
  typedef typename base_type::size_type size_type;
  typedef typename base_type::matrix_type matrix_type;
  typedef typename base_type::vector_type vector_type;
  typedef typename base_type::esoln_type esoln_type;
  std::shared_ptr<const probmat_type> prev_problem_matrix_ptr;
  scalar_type damping_coeff;
  TruncatedOptDampScfState(probmat_type probmat, const overlap_type& overlap_mat)
        : base_type{std::move(probmat), overlap_mat}, prev_problem_matrix_ptr{nullptr}, damping_coeff{1} {}
};

template <typename ScfState>
class TruncatedOptDampScf : public ScfBase<ScfState> {
 public:
  typedef ScfBase<ScfState> base_type;
  typedef typename base_type::state_type state_type;
  typedef typename state_type::probmat_type probmat_type;
  typedef typename state_type::overlap_type overlap_type;
  typedef typename state_type::diagmat_type diagmat_type;
  typedef typename state_type::scalar_type scalar_type;
  typedef typename state_type::real_type real_type;
  typedef typename state_type::size_type size_type;
  typedef typename state_type::matrix_type matrix_type;
  typedef typename state_type::vector_type vector_type;
  static_assert(std::is_base_of<TruncatedOptDampScfState<probmat_type, overlap_type>, ScfState>::value,
                "ScfState needs to be derived off TruncatedOptDampScfState");

  TruncatedOptDampScf() {}
  TruncatedOptDampScf(const krims::GenMap& map) : TruncatedOptDampScf() { update_control_params(map); }
  size_type n_prev_steps = 2;
  real_type min_fock_prefactor = 0.05;
  void update_control_params(const krims::GenMap& map) {
    base_type::update_control_params(map);
    n_prev_steps = map.at(TruncatedOptDampScfKeys::n_prev_steps, n_prev_steps);
    min_fock_prefactor = map.at(TruncatedOptDampScfKeys::min_fock_prefactor, min_fock_prefactor);
  }
  void get_control_params(krims::GenMap& map) const {
    base_type::get_control_params(map);
    map.update(TruncatedOptDampScfKeys::n_prev_steps, n_prev_steps);
    map.update(TruncatedOptDampScfKeys::min_fock_prefactor, min_fock_prefactor);
  }
  void solve_state(state_type& state) const override {
    assert_implemented(n_prev_steps == 2);  // TruncatedOptDampScf.hh:243
    this->run_engine(state);
  }

 protected:
  virtual void on_new_diagmat(state_type&) const {}
  void on_new_diagmat_dispatch(state_type& s) const override { on_new_diagmat(s); }
  typename base_type::Method b200_method() const override { return base_type::Method::Toda; }
  bool b200_hooks_overridden() const override {
#if defined(__GNUC__) && !defined(__clang__)
    typedef typename base_type::b200_hook_fp b200_hook_fp;
    return base_type::b200_hooks_overridden() || GSCF_B200_OVERRIDDEN_(TruncatedOptDampScf, on_new_diagmat);
#else
    return true;
#endif
  }
  void b200_mirror_final(state_type& s, gscf_b200_engine* eng, int p, int P) const override {
    std::vector<double> damp((size_t)P, 1.0);
    if (gscf_b200_get_damping_coeff(eng, damp.data()) == GSCF_B200_OK) s.damping_coeff = damp[p];
  }
  void b200_fill_params(gscf_b200_params& p) const override {
    p.toda_n_prev_steps = (int)n_prev_steps;
    p.toda_min_fock_prefactor = min_fock_prefactor;
  }
  void b200_mirror_iteration(state_type& s, gscf_b200_engine* eng, int ev) const override {
    if (ev == GSCF_B200_EV_NEW_DIAGMAT) {
      double damp = 1.0;
      gscf_b200_get_damping_coeff(eng, &damp);
      s.damping_coeff = damp;
      if (damp == 1) s.prev_problem_matrix_ptr.reset();  // :367
      if (s.prev_problem_matrix_ptr == nullptr) {        // :396-409
        s.diagonalised_matrix_ptr = std::make_shared<diagmat_type>(s.problem_matrix());
      } else {
        s.diagonalised_matrix_ptr = std::make_shared<diagmat_type>(*s.prev_problem_matrix_ptr, 1. - damp);
        s.diagonalised_matrix_ptr->push_term(s.problem_matrix(), damp);
      }
    } else if (ev == GSCF_B200_EV_UPDATE_EIGENPAIRS) {
      if (s.damping_coeff > 0.5) s.prev_problem_matrix_ptr = s.problem_matrix_ptr;  // :273-277
      else s.damping_coeff = 1 - s.damping_coeff;
    }
  }
};

// ---- device Pulay error mixin (the GPU twin of gscfmock::error_wrapped_solvers::ErrorWrapped) -------------------------
namespace b200 {
template <typename Solver>
class DevicePulayError : public Solver {
 public:
  using Solver::Solver;

 protected:
  bool b200_device_error() const override { return true; }
};
}  // namespace b200

}  // namespace gscf
