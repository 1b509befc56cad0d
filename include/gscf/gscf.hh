// gscf-b200 C++ host layer: gscf's solver classes, state structs, plugin interfaces and control
// keys with the reference's names and semantics (src/gscf/*.hh of molsturm/gscf), header-only,
// C++17.  All SCF arithmetic is executed by libgscf_b200.so through the C ABI (gscf_b200.h):
// solve_state() creates a device engine, registers trampolines for the problem-matrix plugin,
// the error function and the hooks, and runs the whole loop device-resident.
//
// What is identical to the reference: class / template names, public data members of the state
// structs, virtual hooks, GenMap keys and defaults, iteration order, exceptions.  What differs:
// the protected "building block" members (update_eigenpairs, ...) are not individually callable -
// the loop lives behind gscf_b200_solve_{plain,diis,toda}; block-diagonal (alpha/beta) operators
// are driven through the C ABI / Python layer only in this round.
#pragma once
#include <cstring>
#include <exception>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "../gscf_b200.h"
#include "../krims/krims.hh"
#include "../lazyten/lazyten.hh"

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

  real_type last_error_norm;

  ScfStateBase(probmat_type prob_mat, const OverlapMatrix& overlap_mat)
        : lazyten::IterativeStateWrapper<lazyten::SolverStateBase>{lazyten::SolverStateBase()},
          last_error_norm{lazyten::Constants<real_type>::invalid},
          problem_matrix_ptr{std::make_shared<probmat_type>(std::move(prob_mat))},
          diagonalised_matrix_ptr{nullptr},
          m_n_mtx_applies(0),
          m_overlap_matrix_ptr{krims::make_subscription(overlap_mat, "ScfState")},
          m_eigensolution{},
          m_prev_eigensolution{} {
    assert_dbg(problem_matrix_ptr->is_hermitian(100 * lazyten::Constants<real_type>::default_tolerance),
               krims::ExcInvalidState("problem matrix not Hermitian"));
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
  std::shared_ptr<diagmat_type> diagonalised_matrix_ptr;

 private:
  EigenproblemStatistics m_eprob_stats;
  size_t m_n_mtx_applies;
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

// ---- OperatorLinearCombination (OperatorLinearCombination.hh:26-47) ----------------------------------------
template <typename Operator, bool BlockDiagonal = lazyten::IsBlockDiagonalMatrix<Operator>::value>
struct OperatorLinearCombination {
  typedef Operator operator_type;
  typedef typename operator_type::scalar_type scalar_type;
  typedef typename operator_type::real_type real_type;
  typedef typename operator_type::size_type size_type;
  typedef typename operator_type::stored_matrix_type stored_matrix_type;
  explicit OperatorLinearCombination(const operator_type& op, scalar_type coeff = 1) { push_term(op, coeff); }
  void push_term(const operator_type& op, scalar_type coeff = 1) { m_terms.emplace_back(&op, coeff); }
  size_t n_terms() const { return m_terms.size(); }
  size_t n_rows() const { return m_terms.front().first->n_rows(); }
  size_t n_cols() const { return m_terms.front().first->n_cols(); }
  // lazily evaluated element of sum_i c_i F_i (the values the device kernel materialises)
  scalar_type operator()(size_t i, size_t j) const {
    scalar_type s = 0;
    for (const auto& t : m_terms) s += t.second * (*t.first)(i, j);
    return s;
  }

 private:
  std::vector<std::pair<const operator_type*, scalar_type>> m_terms;
};

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

 protected:
  virtual void on_update_eigenpairs(state_type&) const {}
  virtual void on_update_problem_matrix(state_type&) const {}
  /** DIIS / TODA override this (PulayDiisScf.hh:202, TruncatedOptDampScf.hh:200). */
  virtual void on_new_diagmat_dispatch(state_type&) const {}

  /** Residual error C_cur - C_prev (ScfBase.hh:268-289); every real user overrides it.  Solvers
   *  wrapped in b200::DevicePulayError use the built-in device kernel instead. */
  virtual matrix_type calculate_error(const state_type& s) const {
    if (s.previous_eigensolution().evectors().n_vectors() == 0)
      return matrix_type{{lazyten::Constants<scalar_type>::invalid}};
    const auto& prev_evec = s.previous_eigensolution().evectors();
    const auto& cur_evec = s.eigensolution().evectors();
    matrix_type ret(prev_evec.n_elem(), prev_evec.n_vectors(), false);
    for (size_type j = 0; j < cur_evec.n_vectors(); ++j)
      for (size_type i = 0; i < cur_evec.n_elem(); ++i) ret(i, j) = cur_evec[j](i) - prev_evec[j](i);
    return ret;
  }
  virtual bool b200_device_error() const { return false; }

  real_type inner_solver_tolerance(state_type&) const { return std::numeric_limits<real_type>::epsilon(); }

  // ---- engine driver shared by the three algorithms ---------------------------------------------------
  enum class Method { Plain, Diis, Toda };
  struct Driver;
  virtual int b200_history() const { return 2; }
  virtual void b200_fill_params(gscf_b200_params&) const {}
  virtual void b200_after_fock(state_type&) const {}
  virtual void b200_mirror_iteration(state_type&, gscf_b200_engine*, int /*event*/) const {}
  void run_engine(state_type& state, Method method) const;
};

template <typename State>
struct ScfBase<State>::Driver {
  const ScfBase<State>* solver;
  state_type* state;
  gscf_b200_engine* eng = nullptr;
  size_t n = 0;
  int n_occ = 0;
  int mirrored_iter = -1;
  std::exception_ptr error = nullptr;
  Method method = Method::Plain;

  static void check(int st) {
    if (st != GSCF_B200_OK) throw b200::Error(st, gscf_b200_last_error());
  }
  void mirror_eigensolution(const double* C, const double* w, int it) {
    if (mirrored_iter == it) return;
    typename state_type::esoln_type es;
    // n_eigenpairs (ScfBase.hh:87): the eigensolution holds the k lowest pairs only
    const int kq = gscf_b200_get_n_eigenpairs(eng);
    const size_t k = (kq > 0 && (size_t)kq < n) ? (size_t)kq : n;
    es.evalues().assign(w, w + k);
    auto mv = std::make_shared<lazyten::MultiVector<vector_type>>(n, k);
    for (size_t j = 0; j < k; ++j)
      for (size_t i = 0; i < n; ++i) (*mv)[j](i) = C[j * n + i];
    es.evectors_ptr = mv;
    state->push_new_eigensolution(std::move(es), {0, 0});
    mirrored_iter = it;
  }
  void mirror_from_engine(int it) {
    if (mirrored_iter == it) return;
    std::vector<double> C(n * n), w(n);
    check(gscf_b200_get_coeff(eng, C.data(), (int)n));
    check(gscf_b200_get_orben(eng, w.data()));
    mirror_eigensolution(C.data(), w.data(), it);
  }
  // host Fock builder: ScfBase::update_problem_matrix (ScfBase.hh:341-370) incl. copy-on-write
  static int fock_tramp(void* ctx, int, int nb, int, const double* C, const double* w, double* F, double* e1,
                        double* e2) {
    Driver* d = static_cast<Driver*>(ctx);
    try {
      state_type& s = *d->state;
      d->mirror_eigensolution(C, w, (int)s.n_iter());
      if (s.problem_matrix_ptr.use_count() > 1)
        s.problem_matrix_ptr = std::make_shared<probmat_type>(*s.problem_matrix_ptr);
      const std::string key = s.problem_matrix().scf_update_key();
      std::shared_ptr<const lazyten::MultiVector<vector_type>> cptr = s.eigensolution().evectors_ptr;
      s.problem_matrix().update({{key, cptr}});
      d->solver->b200_after_fock(s);
      const probmat_type& pm = s.problem_matrix();
      for (int j = 0; j < nb; ++j)
        for (int i = 0; i < nb; ++i) F[(size_t)j * nb + i] = pm(i, j);
      *e1 = pm.energy_1e_terms();
      *e2 = pm.energy_2e_terms();
      return 0;
    } catch (...) {
      d->error = std::current_exception();
      return 1;
    }
  }
  static int error_tramp(void* ctx, int, int nb, int, const double*, const double* C, const double* w, double* E) {
    Driver* d = static_cast<Driver*>(ctx);
    try {
      state_type& s = *d->state;
      d->mirror_eigensolution(C, w, (int)s.n_iter());
      const matrix_type e = d->solver->calculate_error(s);
      if (e.n_rows() != (size_t)nb || e.n_cols() != (size_t)nb) {
        // e.g. the residual error's 1x1 "invalid" matrix of the first iteration: propagate its value
        for (size_t k = 0; k < (size_t)nb * nb; ++k) E[k] = 0.0;
        E[0] = e.n_rows() ? e(0, 0) : lazyten::Constants<scalar_type>::invalid;
        if (e.n_rows() == s.eigensolution().evectors().n_elem())
          for (size_t j = 0; j < e.n_cols() && j < (size_t)nb; ++j)
            for (size_t i = 0; i < e.n_rows(); ++i) E[j * nb + i] = e(i, j);
        return 0;
      }
      for (int j = 0; j < nb; ++j)
        for (int i = 0; i < nb; ++i) E[(size_t)j * nb + i] = e(i, j);
      return 0;
    } catch (...) {
      d->error = std::current_exception();
      return 1;
    }
  }
  static int hook_tramp(void* ctx, int ev, int it) {
    Driver* d = static_cast<Driver*>(ctx);
    try {
      state_type& s = *d->state;
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
          d->mirror_from_engine(it);
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
void ScfBase<State>::run_engine(state_type& state, Method method) const {
  assert_throw(!state.is_failed(), krims::ExcInvalidState("Cannot solve a failed state"));
  const probmat_type& pm0 = state.problem_matrix();
  const size_t n = pm0.n_rows();
  const krims::Range<size_t> occ_a = pm0.indices_orbspace(OrbitalSpace::OCC_ALPHA);
  const krims::Range<size_t> occ_b = pm0.indices_orbspace(OrbitalSpace::OCC_BETA);
  assert_implemented(occ_a == occ_b);  // restricted problems through the C++ layer (see header note)

  Driver drv{this, &state};
  drv.n = n;
  drv.method = method;
  gscf_b200_config cfg{};
  cfg.n_bas = (int)n; cfg.n_problems = 1; cfg.n_blocks = 1;
  cfg.n_alpha = (int)occ_a.length(); cfg.n_beta = (int)occ_b.length();
  cfg.max_history = std::max(2, b200_history());
  cfg.eig_method = b200_eig_method; cfg.shared_overlap = 0; cfg.stream = nullptr;
  Driver::check(gscf_b200_engine_create(&cfg, &drv.eng));
  struct Guard {
    gscf_b200_engine* e;
    ~Guard() { gscf_b200_engine_destroy(e); }
  } guard{drv.eng};

  std::vector<double> buf(n * n);
  const overlap_type& S = state.overlap_matrix();
  for (size_t j = 0; j < n; ++j)
    for (size_t i = 0; i < n; ++i) buf[j * n + i] = S(i, j);
  Driver::check(gscf_b200_set_overlap(drv.eng, buf.data(), (int)n, GSCF_B200_HOST));

  auto* dev = dynamic_cast<const b200::DeviceFockBuilder_i*>(&pm0);
  if (dev) Driver::check(gscf_b200_set_fock_builder_device(drv.eng, dev->device_fock_fn(), dev->device_fock_ctx()));
  else Driver::check(gscf_b200_set_fock_builder_host(drv.eng, &Driver::fock_tramp, &drv));
  if (!b200_device_error()) Driver::check(gscf_b200_set_error_host(drv.eng, &Driver::error_tramp, &drv));
  Driver::check(gscf_b200_set_hook(drv.eng, &Driver::hook_tramp, &drv));

  for (size_t j = 0; j < n; ++j)
    for (size_t i = 0; i < n; ++i) buf[j * n + i] = pm0(i, j);
  const double e1 = pm0.energy_1e_terms(), e2 = pm0.energy_2e_terms();
  Driver::check(gscf_b200_set_problem_matrix(drv.eng, buf.data(), (int)n, GSCF_B200_HOST, &e1, &e2));

  gscf_b200_params prm;
  gscf_b200_default_params(&prm);
  prm.max_iter = (int)this->max_iter;
  prm.max_error_norm = max_error_norm;
  prm.n_eigenpairs = (n_eigenpairs == lazyten::Constants<size_type>::all || n_eigenpairs >= n) ? -1 : (long long)n_eigenpairs;
  b200_fill_params(prm);
  int st = GSCF_B200_OK;
  switch (method) {
    case Method::Plain: st = gscf_b200_solve_plain(drv.eng, &prm); break;
    case Method::Diis: st = gscf_b200_solve_diis(drv.eng, &prm); break;
    case Method::Toda: st = gscf_b200_solve_toda(drv.eng, &prm); break;
  }
  // mirror the final state
  int nit = 0;
  gscf_b200_get_n_iter(drv.eng, &nit);
  state.set_iteration_count((size_t)nit);
  double err = 0;
  gscf_b200_get_error_norm(drv.eng, &err);
  state.last_error_norm = err;
  long long napp = 0;
  gscf_b200_get_n_mtx_applies(drv.eng, &napp);
  state.n_mtx_applies() = (size_t)napp;
  if (dev) {
    double a = 0, b = 0;
    gscf_b200_get_energies(drv.eng, &a, &b);
    const_cast<b200::DeviceFockBuilder_i*>(dev)->set_energies(a, b);
  }
  if (st == GSCF_B200_ERR_CALLBACK && drv.error) {
    state.fail("callback threw");
    std::rethrow_exception(drv.error);
  }
  if (nit > 0 && st != GSCF_B200_ERR_CALLBACK) drv.mirror_from_engine(nit);
  state.diagonalised_matrix_ptr.reset();
  switch (st) {
    case GSCF_B200_OK: return;
    case GSCF_B200_ERR_MAX_ITER:
      solver_assert(false, state, lazyten::ExcMaximumNumberOfIterationsReached(this->max_iter));
      break;
    case GSCF_B200_ERR_DIIS_STEP: solver_assert(false, state, ExcDiisStepFailed(gscf_b200_last_error())); break;
    case GSCF_B200_ERR_EIGENSOLVER:
      solver_assert(false, state, ExcInnerEigensolverFailed(gscf_b200_last_error()));
      break;
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
  void solve_state(state_type& state) const override { this->run_engine(state, base_type::Method::Plain); }

 protected:
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
    this->run_engine(state, base_type::Method::Diis);
  }

  /** The bordered DIIS system of the current state (PulayDiisScf.hh:326-356). */
  matrix_type diis_linear_system_matrix(const state_type& s) const {
    size_type n_errors = s.error_overlaps.size();
    matrix_type B(n_errors + 1, n_errors + 1, false);
    auto it = std::begin(s.error_overlaps);
    for (size_type i = 0; i < n_errors; ++i, ++it) {
      B(i, i) = (*it)[0];
      for (size_type j = 1; j < i + 1; ++j) B(i, i - j) = B(i - j, i) = (*it)[j];
      B(n_errors, i) = B(i, n_errors) = -1.;
    }
    B(n_errors, n_errors) = 0.;
    return B;
  }

 protected:
  virtual void on_new_diagmat(state_type&) const {}
  void on_new_diagmat_dispatch(state_type& s) const override { on_new_diagmat(s); }
  int b200_history() const override { return (int)n_prev_steps; }
  void b200_fill_params(gscf_b200_params& p) const override { p.diis_n_prev_steps = (int)n_prev_steps; }
  void b200_after_fock(state_type& s) const override {
    s.prev_problem_matrix_ptrs.push_back(s.problem_matrix_ptr);  // PulayDiisScf.hh:315
  }
  void b200_mirror_iteration(state_type& s, gscf_b200_engine* eng, int ev) const override {
    if (ev == GSCF_B200_EV_NEW_DIAGMAT) {
      // diis_coefficients (oldest -> newest) and the lazy linear combination (:358-433)
      std::vector<double> c(16, 0.0);
      int k = 0;
      gscf_b200_get_diis_coefficients(eng, c.data(), &k);
      // the coefficients of *this* iteration live on the device; k == history size
      k = (int)s.prev_problem_matrix_ptrs.size();
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
      const size_t n = s.problem_matrix().n_rows();
      std::vector<double> E(n * n);
      gscf_b200_get_error(eng, E.data(), (int)n);
      matrix_type e(n, n, false);
      std::memcpy(e.memptr(), E.data(), n * n * sizeof(double));
      s.errors.push_back(std::move(e));
      const int k = (int)s.errors.size();
      std::vector<double> ov(16, 0.0);
      int len = 0;
      if (gscf_b200_get_error_overlaps(eng, 0, k - 1, ov.data(), &len) == GSCF_B200_OK)
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
    this->run_engine(state, base_type::Method::Toda);
  }

 protected:
  virtual void on_new_diagmat(state_type&) const {}
  void on_new_diagmat_dispatch(state_type& s) const override { on_new_diagmat(s); }
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
