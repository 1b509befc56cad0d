/* gscf_b200.h - C ABI of the B200-native SCF-iteration hot path of molsturm/gscf.
 *
 * This is the drop-in boundary: a host program (the header-only C++ layer in
 * include/gscf/, or any FFI) drives the SCF path through these plain-C entry points.
 * No torch / C++ types cross this boundary: pointers, sizes, POD structs only.
 *
 * Reference interfaces replaced (paths relative to the molsturm/gscf checkout):
 *   - src/gscf/ScfBase.hh:291-339          update_eigenpairs        -> gscf_b200_eigh_generalised,
 *                                                                        engine eigensolve step
 *   - src/gscf/ScfBase.hh:341-370          update_problem_matrix    -> fock-builder callbacks
 *   - src/gscfmock/FockMatrix.cc:198-203   density C_o C_o^T        -> gscf_b200_density
 *   - src/gscfmock/pulay_error.cc:25-44    Pulay error              -> gscf_b200_pulay_error
 *   - src/gscf/PulayDiisScf.hh:326-463     DIIS B matrix / solve /
 *                                          extrapolation / overlaps -> gscf_b200_frob_multidot,
 *                                                                        gscf_b200_diis_solve,
 *                                                                        gscf_b200_axpby_n
 *   - src/gscf/TruncatedOptDampScf.hh:291-409  trace + line search  -> gscf_b200_toda_trace,
 *                                                                        gscf_b200_compute_damping_coeff
 *   - src/gscf/{PlainScf,PulayDiisScf,TruncatedOptDampScf}.hh solve_state
 *                                                                    -> gscf_b200_solve_{plain,diis,toda}
 *
 * All matrices are fp64, column-major.  "Device" pointers are CUDA device pointers of the
 * current device.  Every function returns a gscf_b200_status; gscf_b200_last_error() gives a
 * thread-local description of the most recent failure.  There is no CPU fallback: without a
 * CUDA device every compute entry point returns GSCF_B200_ERR_CUDA.
 */
#ifndef GSCF_B200_H
#define GSCF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSCF_B200_VERSION_MAJOR 0
#define GSCF_B200_VERSION_MINOR 1

typedef enum gscf_b200_status {
  GSCF_B200_OK = 0,
  GSCF_B200_ERR_INVALID_ARG = 1,
  GSCF_B200_ERR_CUDA = 2,
  GSCF_B200_ERR_EIGENSOLVER = 3,  /* ExcInnerEigensolverFailed, ScfBase.hh:33-36 */
  GSCF_B200_ERR_DIIS_STEP = 4,    /* ExcDiisStepFailed, PulayDiisScf.hh:30-32 */
  GSCF_B200_ERR_MAX_ITER = 5,     /* lazyten ExcMaximumNumberOfIterationsReached */
  GSCF_B200_ERR_INVALID_STATE = 6,/* krims::ExcInvalidState ("cannot solve a failed state") */
  GSCF_B200_ERR_CALLBACK = 7,     /* a user callback returned non-zero */
  GSCF_B200_ERR_NOT_IMPLEMENTED = 8
} gscf_b200_status;

typedef enum gscf_b200_memloc { GSCF_B200_HOST = 0, GSCF_B200_DEVICE = 1 } gscf_b200_memloc;

typedef enum gscf_b200_eig_method {
  GSCF_B200_EIG_AUTO = 0,     /* Jacobi for n <= 64, tridiagonal + divide&conquer above */
  GSCF_B200_EIG_JACOBI = 1,   /* batched cyclic Jacobi in shared memory (n <= 128) */
  GSCF_B200_EIG_TRIDIAG_DC = 2/* Householder tridiagonalisation + D&C + back-transformation */
} gscf_b200_eig_method;

const char* gscf_b200_last_error(void);
const char* gscf_b200_version_string(void);
/* Number of CUDA kernels launched by this library since process start (bench.py: gpu_launches). */
uint64_t gscf_b200_launch_count(void);
/* How many of those were GEMM launches whose k-contiguous operand(s) went through tensor-map TMA
 * (cp.async.bulk.tensor); tests use it to prove that path ran.  GSCF_B200_GEMM_TMA=1 keeps it at zero. */
uint64_t gscf_b200_tensor_map_launch_count(void);

/* ======================================================================================
 * 1. Building blocks on raw device pointers (stream = cudaStream_t passed as void*)
 * ====================================================================================== */

/* C = alpha * op(A) * op(B) + beta * C, FP64 tensor-core (DMMA) tiles.  transa/transb: 0 = N,
 * 1 = T.  Strided batches: batch >= 1 with element strides.  Replaces the lazyten matrix
 * products at pulay_error.cc:36-43, FockMatrix.cc:202, TruncatedOptDampScf.hh:332-343. */
int gscf_b200_dgemm(int transa, int transb, int m, int n, int k, double alpha, const double* A,
                    int lda, long long strideA, const double* B, int ldb, long long strideB,
                    double beta, double* C, int ldc, long long strideC, int batch, void* stream);

/* Symmetric eigensolver, all pairs ascending: A (n x n, lower triangle read, destroyed) ->
 * w[n], Z (n x n, orthonormal columns).  batch matrices with the given strides.
 * Like dsyev, a matrix whose max-norm is outside [1e-100, 1e100] is scaled internally;
 * info[q] != 0 reports non-finite input or a failed reduction of matrix q.
 * work: device workspace of gscf_b200_syev_worksize(n, batch, method) bytes. */
size_t gscf_b200_syev_worksize(int n, int batch, int method);
int gscf_b200_syev(int method, int n, double* A, int lda, long long strideA, double* w,
                   long long stride_w, double* Z, int ldz, long long strideZ, int batch,
                   void* work, size_t work_bytes, int* info_dev, void* stream);

/* Partial spectrum (ScfBaseKeys::n_eigenpairs, ScfBase.hh:87,315-316): all n eigenvalues ascending in w, eigenvectors
 * only of the n_eigenpairs lowest in the first n_eigenpairs columns of Z (the other columns are unspecified).  The
 * tridiagonal path skips the eigenvector part of the top-level divide & conquer merge and the back-transformation of
 * the unwanted vectors; the Jacobi path (n <= 64) computes everything.  Same workspace as gscf_b200_syev. */
int gscf_b200_syevx(int method, int n, int n_eigenpairs, double* A, int lda, long long strideA, double* w,
                    long long stride_w, double* Z, int ldz, long long strideZ, int batch, void* work,
                    size_t work_bytes, int* info_dev, void* stream);

/* out[j] = <E_new, E_j>_F for j < count (E_j = ring + slots[j] * slot_stride), the multi-dot of
 * PulayDiisScf.hh:440-463 reading E_new once.  rows x cols elements per matrix, leading dim ld.
 * batch problems: matrix p at + p * problem_stride, results out[p * out_stride + j]. */
int gscf_b200_frob_multidot(const double* e_new, const double* ring, long long slot_stride,
                            const int* slots_host, int count, int rows, int cols, int ld,
                            long long problem_stride_new, long long problem_stride_ring, int batch,
                            double* out_dev, int out_stride, void* work, size_t work_bytes,
                            void* stream);
size_t gscf_b200_frob_multidot_worksize(int count, int rows, int cols, int batch);

/* out = sum_{i<count} coeff[i] * X_i  (X_i = ring + slots[i]*slot_stride): the materialised
 * OperatorLinearCombination (OperatorLinearCombination.hh:26-47).  coeff_dev: device, per
 * problem coeff_dev[p * coeff_stride + i]. */
int gscf_b200_axpby_n(double* out, long long problem_stride_out, const double* ring,
                      long long slot_stride, long long problem_stride_ring, const int* slots_host,
                      int count, const double* coeff_dev, int coeff_stride, int rows, int cols,
                      int ld, int batch, void* stream);

/* DIIS coefficients (PulayDiisScf.hh:326-393) from the device-resident Gram table of error
 * overlaps: gram[p][a*gram_ld + b] = <e_slot_a, e_slot_b>; order_host lists the k valid slots
 * oldest -> newest.  Builds the bordered (k+1)^2 system, solves it by LU with partial pivoting
 * (dgesv semantics, no regularisation), writes c[p][0..k).  fail_dev[p] != 0 on a singular
 * system (ExcDiisStepFailed). */
int gscf_b200_diis_solve(const double* gram_dev, int gram_ld, long long gram_stride,
                         const int* order_host, int k, double* coeff_dev, int coeff_stride,
                         int* fail_dev, int batch, void* stream);

/* D = C_o C_o^T (FockMatrix.cc:198-203), C n x n_occ (first n_occ columns of the coefficient
 * matrix). */
int gscf_b200_density(int n, int n_occ, const double* C, int ldc, long long strideC, double* D,
                      int ldd, long long strideD, int batch, void* stream);

/* Pulay error e = fac*[(S C_o)(F C_o)^T - (F C_o)(S C_o)^T] (pulay_error.cc:25-44, fac = 2
 * closed shell): two skinny products that write [S C_o | F C_o] and [F C_o | -S C_o], then ONE
 * rank-2*n_occ product over the lower tiles whose epilogue stores the antisymmetric mirror image
 * (6 n^2 n_occ flops instead of the reference's 8).  work: gscf_b200_pulay_error_worksize bytes
 * (padded_ld(n) x 4*n_occ doubles per batch entry, stride_work doubles apart). */
int gscf_b200_pulay_error(int n, int n_occ, double fac, const double* S, int lds,
                          long long strideS, const double* F, int ldf, long long strideF,
                          const double* C, int ldc, long long strideC, double* E, int lde,
                          long long strideE, double* work, long long stride_work, int batch,
                          void* stream);
size_t gscf_b200_pulay_error_worksize(int n, int n_occ, int batch);

/* Copy a rows x cols fp64 column-major matrix between host / device buffers (memloc flags). */
int gscf_b200_copy_matrix(void* dst, int ld_dst, int dst_loc, const void* src, int ld_src, int src_loc,
                          int rows, int cols);

/* compute_damping_coeff of TruncatedOptDampScf.hh:370-394 (host, exact control flow). */
double gscf_b200_compute_damping_coeff(double oda_s, double oda_c, double min_fock_prefactor);

/* ======================================================================================
 * 2. Fock-builder plugin interface (ScfProblemMatrix_i / FocklikeMatrix_i)
 * ====================================================================================== */

/* Device plugin: build F(C) for the listed problems.  All pointers are device pointers
 * except active_host.  Problem p, block b (0 = alpha, 1 = beta):
 *   C  (n x n, ld) at C + p*problem_stride + b*block_stride, columns = orbitals,
 *   D = C_occ C_occ^T at the same offsets in D (built by the SCF path, FockMatrix.cc:202),
 *   F to be written at F_out + p*F_problem_stride + b*block_stride (a history slot chosen by
 *   the solver that no retained matrix aliases - the copy-on-write of ScfBase.hh:343-361),
 *   energies_out[p*2+{0,1}] = {E_1e, E_2e} (FocklikeMatrix_i.hh:69,78). */
typedef struct gscf_b200_fock_args {
  int n_bas, n_blocks, n_alpha, n_beta;
  int n_active;
  const int* active_dev;   /* problem indices, device */
  const int* active_host;  /* same, host */
  const double* C;
  const double* D;
  double* F_out;
  int ld;
  long long block_stride;
  long long problem_stride;
  long long F_problem_stride;
  double* energies_out;
  void* stream;
} gscf_b200_fock_args;
typedef int (*gscf_b200_fock_device_fn)(void* ctx, const gscf_b200_fock_args* args);

/* Host plugin (legacy CPU problem matrices, e.g. gscfmock::FockMatrix): one problem at a
 * time; C, F are host column-major n x n per block with ld = n. */
typedef int (*gscf_b200_fock_host_fn)(void* ctx, int problem, int n_bas, int n_blocks,
                                      const double* C_host, const double* orben_host,
                                      double* F_host, double* e1, double* e2);

/* Optional host error callback replacing the built-in Pulay error (ScfBase.hh:217
 * calculate_error override): fills E_host (n_blocks * n * n). */
typedef int (*gscf_b200_error_host_fn)(void* ctx, int problem, int n_bas, int n_blocks,
                                       const double* F_host, const double* C_host,
                                       const double* orben_host, double* E_host);

/* Iteration hooks (ScfBase.hh:195,200; PulayDiisScf.hh:202; lazyten before/after step). */
typedef enum gscf_b200_event {
  GSCF_B200_EV_BEFORE_ITERATION = 0,
  GSCF_B200_EV_NEW_DIAGMAT = 1,
  GSCF_B200_EV_UPDATE_EIGENPAIRS = 2,
  GSCF_B200_EV_UPDATE_PROBLEM_MATRIX = 3,
  GSCF_B200_EV_AFTER_ITERATION = 4
} gscf_b200_event;
typedef int (*gscf_b200_hook_fn)(void* ctx, int event, int n_iter);

/* ======================================================================================
 * 3. The SCF engine: device-resident state for P independent problems in lock step
 * ====================================================================================== */
typedef struct gscf_b200_engine gscf_b200_engine;

typedef struct gscf_b200_config {
  int n_bas;
  int n_problems;    /* P >= 1; P > 1 = batched ensemble in lock-step (all three solvers);
                        P * n_blocks <= 32768 per engine */
  int n_blocks;      /* 1 restricted, 2 unrestricted (alpha/beta blocks) */
  int n_alpha;
  int n_beta;
  int max_history;   /* ring capacity reserved (>= diis n_prev_steps; >= 2 for TODA) */
  int eig_method;    /* gscf_b200_eig_method */
  int shared_overlap;/* 1: one S for all problems, 0: one per problem */
  void* stream;      /* cudaStream_t or NULL for an engine-owned stream */
} gscf_b200_config;

/* krims::GenMap control parameters as a POD (keys: max_iter, n_eigenpairs, max_error_norm,
 * diis_n_prev_steps, toda_n_prev_steps, toda_min_fock_prefactor). */
typedef struct gscf_b200_params {
  int max_iter;                /* 100 */
  long long n_eigenpairs;      /* -1 = all; k < n: the k lowest pairs (k >= occupied orbitals) */
  double max_error_norm;       /* 5e-7 */
  int diis_n_prev_steps;       /* 4 */
  int toda_n_prev_steps;       /* 2 (must be 2, TruncatedOptDampScf.hh:243) */
  double toda_min_fock_prefactor; /* 0.05 */
} gscf_b200_params;
void gscf_b200_default_params(gscf_b200_params* p);

int gscf_b200_engine_create(const gscf_b200_config* cfg, gscf_b200_engine** out);
void gscf_b200_engine_destroy(gscf_b200_engine* e);
void* gscf_b200_engine_stream(gscf_b200_engine* e);

/* Overlap (ScfStateBase.hh:175): n x n (ld) per problem (or one if shared_overlap); the
 * orthogonaliser is factorised once here: Cholesky S = L L^T, X = L^-1 by default, the Loewdin
 * X = S^-1/2 with the environment variable GSCF_B200_ORTHO=loewdin. */
int gscf_b200_set_overlap(gscf_b200_engine* e, const double* S, int ld, int memloc);
int gscf_b200_set_fock_builder_device(gscf_b200_engine* e, gscf_b200_fock_device_fn fn, void* ctx);
int gscf_b200_set_fock_builder_host(gscf_b200_engine* e, gscf_b200_fock_host_fn fn, void* ctx);
int gscf_b200_set_error_host(gscf_b200_engine* e, gscf_b200_error_host_fn fn, void* ctx);
int gscf_b200_set_hook(gscf_b200_engine* e, gscf_b200_hook_fn fn, void* ctx);

/* Initial problem matrix (the probmat handed to ScfBase::solve, ScfBase.hh:140-145) with its
 * energies: F (P * n_blocks matrices, n x n, ld, contiguous stride ld*n), e1/e2 per problem. */
int gscf_b200_set_problem_matrix(gscf_b200_engine* e, const double* F, int ld, int memloc,
                                 const double* e1_host, const double* e2_host);
/* solve_with_guess / obtain_guess_from (ScfStateBase.hh:291-303): stores the guess
 * coefficients (and orbital energies, may be NULL) and calls the Fock builder once. */
int gscf_b200_set_guess(gscf_b200_engine* e, const double* C, int ld, int memloc,
                        const double* orben_host);

/* Guess = eigenvectors of G against S (core guess G = H; one matrix per problem, used for
 * every block), installed like obtain_guess_from: the Fock builder is called once. */
int gscf_b200_set_guess_from_matrix(gscf_b200_engine* e, const double* G, int ld, int memloc);

/* Whole-loop, device-resident drivers (PlainScf.hh:121-147, PulayDiisScf.hh:283-324,
 * TruncatedOptDampScf.hh:239-289).  Return GSCF_B200_OK when every problem converged;
 * GSCF_B200_ERR_MAX_ITER etc. otherwise (per-problem flags via the getters).  All three run the
 * engine's P problems in lock-step; a problem leaves the active set when it converges or fails
 * (TODA: damping coefficient, keep/flip decision and previous matrix are per problem). */
int gscf_b200_solve_plain(gscf_b200_engine* e, const gscf_b200_params* p);
int gscf_b200_solve_diis(gscf_b200_engine* e, const gscf_b200_params* p);
int gscf_b200_solve_toda(gscf_b200_engine* e, const gscf_b200_params* p);

/* Results / state access (host output buffers). */
int gscf_b200_get_n_iter(gscf_b200_engine* e, int* n_iter);            /* [P] */
int gscf_b200_get_converged(gscf_b200_engine* e, int* converged);      /* [P] */
int gscf_b200_get_status(gscf_b200_engine* e, int* status);            /* [P] gscf_b200_status */
int gscf_b200_get_error_norm(gscf_b200_engine* e, double* err);        /* [P] */
int gscf_b200_get_energies(gscf_b200_engine* e, double* e1, double* e2);/* [P] each */
int gscf_b200_get_n_mtx_applies(gscf_b200_engine* e, long long* n);    /* [P] */
int gscf_b200_get_orben(gscf_b200_engine* e, double* orben);           /* [P][n_blocks][n] */
/* Eigenpairs kept by the last solve (ScfBaseKeys::n_eigenpairs, ScfBase.hh:87): the first k entries of every orben
 * row and the first k columns of every coefficient block are the eigensolution; the rest is unspecified. */
int gscf_b200_get_n_eigenpairs(gscf_b200_engine* e);
int gscf_b200_get_coeff(gscf_b200_engine* e, double* C, int ld);       /* [P][n_blocks] n x n */
int gscf_b200_get_prev_coeff(gscf_b200_engine* e, double* C, int ld);
int gscf_b200_get_fock(gscf_b200_engine* e, double* F, int ld);
int gscf_b200_get_diagmat(gscf_b200_engine* e, double* F, int ld);
int gscf_b200_get_density(gscf_b200_engine* e, double* D, int ld);
int gscf_b200_get_error(gscf_b200_engine* e, double* E, int ld);
/* DIIS state mirrors (PulayDiisScfState :56-86): coefficients oldest->newest of the last
 * extrapolation, [P][max_history], count returned through n_coeff. */
int gscf_b200_get_diis_coefficients(gscf_b200_engine* e, double* c, int* n_coeff);
/* Coefficients of the extrapolation being diagonalised right now (valid inside the NEW_DIAGMAT /
 * UPDATE_EIGENPAIRS hooks), [P][k] oldest -> newest. */
int gscf_b200_get_current_diis_coefficients(gscf_b200_engine* e, double* c, int k);
/* Lazy materialisation of the DIIS rings (PulayDiisScfState :56-81) after a solve - a host that overrides no hook
 * never pays a per-iteration device-to-host copy: number of entries in the rings, then entry i (oldest = 0) of the
 * error ring / of the problem-matrix ring, [P][n_blocks] n x n each. */
int gscf_b200_get_history_size(gscf_b200_engine* e);
int gscf_b200_get_error_entry(gscf_b200_engine* e, int entry, double* E, int ld);
int gscf_b200_get_fock_entry(gscf_b200_engine* e, int entry, double* F, int ld);
/* overlap vector of ring entry i (oldest = 0), newest-first like PulayDiisScf.hh:67-70 */
int gscf_b200_get_error_overlaps(gscf_b200_engine* e, int problem, int entry, double* ov, int* len);
int gscf_b200_get_damping_coeff(gscf_b200_engine* e, double* damp);    /* TODA, [P] */
/* Per-iteration records of problem 0..P-1: energy, error norm; rows = iterations. */
int gscf_b200_get_trace(gscf_b200_engine* e, int problem, double* energy, double* error_norm,
                        double* coeff /* max_history per iteration, NaN padded */, int max_rows,
                        int* n_rows);
/* CUDA-event timings of the last solve: milliseconds in the SCF path (excluding the Fock
 * builder) and inside the Fock builder. */
int gscf_b200_get_timings(gscf_b200_engine* e, double* scf_ms, double* fock_ms);
/* Per-phase device time of the last solve, ms: [0] extrapolate/diagmat, [1] eigensolve,
 * [2] density, [3] error + overlaps, [4] coefficient solve, [5] Fock builder, [6] whole loop,
 * [7] orthogonaliser set-up, [8..10] tridiagonalisation / D&C / back-transformation,
 * [14] Fock builder inside the last gscf_b200_set_guess. */
int gscf_b200_get_phase_timings(gscf_b200_engine* e, double* ms, int n);
int gscf_b200_set_phase_timing(gscf_b200_engine* e, int enabled);

/* ======================================================================================
 * 3b. Sharded ensembles: one process per GPU, the ensemble split by problem index,
 *     no data-path collective; the ONLY collective is this all-gather of per-problem records
 *     (NCCL over NVLink, bound at run time with dlopen("libnccl.so.2") - no link-time
 *     dependency).  The reference has no counterpart.
 *     Host protocol: rank 0 calls gscf_b200_comm_unique_id and distributes the 128 bytes by its own
 *     means (MPI_Bcast, a file, torch.distributed ...); every rank calls gscf_b200_comm_create with
 *     the CUDA device of its engine current; after the solve every rank calls
 *     gscf_b200_ensemble_gather with counts[r] = problems owned by rank r.
 * ====================================================================================== */
typedef struct gscf_b200_comm gscf_b200_comm;
typedef struct gscf_b200_record {
  double energy_total;   /* E_1e + E_2e of the final problem matrix */
  double error_norm;     /* last_error_norm */
  int n_iter;
  int converged;
  int status;            /* gscf_b200_status of the problem */
  int reserved;
} gscf_b200_record;
int gscf_b200_nccl_version(void);  /* 0 when no NCCL library can be loaded */
int gscf_b200_comm_unique_id(void* id128);
int gscf_b200_comm_create(const void* id128, int world, int rank, gscf_b200_comm** out);
void gscf_b200_comm_destroy(gscf_b200_comm* c);
/* out: sum(counts) records on the host, ordered by rank then local problem index - identical on
 * every rank.  Runs on the engine's stream; returns after the records are on the host. */
int gscf_b200_ensemble_gather(gscf_b200_engine* e, gscf_b200_comm* c, const int* counts,
                              gscf_b200_record* out);

/* ======================================================================================
 * 4. Synthetic density-fitted Fock builder (device plugin used by bench.py and the tests;
 *    generator of SURVEY.md 8(d); not part of gscf, timed separately)
 * ====================================================================================== */
typedef struct gscf_b200_dfock gscf_b200_dfock;
int gscf_b200_dfock_create(int n_bas, int n_alpha, int n_beta, int n_blocks, int n_aux,
                           int n_problems, const long long* seeds_host, double g,
                           gscf_b200_dfock** out);
void gscf_b200_dfock_destroy(gscf_b200_dfock* f);
/* The device callback + its ctx (= the dfock handle) to register with the engine. */
gscf_b200_fock_device_fn gscf_b200_dfock_callback(void);
/* Device pointers of the generated S, H (P matrices n x n, ld = gscf_b200_padded_ld(n)). */
const double* gscf_b200_dfock_overlap(gscf_b200_dfock* f);
const double* gscf_b200_dfock_core(gscf_b200_dfock* f);
int gscf_b200_padded_ld(int n);

/* Micro-benchmarks used for the roofline denominators (profiles/): FP64 DMMA and DFMA
 * register-resident loops, and a streaming copy.  Return TFLOP/s or GB/s. */
int gscf_b200_measure_peaks(double* dmma_tflops, double* dfma_tflops, double* copy_gbs);

#ifdef __cplusplus
}
#endif
#endif /* GSCF_B200_H */
