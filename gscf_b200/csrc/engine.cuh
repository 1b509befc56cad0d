// The SCF engine: device-resident state of P independent SCF problems driven in lock step.
// Host control flow restates gscf's solve_state loops (PlainScf.hh:121-147,
// PulayDiisScf.hh:283-324, TruncatedOptDampScf.hh:239-289); all arithmetic runs in the
// sm_100a kernels of this library.
#pragma once
#include <vector>

#include "common.cuh"
#include "stream.cuh"

struct gscf_b200_engine {
  gscf_b200_config cfg{};
  int n = 0, ld = 0, P = 0, nblk = 1, M = 0;  // M = ring capacity; slot M = initial matrix
  int occ[2] = {0, 0};
  int omax = 0;
  long long mstride = 0, pstride = 0;  // elements per matrix / per problem
  int eig_method = 0;
  int ortho = 1;             // orthogonaliser: 0 Loewdin X = S^-1/2 (symmetric), 1 Cholesky X = L^-1 (lower triangular)
  double* cholwork = nullptr;
  int neig = 0;  // eigenpairs kept per block (n_eigenpairs, ScfBase.hh:87); 0 = not yet set = all
  cudaStream_t stream = nullptr;
  bool owns_stream = false;

  // ---- device state ------------------------------------------------------------------------
  double* S = nullptr;      // [P*nblk] (or 1 if shared) overlap
  double* X = nullptr;      // same shape: the orthogonaliser (L^-1, or S^-1/2 with GSCF_B200_ORTHO=loewdin)
  long long sx_stride = 0;  // 0 when shared
  double* Fring = nullptr;  // [P][M+1][nblk] problem matrices (slot M: initial matrix)
  double* Ering = nullptr;  // [P][M][nblk] error matrices
  double* Fdiag = nullptr;  // [P][nblk] matrix to diagonalise when it is a combination
  double* Cbuf = nullptr;   // [2][P][nblk] coefficients (current / previous ping-pong)
  double* orben = nullptr;  // [2][P][nblk][n]
  double* D = nullptr;      // [P][nblk] densities
  double* Aeig = nullptr;   // [P][nblk] eigensolver input (X^T F X)
  double* Zeig = nullptr;   // [P][nblk] eigenvectors in the orthogonal basis
  double* Tmp = nullptr;    // [P][nblk] GEMM temporary
  double* errwork = nullptr;  // [P][nblk][ld x 2*omax]
  double* gram = nullptr;   // [P][M*M]
  double* coeff = nullptr;  // [P][M]
  double* dots = nullptr;   // [P][kMaxHistory]
  double* energies = nullptr;  // [P][2]
  double* scal = nullptr;   // [P][4] scratch scalars (TODA trace ...)
  int* fail = nullptr;      // [P] DIIS failure flags
  int* eiginfo = nullptr;   // [P*nblk]
  void* eigwork = nullptr;
  size_t eigwork_bytes = 0;
  double* mdwork = nullptr;
  int* act_p = nullptr;     // active problems
  int* keep_p = nullptr;    // TODA: problems that keep their matrix this iteration
  int* act_m = nullptr;     // active matrices (p*nblk+b)
  int* act_mb[2] = {nullptr, nullptr};
  // pinned host staging
  double* h_stage = nullptr;  // [P * (2 kMaxHistory + 2)]: dots | energies | coefficients
  int* h_istage = nullptr;    // [P * (1 + nblk)]: DIIS failure flags | eigensolver info

  // ---- host mirrors --------------------------------------------------------------------------
  int cur = 0;  // index of the current coefficient buffer in Cbuf/orben
  int cur_slot = -1;  // Fring slot holding the current problem matrix (-1: none set)
  bool have_overlap = false, have_coeff = false;
  std::vector<int> n_iter, converged, status, active;
  std::vector<long long> n_mtx_applies;
  std::vector<double> errnorm, e1, e2, damp;
  std::vector<int> hist_order;  // ring slots oldest -> newest
  int n_coeff = 0;              // number of coefficients of the last extrapolation
  int diagmat_kind = 0;         // 0: Fring[cur_slot], 1: Fring slot diagmat_slot, 2: Fdiag buffer
  int diagmat_slot = 0;
  int last_error_slot = -1;
  std::vector<std::vector<double>> trace_energy, trace_error;
  std::vector<std::vector<double>> trace_coeff;  // rows of kMaxHistory
  // callbacks
  gscf_b200_fock_device_fn fock_dev = nullptr;
  void* fock_dev_ctx = nullptr;
  gscf_b200_fock_host_fn fock_host = nullptr;
  void* fock_host_ctx = nullptr;
  gscf_b200_error_host_fn error_host = nullptr;
  void* error_host_ctx = nullptr;
  gscf_b200_hook_fn hook = nullptr;
  void* hook_ctx = nullptr;
  // timings
  double scf_ms = 0.0, fock_ms = 0.0, overlap_ms = 0.0;
  double guess_fock_ms = 0.0;  // Fock-builder time inside the last set_guess
  bool phase_timing = false;
  double phase_ms[16] = {0};
  long long phase_count[16] = {0};
  cudaEvent_t phase_open[16] = {nullptr};
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev_spans;  // (phase, (a,b))

  double* Fslot(int slot) const { return Fring + (long long)slot * pstride; }
  double* Eslot(int slot) const { return Ering + (long long)slot * pstride; }
  long long fring_pstride() const { return (long long)(M + 1) * pstride; }
  long long ering_pstride() const { return (long long)M * pstride; }
  double* Ccur() const { return Cbuf + (long long)cur * P * pstride; }
  double* Cprev() const { return Cbuf + (long long)(1 - cur) * P * pstride; }
  double* orben_cur() const { return orben + (long long)cur * P * nblk * n; }
  double* orben_prev() const { return orben + (long long)(1 - cur) * P * nblk * n; }
};
