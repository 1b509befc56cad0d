// SCF engine implementation.  See engine.cuh / include/gscf_b200.h.
#include "engine.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "chol.cuh"
#include "eig.cuh"
#include "gemm.cuh"
#include "smallsolve.hpp"

using namespace gscfb;
typedef gscf_b200_engine Engine;

namespace {

enum Phase { PH_DIAGMAT = 0, PH_EIG = 1, PH_DENSITY = 2, PH_ERROR = 3, PH_COEFF = 4, PH_FOCK = 5,
             PH_TOTAL = 6 };

cudaEvent_t take_event(Engine* e) {
  if (e->ev_used == e->ev_pool.size()) {
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    e->ev_pool.push_back(ev);
  }
  return e->ev_pool[e->ev_used++];
}

struct Span {
  Engine* e;
  int phase;
  cudaEvent_t a = nullptr;
  bool on;
  Span(Engine* e_, int phase_, bool force = false) : e(e_), phase(phase_) {
    on = force || e->phase_timing;
    if (on) {
      a = take_event(e);
      cudaEventRecord(a, e->stream);
    }
  }
  ~Span() {
    if (on) {
      cudaEvent_t b = take_event(e);
      cudaEventRecord(b, e->stream);
      e->ev_spans.push_back({phase, {a, b}});
    }
  }
};

void collect_timings(Engine* e) {
  for (int i = 0; i < 16; ++i) { e->phase_ms[i] = 0.0; e->phase_count[i] = 0; }
  for (auto& s : e->ev_spans) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.second.first, s.second.second) == cudaSuccess) {
      e->phase_ms[s.first] += ms;
      e->phase_count[s.first] += 1;
    }
  }
  e->ev_spans.clear();
  e->ev_used = 0;
  e->fock_ms = e->phase_ms[PH_FOCK];
  e->scf_ms = e->phase_ms[PH_TOTAL] - e->fock_ms;
  e->phase_ms[7] = e->overlap_ms;
  e->phase_ms[14] = e->guess_fock_ms;
}

thread_local Engine* g_cur_engine = nullptr;
struct EngineScope {
  explicit EngineScope(Engine* e) { g_cur_engine = e; }
  ~EngineScope() { g_cur_engine = nullptr; }
};

template <typename T>
int dev_alloc(T** p, size_t count) {
  *p = nullptr;
  if (count == 0) return GSCF_B200_OK;
  GSCF_CUDA_TRY(cudaMalloc((void**)p, count * sizeof(T)));
  GSCF_CUDA_TRY(cudaMemset(*p, 0, count * sizeof(T)));
  return GSCF_B200_OK;
}

__global__ void gram_update_kernel(double* __restrict__ gram, int M, const double* __restrict__ dots,
                                   SlotList slots, int head, const int* __restrict__ act, int nact) {
  const int z = blockIdx.x * blockDim.x + threadIdx.x;
  if (z >= nact) return;
  const long long p = act[z];
  double* g = gram + p * M * M;
  for (int j = 0; j < slots.n; ++j) {
    const double v = dots[p * kMaxHistory + j];
    g[head * M + slots.slot[j]] = v;
    g[slots.slot[j] * M + head] = v;
  }
}

// Upload the active lists.
int upload_active(Engine* e) {
  std::vector<int> ap, am, ab[2];
  for (int p = 0; p < e->P; ++p)
    if (e->active[p]) {
      ap.push_back(p);
      for (int b = 0; b < e->nblk; ++b) {
        am.push_back(p * e->nblk + b);
        ab[b].push_back(p * e->nblk + b);
      }
    }
  if (ap.empty()) return GSCF_B200_OK;
  GSCF_CUDA_TRY(cudaMemcpyAsync(e->act_p, ap.data(), ap.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  GSCF_CUDA_TRY(cudaMemcpyAsync(e->act_m, am.data(), am.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  for (int b = 0; b < e->nblk; ++b)
    GSCF_CUDA_TRY(cudaMemcpyAsync(e->act_mb[b], ab[b].data(), ab[b].size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  // the host vectors go out of scope: make the copies complete first
  GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
  return GSCF_B200_OK;
}

int n_active(const Engine* e) {
  int c = 0;
  for (int p = 0; p < e->P; ++p) c += e->active[p] ? 1 : 0;
  return c;
}

// ---- host <-> device matrix transfers ------------------------------------------------------------
// `count` groups of `inner` consecutive matrices in ONE strided copy (an ensemble is thousands of small matrices: one
// cudaMemcpy2DAsync each was 4096 x 3 launches per solve at C4).  Device side: group i at dev + i * dev_gstride, the
// matrices of a group mstride = ld * n apart (so a group is inner * n rows of pitch ld); other side (host or device):
// matrices ld_o * n apart, groups o_gstride elements apart.  Both group strides are whole numbers of rows.
int copy_matrices(Engine* e, double* dev, long long dev_gstride, double* other, int ld_o, long long o_gstride,
                  int inner, int count, cudaMemcpyKind kind, bool to_device) {
  if (count <= 0 || inner <= 0) return GSCF_B200_OK;
  const size_t rows = (size_t)inner * e->n;
  if (count > 1 && (dev_gstride % e->ld != 0 || o_gstride % ld_o != 0 || (size_t)(dev_gstride / e->ld) < rows ||
                    (size_t)(o_gstride / ld_o) < rows)) {
    for (int i = 0; i < count; ++i)
      GSCF_TRY(copy_matrices(e, dev + i * dev_gstride, dev_gstride, other + i * o_gstride, ld_o, o_gstride, inner, 1,
                             kind, to_device));
    return GSCF_B200_OK;
  }
  cudaMemcpy3DParms p = {};
  const cudaPitchedPtr dptr = make_cudaPitchedPtr(dev, (size_t)e->ld * sizeof(double), (size_t)e->n * sizeof(double),
                                                  count > 1 ? (size_t)(dev_gstride / e->ld) : rows);
  const cudaPitchedPtr optr = make_cudaPitchedPtr(other, (size_t)ld_o * sizeof(double), (size_t)e->n * sizeof(double),
                                                  count > 1 ? (size_t)(o_gstride / ld_o) : rows);
  p.srcPtr = to_device ? optr : dptr;
  p.dstPtr = to_device ? dptr : optr;
  p.extent = make_cudaExtent((size_t)e->n * sizeof(double), rows, (size_t)count);
  p.kind = kind;
  GSCF_CUDA_TRY(cudaMemcpy3DAsync(&p, e->stream));
  return GSCF_B200_OK;
}

// `count` matrices, dst_stride apart on the device, contiguous (ld_src * n apart) at the source
int copy_in(Engine* e, double* dst_dev, long long dst_stride, const double* src, int ld_src,
            int memloc, int count) {
  const cudaMemcpyKind kind = memloc == GSCF_B200_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  return copy_matrices(e, dst_dev, dst_stride, const_cast<double*>(src), ld_src, (long long)ld_src * e->n, 1, count, kind,
                       true);
}

// P groups of n_blocks matrices each (groups group_stride apart on the device; blocks mstride apart) from a source
// that holds them contiguously
int copy_in_groups(Engine* e, double* dst_dev, long long group_stride, const double* src, int ld_src, int memloc) {
  const cudaMemcpyKind kind = memloc == GSCF_B200_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  return copy_matrices(e, dst_dev, group_stride, const_cast<double*>(src), ld_src, (long long)e->nblk * ld_src * e->n,
                       e->nblk, e->P, kind, true);
}

int copy_out(Engine* e, double* dst_host, int ld_dst, const double* src_dev, long long src_pstride,
             long long src_bstride) {
  if (src_bstride == e->mstride) {
    GSCF_TRY(copy_matrices(e, const_cast<double*>(src_dev), src_pstride, dst_host, ld_dst,
                           (long long)e->nblk * ld_dst * e->n, e->nblk, e->P, cudaMemcpyDeviceToHost, false));
  } else {
    for (int p = 0; p < e->P; ++p)
      for (int b = 0; b < e->nblk; ++b)
        GSCF_CUDA_TRY(cudaMemcpy2DAsync(dst_host + ((size_t)p * e->nblk + b) * ld_dst * e->n,
                                        (size_t)ld_dst * sizeof(double),
                                        src_dev + p * src_pstride + b * src_bstride,
                                        (size_t)e->ld * sizeof(double), (size_t)e->n * sizeof(double),
                                        e->n, cudaMemcpyDeviceToHost, e->stream));
  }
  GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
  return GSCF_B200_OK;
}

int call_hook(Engine* e, int ev, int it) {
  if (!e->hook) return GSCF_B200_OK;
  GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
  if (e->hook(e->hook_ctx, ev, it) != 0) {
    set_last_error("hook callback returned non-zero");
    return GSCF_B200_ERR_CALLBACK;
  }
  return GSCF_B200_OK;
}

// ---- path steps ------------------------------------------------------------------------------

// Generalised eigensolve of the matrices at Fd (+p*fd_pstride + b*mstride) for the active set (ScfBase.hh:291-339):
//   default (Cholesky, S = L L^T, X = L^-1 lower triangular):  A' = X Fd X^T,  A' Z = Z w,  C = X^T Z
//   GSCF_B200_ORTHO=loewdin (X = S^-1/2 symmetric):            A' = X Fd X,    A' Z = Z w,  C = X Z
int step_eigensolve(Engine* e, const double* Fd, long long fd_pstride, int nact) {
  Span sp(e, PH_EIG);
  const int n = e->n, ld = e->ld;
  const int nm = nact * e->nblk;
  // push_new_eigensolution: current -> previous (ScfStateBase.hh:280-289).  A copy restricted to
  // the active problems (not a pointer swap): converged ensemble members keep their final C.
  if (e->have_coeff) {
    GSCF_TRY(scale_copy(e->Cprev(), ld, e->pstride, e->Ccur(), ld, e->pstride, 1.0, n, n * e->nblk, nact,
                        e->act_p, e->stream));
    GSCF_TRY(scale_copy(e->orben_prev(), n * e->nblk, (long long)n * e->nblk, e->orben_cur(), n * e->nblk,
                        (long long)n * e->nblk, 1.0, n * e->nblk, 1, nact, e->act_p, e->stream));
  }
  // matrix-level strides: Fd may live in the ring (problem stride != nblk*mstride); handle per block
  const bool tri = e->ortho == 1;  // X = L^-1 lower triangular:  A' = X Fd X^T  (else X symmetric: A' = X Fd X)
  for (int b = 0; b < e->nblk; ++b) {
    // T = Fd * X^T
    GemmParams g = gemm_params(n, n, n, 1.0, Fd + b * e->mstride, ld, e->X + (e->sx_stride ? b * e->mstride : 0), ld,
                               0.0, e->Tmp + b * e->mstride, ld);
    g.sA = fd_pstride; g.sB = e->sx_stride ? e->pstride : 0; g.sC = e->pstride; g.batch_map = e->act_p; g.entries = e->P;
    g.ktri = tri ? 1 : 0;  // X^T(k, j) = X(j, k) = 0 for k > j
    GSCF_TRY(launch_gemm(false, tri, g, nact, n, n, e->stream));
    // A' = X * T
    GemmParams h = gemm_params(n, n, n, 1.0, e->X + (e->sx_stride ? b * e->mstride : 0), ld, e->Tmp + b * e->mstride, ld,
                               0.0, e->Aeig + b * e->mstride, ld);
    h.sA = e->sx_stride ? e->pstride : 0; h.sB = e->pstride; h.sC = e->pstride; h.batch_map = e->act_p; h.entries = e->P;
    h.ktri = tri ? 2 : 0;  // X(i, k) = 0 for k > i
    h.lower_only = tri ? 1 : 0;  // only the tiles on and below the diagonal are computed ...
    h.mirror = tri ? 1 : 0;      // ... and the epilogue stores every lower entry on both sides: A' exactly symmetric
    GSCF_TRY(launch_gemm(false, false, h, nact, n, n, e->stream));
  }
  const int nv = (e->neig > 0 && e->neig < n) ? e->neig : n;
  GSCF_TRY(syev_batched(e->eig_method, n, e->Aeig, ld, e->mstride, e->orben_cur(), n, e->Zeig, ld,
                        e->mstride, nm, e->act_m, e->eigwork, e->eigwork_bytes, e->eiginfo, e->stream, nv));
  for (int b = 0; b < e->nblk; ++b) {
    GemmParams g = gemm_params(n, nv, n, 1.0, e->X + (e->sx_stride ? b * e->mstride : 0), ld, e->Zeig + b * e->mstride, ld,
                               0.0, e->Ccur() + b * e->mstride, ld);
    g.sA = e->sx_stride ? e->pstride : 0; g.sB = e->pstride; g.sC = e->pstride; g.batch_map = e->act_p; g.entries = e->P;
    g.ktri = tri ? 3 : 0;  // C = X^T Z:  X^T(i, k) = X(k, i) = 0 for k < i
    GSCF_TRY(launch_gemm(tri, false, g, nact, n, nv, e->stream));
  }
  e->have_coeff = true;
  return GSCF_B200_OK;
}

// D = C_o C_o^T per block (FockMatrix.cc:198-203)
int step_density(Engine* e, int nact) {
  Span sp(e, PH_DENSITY);
  for (int b = 0; b < e->nblk; ++b) {
    GemmParams g = gemm_params(e->n, e->n, e->occ[b], 1.0, e->Ccur() + b * e->mstride, e->ld,
                               e->Ccur() + b * e->mstride, e->ld, 0.0, e->D + b * e->mstride, e->ld);
    g.sA = g.sB = g.sC = e->pstride; g.batch_map = e->act_p; g.entries = e->P;
    GSCF_TRY(launch_gemm(false, true, g, nact, e->n, e->n, e->stream));
  }
  return GSCF_B200_OK;
}

// Host-plugin adapters (slow path: D2H of C, H2D of F) ------------------------------------------
int fock_via_host(Engine* e, int slot) {
  const int n = e->n;
  std::vector<double> Ch((size_t)e->nblk * n * n), Fh((size_t)e->nblk * n * n), wh((size_t)e->nblk * n);
  for (int p = 0; p < e->P; ++p) {
    if (!e->active[p]) continue;
    for (int b = 0; b < e->nblk; ++b)
      GSCF_CUDA_TRY(cudaMemcpy2DAsync(Ch.data() + (size_t)b * n * n, (size_t)n * sizeof(double),
                                      e->Ccur() + p * e->pstride + b * e->mstride,
                                      (size_t)e->ld * sizeof(double), (size_t)n * sizeof(double), n,
                                      cudaMemcpyDeviceToHost, e->stream));
    GSCF_CUDA_TRY(cudaMemcpyAsync(wh.data(), e->orben_cur() + (size_t)p * e->nblk * n,
                                  (size_t)e->nblk * n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
    double e1 = 0, e2 = 0;
    if (e->fock_host(e->fock_host_ctx, p, n, e->nblk, Ch.data(), wh.data(), Fh.data(), &e1, &e2) != 0) {
      set_last_error("host Fock builder returned non-zero");
      return GSCF_B200_ERR_CALLBACK;
    }
    for (int b = 0; b < e->nblk; ++b)
      GSCF_CUDA_TRY(cudaMemcpy2DAsync(e->Fslot(slot) + p * e->fring_pstride() + b * e->mstride,
                                      (size_t)e->ld * sizeof(double), Fh.data() + (size_t)b * n * n,
                                      (size_t)n * sizeof(double), (size_t)n * sizeof(double), n,
                                      cudaMemcpyHostToDevice, e->stream));
    const double en[2] = {e1, e2};
    GSCF_CUDA_TRY(cudaMemcpyAsync(e->energies + 2 * p, en, 2 * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
  }
  return GSCF_B200_OK;
}

// update_problem_matrix (ScfBase.hh:341-370): F(C) written into ring slot `slot`.
int step_fock(Engine* e, int slot, int nact, const std::vector<int>& act_host) {
  Span sp(e, PH_FOCK, /*force=*/true);
  if (e->fock_dev) {
    gscf_b200_fock_args a{};
    a.n_bas = e->n; a.n_blocks = e->nblk; a.n_alpha = e->occ[0]; a.n_beta = e->occ[1];
    a.n_active = nact; a.active_dev = e->act_p; a.active_host = act_host.data();
    a.C = e->Ccur(); a.D = e->D; a.F_out = e->Fslot(slot); a.ld = e->ld;
    a.block_stride = e->mstride; a.problem_stride = e->pstride; a.F_problem_stride = e->fring_pstride();
    a.energies_out = e->energies; a.stream = e->stream;
    if (e->fock_dev(e->fock_dev_ctx, &a) != 0) {
      if (std::string(last_error_cstr()).empty()) set_last_error("device Fock builder returned non-zero");
      return GSCF_B200_ERR_CALLBACK;
    }
  } else if (e->fock_host) {
    GSCF_TRY(fock_via_host(e, slot));
  } else {
    set_last_error("no Fock builder registered (gscf_b200_set_fock_builder_*)");
    return GSCF_B200_ERR_INVALID_STATE;
  }
  e->cur_slot = slot;
  return GSCF_B200_OK;
}

// calculate_error: built-in Pulay error (pulay_error.cc:25-44) of (F in slot, current C) into Edst.
int step_error(Engine* e, int fslot, double* Edst, long long e_pstride, int nact) {
  const int n = e->n, ld = e->ld;
  if (e->error_host) {
    std::vector<double> Fh((size_t)e->nblk * n * n), Ch((size_t)e->nblk * n * n), Eh((size_t)e->nblk * n * n),
        wh((size_t)e->nblk * n);
    for (int p = 0; p < e->P; ++p) {
      if (!e->active[p]) continue;
      for (int b = 0; b < e->nblk; ++b) {
        GSCF_CUDA_TRY(cudaMemcpy2DAsync(Fh.data() + (size_t)b * n * n, (size_t)n * sizeof(double),
                                        e->Fslot(fslot) + p * e->fring_pstride() + b * e->mstride,
                                        (size_t)ld * sizeof(double), (size_t)n * sizeof(double), n,
                                        cudaMemcpyDeviceToHost, e->stream));
        GSCF_CUDA_TRY(cudaMemcpy2DAsync(Ch.data() + (size_t)b * n * n, (size_t)n * sizeof(double),
                                        e->Ccur() + p * e->pstride + b * e->mstride,
                                        (size_t)ld * sizeof(double), (size_t)n * sizeof(double), n,
                                        cudaMemcpyDeviceToHost, e->stream));
      }
      GSCF_CUDA_TRY(cudaMemcpyAsync(wh.data(), e->orben_cur() + (size_t)p * e->nblk * n,
                                    (size_t)e->nblk * n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
      GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
      if (e->error_host(e->error_host_ctx, p, n, e->nblk, Fh.data(), Ch.data(), wh.data(), Eh.data()) != 0) {
        set_last_error("host error callback returned non-zero");
        return GSCF_B200_ERR_CALLBACK;
      }
      for (int b = 0; b < e->nblk; ++b)
        GSCF_CUDA_TRY(cudaMemcpy2DAsync(Edst + p * e_pstride + b * e->mstride, (size_t)ld * sizeof(double),
                                        Eh.data() + (size_t)b * n * n, (size_t)n * sizeof(double),
                                        (size_t)n * sizeof(double), n, cudaMemcpyHostToDevice, e->stream));
      GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
    }
    return GSCF_B200_OK;
  }
  const double fac = e->nblk == 1 ? 2.0 : 1.0;
  const long long wstride = (long long)pulay_error_work_doubles(n, e->omax);  // per matrix: [SC | FC | FC | -SC]
  for (int b = 0; b < e->nblk; ++b) {
    PulayErrorArgs a{};
    a.n = n; a.o = e->occ[b]; a.fac = fac;
    a.S = e->S + (e->sx_stride ? b * e->mstride : 0); a.lds = ld; a.sS = e->sx_stride ? e->pstride : 0;
    a.F = e->Fslot(fslot) + b * e->mstride; a.ldf = ld; a.sF = e->fring_pstride();
    a.C = e->Ccur() + b * e->mstride; a.ldc = ld; a.sC = e->pstride;
    a.E = Edst + b * e->mstride; a.lde = ld; a.sE = e_pstride;
    a.work = e->errwork + b * wstride; a.swork = e->nblk * wstride;
    a.batch = nact; a.batch_map = e->act_p; a.entries = e->P;
    GSCF_TRY(pulay_error_fused(a, e->stream));
  }
  return GSCF_B200_OK;
}

// <E_new, E_slot_j> for the listed slots into dots[p][j]
int step_multidot(Engine* e, const double* Enew, long long enew_pstride, const SlotList& slots, int nact) {
  return multidot(Enew, enew_pstride, e->Ering, e->pstride, e->ering_pstride(), slots, e->n,
                  e->n * e->nblk, e->ld, nact, e->act_p, e->dots, kMaxHistory, e->mdwork, e->stream);
}

struct IterScalars {
  std::vector<double> dots, energies;
  std::vector<int> fail, eiginfo;
};

// One D2H round trip per iteration: dots, energies, failure flags (and the DIIS coefficients when asked for), through
// pinned staging buffers - into pageable memory every cudaMemcpyAsync is its own blocking transfer (~15 us each, five per
// iteration: 1 % of a C2 iteration).
int fetch_scalars(Engine* e, IterScalars& s, double* coeff_out = nullptr) {
  const size_t nd = (size_t)e->P * kMaxHistory, ne = (size_t)e->P * 2, nf = (size_t)e->P, ni = (size_t)e->P * e->nblk;
  double* hd = e->h_stage;           // [nd] dots | [ne] energies | [nd] coefficients
  int* hi = e->h_istage;             // [nf] fail | [ni] eiginfo
  GSCF_CUDA_TRY(cudaMemcpyAsync(hd, e->dots, nd * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  GSCF_CUDA_TRY(cudaMemcpyAsync(hd + nd, e->energies, ne * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  if (coeff_out)
    GSCF_CUDA_TRY(cudaMemcpyAsync(hd + nd + ne, e->coeff, nd * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  GSCF_CUDA_TRY(cudaMemcpyAsync(hi, e->fail, nf * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  GSCF_CUDA_TRY(cudaMemcpyAsync(hi + nf, e->eiginfo, ni * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
  s.dots.assign(hd, hd + nd);
  s.energies.assign(hd + nd, hd + nd + ne);
  s.fail.assign(hi, hi + nf);
  s.eiginfo.assign(hi + nf, hi + nf + ni);
  if (coeff_out) std::copy(hd + nd + ne, hd + nd + ne + nd, coeff_out);
  return GSCF_B200_OK;
}

void begin_solve(Engine* e) {
  for (int p = 0; p < e->P; ++p) {
    e->n_iter[p] = 0;
    e->converged[p] = 0;
    e->status[p] = GSCF_B200_OK;
    e->active[p] = 1;
    e->n_mtx_applies[p] = 0;
    e->errnorm[p] = std::numeric_limits<double>::quiet_NaN();  // Constants<real>::invalid
    e->damp[p] = 1.0;
    e->trace_energy[p].clear();
    e->trace_error[p].clear();
    e->trace_coeff[p].clear();
  }
  e->hist_order.clear();
  e->n_coeff = 0;
  e->last_error_slot = -1;
  e->ev_spans.clear();
  e->ev_used = 0;
}

// convergence_reached (lazyten IterativeWrapper) + is_converged (ScfBase.hh:123-126), per problem
bool update_active(Engine* e, const gscf_b200_params* prm) {
  bool changed = false;
  for (int p = 0; p < e->P; ++p) {
    if (!e->active[p]) continue;
    if (e->errnorm[p] < prm->max_error_norm) {
      e->converged[p] = 1;
      e->active[p] = 0;
      changed = true;
    } else if (e->n_iter[p] >= prm->max_iter) {
      e->status[p] = GSCF_B200_ERR_MAX_ITER;
      e->active[p] = 0;
      changed = true;
    }
  }
  return changed;
}

int final_status(Engine* e) {
  int st = GSCF_B200_OK;
  for (int p = 0; p < e->P; ++p)
    if (e->status[p] != GSCF_B200_OK) {
      st = e->status[p];
      set_last_error("problem " + std::to_string(p) + " failed with status " + std::to_string(st) +
                     " after " + std::to_string(e->n_iter[p]) + " iterations");
    }
  return st;
}

int check_ready(Engine* e, const gscf_b200_params* prm) {
  if (!e || !prm) return GSCF_B200_ERR_INVALID_ARG;
  if (!e->have_overlap) { set_last_error("overlap not set"); return GSCF_B200_ERR_INVALID_STATE; }
  if (e->cur_slot < 0) { set_last_error("problem matrix not set"); return GSCF_B200_ERR_INVALID_STATE; }
  // n_eigenpairs (ScfBase.hh:87, :315-316): the lowest k pairs.  All n eigenvalues are still computed (the Gu-Eisenstat
  // merges need every root); what k < n saves is the eigenvector part of the top-level merge of the divide & conquer,
  // the back-transformation and X Z of the n - k unwanted vectors.
  e->neig = e->n;
  if (prm->n_eigenpairs >= 0 && prm->n_eigenpairs < e->n) {
    const int kmin = std::max(e->occ[0], e->nblk > 1 ? e->occ[1] : 0);
    if (prm->n_eigenpairs < kmin) {
      set_last_error("n_eigenpairs is smaller than the number of occupied orbitals");
      return GSCF_B200_ERR_INVALID_ARG;
    }
    e->neig = (int)prm->n_eigenpairs;
  }
  for (int p = 0; p < e->P; ++p)
    if (e->status[p] != GSCF_B200_OK) { set_last_error("cannot solve a failed state"); return GSCF_B200_ERR_INVALID_STATE; }
  return GSCF_B200_OK;
}

void record_trace(Engine* e, int p, const double* coeff, int ncoeff) {
  e->trace_energy[p].push_back(e->e1[p] + e->e2[p]);
  e->trace_error[p].push_back(e->errnorm[p]);
  for (int j = 0; j < kMaxHistory; ++j)
    e->trace_coeff[p].push_back(j < ncoeff && coeff ? coeff[j] : std::numeric_limits<double>::quiet_NaN());
}

int absorb_scalars(Engine* e, const IterScalars& s, int self_index, bool diis_flags) {
  for (int p = 0; p < e->P; ++p) {
    if (!e->active[p]) continue;
    e->e1[p] = s.energies[2 * p];
    e->e2[p] = s.energies[2 * p + 1];
    const double nn = s.dots[(size_t)p * kMaxHistory + self_index];
    e->errnorm[p] = std::sqrt(nn);
    for (int b = 0; b < e->nblk; ++b)
      if (s.eiginfo[(size_t)p * e->nblk + b] != 0) {
        e->status[p] = GSCF_B200_ERR_EIGENSOLVER;  // ExcInnerEigensolverFailed
        e->active[p] = 0;
      }
    if (diis_flags && s.fail[p] != 0) {
      e->status[p] = GSCF_B200_ERR_DIIS_STEP;  // ExcDiisStepFailed
      e->active[p] = 0;
    }
    if (e->active[p] && !std::isfinite(e->errnorm[p]) && !e->error_host) {
      // non-finite built-in Pulay error: the inner solvers produced garbage.  With a HOST error function a non-finite
      // norm is the caller's business: the default residual error returns Constants<>::invalid on the first iteration
      // (ScfBase.hh:271-275), which only means "not converged" (ScfBase.hh:123-126) - real eigensolver failures are
      // still caught by eiginfo above.
      e->status[p] = GSCF_B200_ERR_EIGENSOLVER;
      e->active[p] = 0;
    }
  }
  return GSCF_B200_OK;
}

}  // namespace

namespace gscfb {
// Sub-phase timing hook used by the eigensolver (ids 8..15, see bench.py): events are only
// recorded while an engine with phase timing enabled is solving on this thread.
void phase_mark(int id, bool begin, cudaStream_t st) {
  Engine* e = g_cur_engine;
  if (!e || !e->phase_timing || id < 0 || id >= 16) return;
  if (begin) {
    e->phase_open[id] = take_event(e);
    cudaEventRecord(e->phase_open[id], st);
  } else if (e->phase_open[id]) {
    cudaEvent_t b = take_event(e);
    cudaEventRecord(b, st);
    e->ev_spans.push_back({id, {e->phase_open[id], b}});
    e->phase_open[id] = nullptr;
  }
}
}  // namespace gscfb

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

void gscf_b200_default_params(gscf_b200_params* p) {
  if (!p) return;
  p->max_iter = 100;
  p->n_eigenpairs = -1;
  p->max_error_norm = 5e-7;
  p->diis_n_prev_steps = 4;
  p->toda_n_prev_steps = 2;
  p->toda_min_fock_prefactor = 0.05;
}

int gscf_b200_engine_create(const gscf_b200_config* cfg, gscf_b200_engine** out) {
  if (!cfg || !out) return GSCF_B200_ERR_INVALID_ARG;
  *out = nullptr;
  if (cfg->n_bas < 1 || cfg->n_problems < 1 || (cfg->n_blocks != 1 && cfg->n_blocks != 2) ||
      cfg->n_alpha < 0 || cfg->n_beta < 0 || cfg->n_alpha > cfg->n_bas || cfg->n_beta > cfg->n_bas ||
      cfg->max_history < 0 || cfg->max_history > kMaxHistory) {
    set_last_error("engine_create: invalid configuration");
    return GSCF_B200_ERR_INVALID_ARG;
  }
  if ((long long)cfg->n_problems * cfg->n_blocks > 32768) {
    // the batched kernels index the matrices of one launch with gridDim.y / .z (<= 65535)
    set_last_error("engine_create: at most 32768 matrices (n_problems x n_blocks) per engine - shard larger "
                   "ensembles over engines / GPUs (gscf_b200/ensemble.py)");
    return GSCF_B200_ERR_INVALID_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_last_error("no CUDA device: gscf_b200 has no CPU fallback");
    return GSCF_B200_ERR_CUDA;
  }
  Engine* e = new Engine();
  e->cfg = *cfg;
  e->n = cfg->n_bas;
  e->ld = padded_ld(e->n);
  e->P = cfg->n_problems;
  e->nblk = cfg->n_blocks;
  e->M = std::max(cfg->max_history, 2);
  e->occ[0] = cfg->n_alpha;
  e->occ[1] = cfg->n_blocks == 2 ? cfg->n_beta : cfg->n_alpha;
  e->omax = std::max(1, std::max(e->occ[0], e->occ[1]));
  e->mstride = (long long)e->ld * e->n;
  e->pstride = e->mstride * e->nblk;
  e->eig_method = resolve_eig_method(cfg->eig_method, e->n);
  if (e->eig_method < 0) { delete e; return GSCF_B200_ERR_INVALID_ARG; }
  if (cfg->stream) {
    e->stream = (cudaStream_t)cfg->stream;
  } else {
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete e;
      set_last_error("cudaStreamCreate failed");
      return GSCF_B200_ERR_CUDA;
    }
    e->owns_stream = true;
  }
  const size_t P = e->P, PM = P * e->nblk;
  const size_t nS = cfg->shared_overlap ? (size_t)1 : PM;
  e->sx_stride = cfg->shared_overlap ? 0 : e->mstride;
  int st = GSCF_B200_OK;
  auto A = [&](auto** ptr, size_t cnt) { if (st == GSCF_B200_OK) st = dev_alloc(ptr, cnt); };
  A(&e->S, nS * e->mstride);
  A(&e->X, nS * e->mstride);
  {
    const char* oe = getenv("GSCF_B200_ORTHO");
    e->ortho = (oe && (oe[0] == 'l' || oe[0] == 'L' || oe[0] == '0')) ? 0 : 1;
    void* cw = nullptr;
    if (cudaMalloc(&cw, cholesky_inverse_worksize(e->n, (int)nS)) != cudaSuccess) cw = nullptr;
    e->cholwork = static_cast<double*>(cw);
    if (!cw) e->ortho = 0;
  }
  A(&e->Fring, P * (e->M + 1) * e->pstride);
  A(&e->Ering, P * e->M * e->pstride);
  A(&e->Fdiag, P * e->pstride);
  A(&e->Cbuf, 2 * P * e->pstride);
  A(&e->orben, 2 * PM * e->n);
  A(&e->D, P * e->pstride);
  A(&e->Aeig, P * e->pstride);
  A(&e->Zeig, P * e->pstride);
  A(&e->Tmp, P * e->pstride);
  A(&e->errwork, PM * pulay_error_work_doubles(e->n, e->omax));
  A(&e->gram, P * e->M * e->M);
  A(&e->coeff, P * kMaxHistory);
  A(&e->dots, P * kMaxHistory);
  A(&e->energies, P * 2);
  A(&e->scal, P * 4);
  A(&e->fail, P);
  A(&e->eiginfo, PM);
  A(&e->act_p, P);
  A(&e->keep_p, P);
  if (st == GSCF_B200_OK &&
      (cudaMallocHost((void**)&e->h_stage, (P * (2 * kMaxHistory + 2)) * sizeof(double)) != cudaSuccess ||
       cudaMallocHost((void**)&e->h_istage, (P + PM) * sizeof(int)) != cudaSuccess)) {
    set_last_error("cudaMallocHost failed");
    st = GSCF_B200_ERR_CUDA;
  }
  A(&e->act_m, PM);
  A(&e->act_mb[0], P);
  A(&e->act_mb[1], P);
  e->eigwork_bytes = syev_worksize(e->n, (int)PM, e->eig_method);
  if (st == GSCF_B200_OK && e->eigwork_bytes) {
    if (cudaMalloc(&e->eigwork, e->eigwork_bytes) != cudaSuccess) { set_last_error("cudaMalloc eig workspace failed"); st = GSCF_B200_ERR_CUDA; }
  }
  {
    const size_t md = multidot_worksize(kMaxHistory, e->n, e->n * e->nblk, e->P);
    if (st == GSCF_B200_OK && cudaMalloc((void**)&e->mdwork, md) != cudaSuccess) { set_last_error("cudaMalloc multidot workspace failed"); st = GSCF_B200_ERR_CUDA; }
  }
  if (st != GSCF_B200_OK) {
    gscf_b200_engine_destroy(e);
    return st;
  }
  e->n_iter.assign(P, 0); e->converged.assign(P, 0); e->status.assign(P, 0); e->active.assign(P, 1);
  e->n_mtx_applies.assign(P, 0);
  e->errnorm.assign(P, std::numeric_limits<double>::quiet_NaN());
  e->e1.assign(P, 0.0); e->e2.assign(P, 0.0); e->damp.assign(P, 1.0);
  e->trace_energy.resize(P); e->trace_error.resize(P); e->trace_coeff.resize(P);
  // identity active lists
  {
    std::vector<int> ids(PM);
    for (size_t i = 0; i < PM; ++i) ids[i] = (int)i;
    cudaMemcpy(e->act_m, ids.data(), PM * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemcpy(e->act_p, ids.data(), P * sizeof(int), cudaMemcpyHostToDevice);
    for (int b = 0; b < e->nblk; ++b) {
      for (size_t p = 0; p < P; ++p) ids[p] = (int)(p * e->nblk + b);
      cudaMemcpy(e->act_mb[b], ids.data(), P * sizeof(int), cudaMemcpyHostToDevice);
    }
  }
  *out = e;
  return GSCF_B200_OK;
}

void gscf_b200_engine_destroy(gscf_b200_engine* e) {
  if (!e) return;
  if (e->stream) cudaStreamSynchronize(e->stream);
  cudaFree(e->cholwork);
  cudaFree(e->S); cudaFree(e->X); cudaFree(e->Fring); cudaFree(e->Ering); cudaFree(e->Fdiag);
  cudaFree(e->Cbuf); cudaFree(e->orben); cudaFree(e->D); cudaFree(e->Aeig); cudaFree(e->Zeig);
  cudaFree(e->Tmp); cudaFree(e->errwork); cudaFree(e->gram); cudaFree(e->coeff); cudaFree(e->dots);
  cudaFree(e->energies); cudaFree(e->scal); cudaFree(e->fail); cudaFree(e->eiginfo);
  cudaFree(e->eigwork); cudaFree(e->mdwork); cudaFree(e->act_p); cudaFree(e->keep_p); cudaFree(e->act_m);
  cudaFree(e->act_mb[0]); cudaFree(e->act_mb[1]);
  cudaFreeHost(e->h_stage); cudaFreeHost(e->h_istage);
  for (auto ev : e->ev_pool) cudaEventDestroy(ev);
  if (e->owns_stream && e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

void* gscf_b200_engine_stream(gscf_b200_engine* e) { return e ? (void*)e->stream : nullptr; }

int gscf_b200_set_overlap(gscf_b200_engine* e, const double* S, int ld, int memloc) {
  if (!e || !S || ld < e->n) return GSCF_B200_ERR_INVALID_ARG;
  const int n = e->n;
  const int cnt_in = e->cfg.shared_overlap ? 1 : e->P;
  cudaEvent_t ov_a = nullptr, ov_b = nullptr;
  cudaEventCreate(&ov_a);
  cudaEventCreate(&ov_b);
  // per-problem overlaps are duplicated per block so that every matrix-level GEMM sees one stride
  if (e->cfg.shared_overlap) {
    GSCF_TRY(copy_in(e, e->S, e->mstride, S, ld, memloc, 1));
  } else {
    for (int b = 0; b < e->nblk; ++b) GSCF_TRY(copy_in(e, e->S + b * e->mstride, e->pstride, S, ld, memloc, cnt_in));
  }
  const int nm = e->cfg.shared_overlap ? 1 : e->P * e->nblk;
  cudaEventRecord(ov_a, e->stream);  // S is resident from here: the timed part is the factorisation
  // Aeig <- S (both set-ups destroy their input)
  GSCF_CUDA_TRY(cudaMemcpyAsync(e->Aeig, e->S, (size_t)nm * e->mstride * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
  std::vector<int> info(nm);
  std::vector<double> wh;
  if (e->ortho == 1) {
    // Cholesky: S = L L^T, X = L^-1 (chol.cu); every later product with X is triangular
    GSCF_TRY(cholesky_inverse_batched(n, e->Aeig, e->ld, e->mstride, e->Tmp, e->ld, e->mstride, e->X, e->ld, e->mstride,
                                      e->cholwork, nm, e->eiginfo, e->stream));
  } else {
    // Loewdin: X = S^-1/2 = U diag(w^-1/2) U^T, through our own eigensolver
    double* wtmp = e->orben;  // scratch: overwritten by the first eigensolve anyway
    GSCF_TRY(syev_batched(e->eig_method, n, e->Aeig, e->ld, e->mstride, wtmp, n, e->Zeig, e->ld, e->mstride,
                          nm, nullptr, e->eigwork, e->eigwork_bytes, e->eiginfo, e->stream));
    // U~ = U diag(w^-1/4);  X = U~ U~^T
    GSCF_TRY(scale_columns_pow(e->Tmp, e->ld, e->mstride, e->Zeig, e->ld, e->mstride, wtmp, n, 1, n, n, nm, e->stream));
    GSCF_TRY(gemm_batched(false, true, n, n, n, 1.0, e->Tmp, e->ld, e->mstride, e->Tmp, e->ld, e->mstride,
                          0.0, e->X, e->ld, e->mstride, nm, nullptr, e->stream));
    wh.resize((size_t)nm * n);
    GSCF_CUDA_TRY(cudaMemcpyAsync(wh.data(), wtmp, wh.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  }
  GSCF_CUDA_TRY(cudaMemcpyAsync(info.data(), e->eiginfo, nm * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  cudaEventRecord(ov_b, e->stream);
  GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
  {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ov_a, ov_b);
    e->overlap_ms = ms;  // orthogonaliser set-up (S -> X, without the copy-in of S), reported as phase 7
    cudaEventDestroy(ov_a);
    cudaEventDestroy(ov_b);
  }
  for (int i = 0; i < nm; ++i) {
    if (e->ortho == 1) {
      if (info[i] != 0) { set_last_error("overlap matrix is not positive definite"); return GSCF_B200_ERR_INVALID_ARG; }
    } else {
      if (info[i] != 0) { set_last_error("eigensolver failed on the overlap matrix"); return GSCF_B200_ERR_EIGENSOLVER; }
      if (!(wh[(size_t)i * n] > 0.0)) { set_last_error("overlap matrix is not positive definite"); return GSCF_B200_ERR_INVALID_ARG; }
    }
  }
  e->have_overlap = true;
  return GSCF_B200_OK;
}

int gscf_b200_set_fock_builder_device(gscf_b200_engine* e, gscf_b200_fock_device_fn fn, void* ctx) {
  if (!e) return GSCF_B200_ERR_INVALID_ARG;
  e->fock_dev = fn; e->fock_dev_ctx = ctx; e->fock_host = nullptr;
  return GSCF_B200_OK;
}
int gscf_b200_set_fock_builder_host(gscf_b200_engine* e, gscf_b200_fock_host_fn fn, void* ctx) {
  if (!e) return GSCF_B200_ERR_INVALID_ARG;
  e->fock_host = fn; e->fock_host_ctx = ctx; e->fock_dev = nullptr;
  return GSCF_B200_OK;
}
int gscf_b200_set_error_host(gscf_b200_engine* e, gscf_b200_error_host_fn fn, void* ctx) {
  if (!e) return GSCF_B200_ERR_INVALID_ARG;
  e->error_host = fn; e->error_host_ctx = ctx;
  return GSCF_B200_OK;
}
int gscf_b200_set_hook(gscf_b200_engine* e, gscf_b200_hook_fn fn, void* ctx) {
  if (!e) return GSCF_B200_ERR_INVALID_ARG;
  e->hook = fn; e->hook_ctx = ctx;
  return GSCF_B200_OK;
}

int gscf_b200_set_problem_matrix(gscf_b200_engine* e, const double* F, int ld, int memloc,
                                 const double* e1_host, const double* e2_host) {
  if (!e || !F || ld < e->n) return GSCF_B200_ERR_INVALID_ARG;
  GSCF_TRY(copy_in_groups(e, e->Fslot(e->M), e->fring_pstride(), F, ld, memloc));
  std::vector<double> en((size_t)e->P * 2, 0.0);
  for (int p = 0; p < e->P; ++p) {
    e->e1[p] = e1_host ? e1_host[p] : 0.0;
    e->e2[p] = e2_host ? e2_host[p] : 0.0;
    en[2 * p] = e->e1[p]; en[2 * p + 1] = e->e2[p];
    e->status[p] = GSCF_B200_OK;
  }
  GSCF_CUDA_TRY(cudaMemcpyAsync(e->energies, en.data(), en.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
  e->cur_slot = e->M;
  return GSCF_B200_OK;
}

int gscf_b200_set_guess(gscf_b200_engine* e, const double* C, int ld, int memloc, const double* orben_host) {
  if (!e || !C || ld < e->n) return GSCF_B200_ERR_INVALID_ARG;
  GSCF_TRY(copy_in_groups(e, e->Ccur(), e->pstride, C, ld, memloc));
  if (orben_host)
    GSCF_CUDA_TRY(cudaMemcpyAsync(e->orben_cur(), orben_host, (size_t)e->P * e->nblk * e->n * sizeof(double), cudaMemcpyHostToDevice, e->stream));
  e->have_coeff = true;
  for (int p = 0; p < e->P; ++p) { e->active[p] = 1; e->status[p] = GSCF_B200_OK; }
  GSCF_TRY(upload_active(e));
  std::vector<int> act(e->P);
  for (int p = 0; p < e->P; ++p) act[p] = p;
  GSCF_TRY(step_density(e, e->P));
  cudaEvent_t fk_a = nullptr, fk_b = nullptr;
  cudaEventCreate(&fk_a);
  cudaEventCreate(&fk_b);
  cudaEventRecord(fk_a, e->stream);
  GSCF_TRY(step_fock(e, e->M, e->P, act));  // ScfStateBase.hh:301-302
  cudaEventRecord(fk_b, e->stream);
  std::vector<double> en((size_t)e->P * 2);
  GSCF_CUDA_TRY(cudaMemcpyAsync(en.data(), e->energies, en.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
  {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, fk_a, fk_b);
    e->guess_fock_ms = ms;  // the plugin's share of installing the guess (phase 14): not part of the SCF path
    cudaEventDestroy(fk_a);
    cudaEventDestroy(fk_b);
  }
  for (int p = 0; p < e->P; ++p) { e->e1[p] = en[2 * p]; e->e2[p] = en[2 * p + 1]; }
  e->ev_spans.clear();
  e->ev_used = 0;
  return GSCF_B200_OK;
}

// Guess from a matrix: eigenvectors of G against S (e.g. the core guess G = H, or the Loewdin
// guess of examples/hartree_fock/loewdin_guess.cc with G = S handled by the caller), installed
// like obtain_guess_from (ScfStateBase.hh:291-303): the Fock builder is called once.
// G holds P matrices (n x n, ld) that are used for every block.
int gscf_b200_set_guess_from_matrix(gscf_b200_engine* e, const double* G, int ld, int memloc) {
  if (!e || !G || ld < e->n) return GSCF_B200_ERR_INVALID_ARG;
  if (!e->have_overlap) { set_last_error("overlap not set"); return GSCF_B200_ERR_INVALID_STATE; }
  for (int b = 0; b < e->nblk; ++b)
    GSCF_TRY(copy_in(e, e->Fslot(e->M) + b * e->mstride, e->fring_pstride(), G, ld, memloc, e->P));
  for (int p = 0; p < e->P; ++p) { e->active[p] = 1; e->status[p] = GSCF_B200_OK; }
  GSCF_TRY(upload_active(e));
  std::vector<int> act(e->P);
  for (int p = 0; p < e->P; ++p) act[p] = p;
  GSCF_TRY(step_eigensolve(e, e->Fslot(e->M), e->fring_pstride(), e->P));
  GSCF_TRY(step_density(e, e->P));
  GSCF_TRY(step_fock(e, e->M, e->P, act));
  std::vector<double> en((size_t)e->P * 2);
  std::vector<int> info((size_t)e->P * e->nblk);
  GSCF_CUDA_TRY(cudaMemcpyAsync(en.data(), e->energies, en.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  GSCF_CUDA_TRY(cudaMemcpyAsync(info.data(), e->eiginfo, info.size() * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
  for (int p = 0; p < e->P; ++p) { e->e1[p] = en[2 * p]; e->e2[p] = en[2 * p + 1]; }
  for (int i : info) if (i != 0) { set_last_error("eigensolver failed in the guess"); return GSCF_B200_ERR_EIGENSOLVER; }
  e->ev_spans.clear();
  e->ev_used = 0;
  return GSCF_B200_OK;
}

// ------------------------------------------------------------------------------------------------
// PlainScf::solve_state (PlainScf.hh:121-147)
// ------------------------------------------------------------------------------------------------
int gscf_b200_solve_plain(gscf_b200_engine* e, const gscf_b200_params* prm) {
  GSCF_TRY(check_ready(e, prm));
  EngineScope scope(e);
  begin_solve(e);
  IterScalars sc;
  SlotList self{};
  self.n = 1; self.slot[0] = 0;  // error kept in error-ring slot 0
  {
    Span total(e, PH_TOTAL, true);
    bool first = true;
    while (true) {
      const bool changed = update_active(e, prm);
      const int nact = n_active(e);
      if (nact == 0) break;
      if (changed || first) GSCF_TRY(upload_active(e));
      first = false;
      std::vector<int> act;
      for (int p = 0; p < e->P; ++p) if (e->active[p]) { act.push_back(p); e->n_iter[p]++; }
      const int it = e->n_iter[act[0]];
      GSCF_TRY(call_hook(e, GSCF_B200_EV_BEFORE_ITERATION, it));
      e->diagmat_kind = 0;
      e->n_coeff = 0;
      GSCF_TRY(step_eigensolve(e, e->Fslot(e->cur_slot), e->fring_pstride(), nact));  // :129-130
      GSCF_TRY(call_hook(e, GSCF_B200_EV_UPDATE_EIGENPAIRS, it));
      GSCF_TRY(step_density(e, nact));
      GSCF_TRY(step_fock(e, e->cur_slot, nact, act));  // in place: sole owner (:139-140)
      GSCF_TRY(call_hook(e, GSCF_B200_EV_UPDATE_PROBLEM_MATRIX, it));
      {
        Span sp(e, PH_ERROR);
        GSCF_TRY(step_error(e, e->cur_slot, e->Eslot(0), e->ering_pstride(), nact));  // :143
        GSCF_TRY(step_multidot(e, e->Eslot(0), e->ering_pstride(), self, nact));
      }
      e->last_error_slot = 0;
      GSCF_TRY(fetch_scalars(e, sc));
      absorb_scalars(e, sc, 0, false);
      for (int p : act) record_trace(e, p, nullptr, 0);
      GSCF_TRY(call_hook(e, GSCF_B200_EV_AFTER_ITERATION, it));
    }
  }
  GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
  collect_timings(e);
  return final_status(e);
}

// ------------------------------------------------------------------------------------------------
// PulayDiisScf::solve_state (PulayDiisScf.hh:283-324)
// ------------------------------------------------------------------------------------------------
int gscf_b200_solve_diis(gscf_b200_engine* e, const gscf_b200_params* prm) {
  GSCF_TRY(check_ready(e, prm));
  EngineScope scope(e);
  const int m = prm->diis_n_prev_steps;
  if (m < 1 || m > e->M) {
    set_last_error("diis_n_prev_steps must be in [1, max_history]");
    return GSCF_B200_ERR_INVALID_ARG;
  }
  begin_solve(e);
  // the initial matrix must not live in a ring slot < m: it is in slot M (set_problem_matrix)
  if (e->cur_slot != e->M) {
    // a previous solve left the current matrix in a ring slot: move it to the spare slot
    for (int p = 0; p < e->P; ++p)
      GSCF_CUDA_TRY(cudaMemcpyAsync(e->Fslot(e->M) + p * e->fring_pstride(), e->Fslot(e->cur_slot) + p * e->fring_pstride(),
                                    e->pstride * sizeof(double), cudaMemcpyDeviceToDevice, e->stream));
    e->cur_slot = e->M;
  }
  IterScalars sc;
  std::vector<double> hcoeff((size_t)e->P * kMaxHistory, 0.0);
  int head = 0;  // next ring slot to (over)write
  {
    Span total(e, PH_TOTAL, true);
    bool first = true;
    while (true) {
      const bool changed = update_active(e, prm);
      const int nact = n_active(e);
      if (nact == 0) break;
      if (changed || first) GSCF_TRY(upload_active(e));
      first = false;
      std::vector<int> act;
      for (int p = 0; p < e->P; ++p) if (e->active[p]) { act.push_back(p); e->n_iter[p]++; }
      const int it = e->n_iter[act[0]];
      GSCF_TRY(call_hook(e, GSCF_B200_EV_BEFORE_ITERATION, it));

      // update_diis_coefficients (:358-393)
      const int k = (int)e->hist_order.size();
      SlotList order{};
      order.n = k;
      for (int i = 0; i < k; ++i) order.slot[i] = e->hist_order[i];
      {
        Span sp(e, PH_COEFF);
        if (k >= 2) GSCF_TRY(diis_solve(e->gram, e->M, (long long)e->M * e->M, order, e->coeff, kMaxHistory, e->fail, nact, e->act_p, e->stream));
      }
      e->n_coeff = k;
      // update_diis_diagmat (:395-433)
      const double* Fd;
      long long fd_pstride;
      {
        Span sp(e, PH_DIAGMAT);
        if (k == 0) {
          e->diagmat_kind = 0;
          Fd = e->Fslot(e->cur_slot); fd_pstride = e->fring_pstride();
        } else if (k == 1) {
          e->diagmat_kind = 1; e->diagmat_slot = order.slot[0];  // c = [1]
          Fd = e->Fslot(order.slot[0]); fd_pstride = e->fring_pstride();
        } else {
          e->diagmat_kind = 2;
          GSCF_TRY(axpby_n(e->Fdiag, e->pstride, e->Fring, e->pstride, e->fring_pstride(), order, e->coeff,
                           kMaxHistory, e->n, e->n * e->nblk, e->ld, nact, e->act_p, e->stream));
          Fd = e->Fdiag; fd_pstride = e->pstride;
        }
      }
      GSCF_TRY(call_hook(e, GSCF_B200_EV_NEW_DIAGMAT, it));
      GSCF_TRY(step_eigensolve(e, Fd, fd_pstride, nact));  // :299
      for (int p : act) e->n_mtx_applies[p] += 0;  // dense solver reports 0 applies (:304-305)
      GSCF_TRY(call_hook(e, GSCF_B200_EV_UPDATE_EIGENPAIRS, it));
      GSCF_TRY(step_density(e, nact));
      // update_problem_matrix + push_back (:314-315): the new matrix goes into ring slot `head`,
      // evicting the oldest entry when the ring is full
      const int slot = head;
      if ((int)e->hist_order.size() == m) e->hist_order.erase(e->hist_order.begin());
      GSCF_TRY(step_fock(e, slot, nact, act));
      e->hist_order.push_back(slot);
      head = (head + 1) % m;
      GSCF_TRY(call_hook(e, GSCF_B200_EV_UPDATE_PROBLEM_MATRIX, it));
      // errors.push_back(calculate_error) + overlaps (:318-320)
      SlotList valid{};
      valid.n = (int)e->hist_order.size();
      // newest first, like the reference's overlap vector (:440-463)
      for (int i = 0; i < valid.n; ++i) valid.slot[i] = e->hist_order[valid.n - 1 - i];
      {
        Span sp(e, PH_ERROR);
        GSCF_TRY(step_error(e, slot, e->Eslot(slot), e->ering_pstride(), nact));
        GSCF_TRY(step_multidot(e, e->Eslot(slot), e->ering_pstride(), valid, nact));
        gram_update_kernel<<<ceil_div(nact, 64), 64, 0, e->stream>>>(e->gram, e->M, e->dots, valid, slot, e->act_p, nact);
        count_launch();
        GSCF_CHECK_LAUNCH();
      }
      e->last_error_slot = slot;
      GSCF_TRY(fetch_scalars(e, sc, k >= 2 ? hcoeff.data() : nullptr));
      if (k == 1) for (int p : act) hcoeff[(size_t)p * kMaxHistory] = 1.0;
      absorb_scalars(e, sc, 0, k >= 2);
      for (int p : act) record_trace(e, p, hcoeff.data() + (size_t)p * kMaxHistory, k);
      GSCF_TRY(call_hook(e, GSCF_B200_EV_AFTER_ITERATION, it));
    }
  }
  GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
  collect_timings(e);
  return final_status(e);
}

// ------------------------------------------------------------------------------------------------
// TruncatedOptDampScf::solve_state (TruncatedOptDampScf.hh:239-289), P problems in lock-step
//
// Per problem the reference keeps prev_problem_matrix_ptr (:61), its energies and damping_coeff (:70).  Here the
// current matrix of every problem lives in one ring slot (CUR) and the kept previous one in another (PREV): when a
// problem keeps its matrix (:273-275) the matrix is copied CUR -> PREV (one n x n copy per keeper and iteration,
// nothing against the eigensolve), so that all kernels still see one slot per role for the whole ensemble.
// ------------------------------------------------------------------------------------------------
int gscf_b200_solve_toda(gscf_b200_engine* e, const gscf_b200_params* prm) {
  GSCF_TRY(check_ready(e, prm));
  EngineScope scope(e);
  if (prm->toda_n_prev_steps != 2) {  // :243 assert_implemented
    set_last_error("toda_n_prev_steps must be 2");
    return GSCF_B200_ERR_NOT_IMPLEMENTED;
  }
  begin_solve(e);
  IterScalars sc;
  SlotList self{};
  self.n = 1; self.slot[0] = 0;
  const int n = e->n, ld = e->ld, P = e->P;
  const int CUR = e->cur_slot;
  const int PREV = CUR == 0 ? 1 : 0;
  std::vector<char> prev_valid(P, 0);                 // prev_problem_matrix_ptr != nullptr
  std::vector<double> e1_prev(P, 0.0), e2_prev(P, 0.0);  // energies of the matrix in PREV
  std::vector<double> damp(P, 1.0), lam(P, 1.0);      // damping_coeff (:70, :76)
  std::vector<double> orb, tr4((size_t)P * 4), cf((size_t)P * kMaxHistory, 0.0);
  std::vector<int> keep;
  // PREV must hold finite numbers for the problems without a previous matrix (their coefficient is an exact 0)
  for (int p = 0; p < P; ++p)
    GSCF_CUDA_TRY(cudaMemsetAsync(e->Fslot(PREV) + p * e->fring_pstride(), 0, (size_t)e->pstride * sizeof(double), e->stream));
  {
    Span total(e, PH_TOTAL, true);
    bool first = true;
    while (true) {
      const bool changed = update_active(e, prm);
      const int nact = n_active(e);
      if (nact == 0) break;
      if (changed || first) GSCF_TRY(upload_active(e));
      first = false;
      std::vector<int> act;
      for (int p = 0; p < P; ++p) if (e->active[p]) { act.push_back(p); e->n_iter[p]++; }
      const int it = e->n_iter[act[0]];
      GSCF_TRY(call_hook(e, GSCF_B200_EV_BEFORE_ITERATION, it));
      // ---- update_damping_coefficients (:349-368) -------------------------------------------------
      {
        Span sp(e, PH_COEFF);
        bool need_short = false, need_gemm = false;
        for (int p : act)
          if (prev_valid[p]) (damp[p] == 1 ? need_short : need_gemm) = true;
        if (need_short) {  // trace_fprev_dcur shortcut (:309-319): sums of occupied orbital energies
          orb.resize((size_t)P * e->nblk * n);
          GSCF_CUDA_TRY(cudaMemcpyAsync(orb.data(), e->orben_cur(), orb.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        }
        if (need_gemm) {  // :325-345: tr(C_o^T F_prev C_o) per block, never forming the n x n product
          for (int b = 0; b < e->nblk; ++b) {
            const int o = e->occ[b];
            GemmParams g = gemm_params(n, o, n, 1.0, e->Fslot(PREV) + b * e->mstride, ld, e->Ccur() + b * e->mstride, ld,
                                       0.0, e->Tmp + b * e->mstride, ld);
            g.sA = e->fring_pstride(); g.sB = e->pstride; g.sC = e->pstride; g.batch_map = e->act_p; g.entries = e->P;
            GSCF_TRY(launch_gemm(false, false, g, nact, n, o, e->stream));
            GSCF_TRY(multidot(e->Ccur() + b * e->mstride, e->pstride, e->Tmp + b * e->mstride, 0, e->pstride, self, n, o,
                              ld, nact, e->act_p, e->scal + b, 4, e->mdwork, e->stream));
          }
          GSCF_CUDA_TRY(cudaMemcpyAsync(tr4.data(), e->scal, tr4.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        }
        if (need_short || need_gemm) GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
        for (int p : act) {
          if (!prev_valid[p]) { damp[p] = 1.0; continue; }
          double tr;
          if (damp[p] == 1) {
            const double* w = orb.data() + (size_t)p * e->nblk * n;
            double sum_a = 0.0;
            for (int i = 0; i < e->occ[0]; ++i) sum_a += w[i];
            if (e->nblk == 1) {
              tr = 2. * sum_a;
            } else {
              double s2 = sum_a;
              for (int i = 0; i < e->occ[1]; ++i) s2 += w[(size_t)n + i];
              tr = s2;
            }
          } else {
            e->n_mtx_applies[p] += 1;
            tr = e->nblk == 1 ? 2. * tr4[(size_t)p * 4] : tr4[(size_t)p * 4] + tr4[(size_t)p * 4 + 1];
          }
          const double oda_s = tr - e1_prev[p] - 2. * e2_prev[p];              // :360-361
          const double oda_c = e->e1[p] - tr + e->e2[p] + e2_prev[p];          // :362-364
          damp[p] = compute_damping_coeff(oda_s, oda_c, prm->toda_min_fock_prefactor);
          if (damp[p] == 1) prev_valid[p] = 0;                                 // :367
        }
      }
      for (int p : act) e->damp[p] = damp[p];
      // ---- update_oda_diagmat (:396-409) -------------------------------------------------------------
      const double* Fd;
      long long fd_pstride;
      {
        Span sp(e, PH_DIAGMAT);
        bool any_prev = false;
        for (int p : act) any_prev = any_prev || prev_valid[p];
        if (!any_prev) {
          e->diagmat_kind = 0;
          Fd = e->Fslot(CUR); fd_pstride = e->fring_pstride();
          e->n_coeff = 0;
        } else {
          e->diagmat_kind = 2;
          SlotList two{};
          two.n = 2; two.slot[0] = PREV; two.slot[1] = CUR;
          for (int p : act) {  // no previous matrix: 0 * PREV + 1 * CUR is CUR exactly
            cf[(size_t)p * kMaxHistory] = prev_valid[p] ? 1. - damp[p] : 0.0;
            cf[(size_t)p * kMaxHistory + 1] = prev_valid[p] ? damp[p] : 1.0;
          }
          GSCF_CUDA_TRY(cudaMemcpyAsync(e->coeff, cf.data(), cf.size() * sizeof(double), cudaMemcpyHostToDevice, e->stream));
          GSCF_TRY(axpby_n(e->Fdiag, e->pstride, e->Fring, e->pstride, e->fring_pstride(), two, e->coeff,
                           kMaxHistory, n, n * e->nblk, ld, nact, e->act_p, e->stream));
          Fd = e->Fdiag; fd_pstride = e->pstride;
          e->n_coeff = 2;
        }
      }
      GSCF_TRY(call_hook(e, GSCF_B200_EV_NEW_DIAGMAT, it));
      GSCF_TRY(step_eigensolve(e, Fd, fd_pstride, nact));  // :253
      GSCF_TRY(call_hook(e, GSCF_B200_EV_UPDATE_EIGENPAIRS, it));
      // ---- keep / flip (:273-277) ---------------------------------------------------------------------
      keep.clear();
      for (int p : act) {
        lam[p] = damp[p];
        if (damp[p] > 0.5) {
          keep.push_back(p);
          prev_valid[p] = 1;
          e1_prev[p] = e->e1[p]; e2_prev[p] = e->e2[p];
        } else {
          damp[p] = 1 - damp[p];
        }
        e->damp[p] = damp[p];
      }
      if (!keep.empty()) {
        GSCF_CUDA_TRY(cudaMemcpyAsync(e->keep_p, keep.data(), keep.size() * sizeof(int), cudaMemcpyHostToDevice, e->stream));
        GSCF_TRY(scale_copy(e->Fslot(PREV), ld, e->fring_pstride(), e->Fslot(CUR), ld, e->fring_pstride(), 1.0, n,
                            n * e->nblk, (int)keep.size(), e->keep_p, e->stream));
        GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));  // `keep` is reused next iteration
      }
      // ---- update_problem_matrix (:282): in place, the kept copy is in PREV --------------------------------
      GSCF_TRY(step_density(e, nact));
      GSCF_TRY(step_fock(e, CUR, nact, act));
      GSCF_TRY(call_hook(e, GSCF_B200_EV_UPDATE_PROBLEM_MATRIX, it));
      {
        Span sp(e, PH_ERROR);
        GSCF_TRY(step_error(e, CUR, e->Eslot(0), e->ering_pstride(), nact));  // :285
        GSCF_TRY(step_multidot(e, e->Eslot(0), e->ering_pstride(), self, nact));
      }
      e->last_error_slot = 0;
      GSCF_TRY(fetch_scalars(e, sc));
      absorb_scalars(e, sc, 0, false);
      for (int p : act) record_trace(e, p, &lam[p], 1);
      GSCF_TRY(call_hook(e, GSCF_B200_EV_AFTER_ITERATION, it));
    }
  }
  GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
  collect_timings(e);
  return final_status(e);
}

// ------------------------------------------------------------------------------------------------
// getters
// ------------------------------------------------------------------------------------------------
int gscf_b200_get_n_iter(gscf_b200_engine* e, int* v) { if (!e || !v) return GSCF_B200_ERR_INVALID_ARG; std::copy(e->n_iter.begin(), e->n_iter.end(), v); return GSCF_B200_OK; }
int gscf_b200_get_converged(gscf_b200_engine* e, int* v) { if (!e || !v) return GSCF_B200_ERR_INVALID_ARG; std::copy(e->converged.begin(), e->converged.end(), v); return GSCF_B200_OK; }
int gscf_b200_get_status(gscf_b200_engine* e, int* v) { if (!e || !v) return GSCF_B200_ERR_INVALID_ARG; std::copy(e->status.begin(), e->status.end(), v); return GSCF_B200_OK; }
int gscf_b200_get_error_norm(gscf_b200_engine* e, double* v) { if (!e || !v) return GSCF_B200_ERR_INVALID_ARG; std::copy(e->errnorm.begin(), e->errnorm.end(), v); return GSCF_B200_OK; }
int gscf_b200_get_energies(gscf_b200_engine* e, double* a, double* b) {
  if (!e) return GSCF_B200_ERR_INVALID_ARG;
  if (a) std::copy(e->e1.begin(), e->e1.end(), a);
  if (b) std::copy(e->e2.begin(), e->e2.end(), b);
  return GSCF_B200_OK;
}
int gscf_b200_get_n_mtx_applies(gscf_b200_engine* e, long long* v) { if (!e || !v) return GSCF_B200_ERR_INVALID_ARG; std::copy(e->n_mtx_applies.begin(), e->n_mtx_applies.end(), v); return GSCF_B200_OK; }
int gscf_b200_get_damping_coeff(gscf_b200_engine* e, double* v) { if (!e || !v) return GSCF_B200_ERR_INVALID_ARG; std::copy(e->damp.begin(), e->damp.end(), v); return GSCF_B200_OK; }

int gscf_b200_get_n_eigenpairs(gscf_b200_engine* e) { return e ? ((e->neig > 0 && e->neig < e->n) ? e->neig : e->n) : -1; }

int gscf_b200_get_orben(gscf_b200_engine* e, double* w) {
  if (!e || !w) return GSCF_B200_ERR_INVALID_ARG;
  GSCF_CUDA_TRY(cudaMemcpyAsync(w, e->orben_cur(), (size_t)e->P * e->nblk * e->n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
  return GSCF_B200_OK;
}
int gscf_b200_get_coeff(gscf_b200_engine* e, double* C, int ld) {
  if (!e || !C || ld < e->n) return GSCF_B200_ERR_INVALID_ARG;
  return copy_out(e, C, ld, e->Ccur(), e->pstride, e->mstride);
}
int gscf_b200_get_prev_coeff(gscf_b200_engine* e, double* C, int ld) {
  if (!e || !C || ld < e->n) return GSCF_B200_ERR_INVALID_ARG;
  return copy_out(e, C, ld, e->Cprev(), e->pstride, e->mstride);
}
int gscf_b200_get_fock(gscf_b200_engine* e, double* F, int ld) {
  if (!e || !F || ld < e->n || e->cur_slot < 0) return GSCF_B200_ERR_INVALID_ARG;
  return copy_out(e, F, ld, e->Fslot(e->cur_slot), e->fring_pstride(), e->mstride);
}
int gscf_b200_get_diagmat(gscf_b200_engine* e, double* F, int ld) {
  if (!e || !F || ld < e->n) return GSCF_B200_ERR_INVALID_ARG;
  if (e->diagmat_kind == 2) return copy_out(e, F, ld, e->Fdiag, e->pstride, e->mstride);
  const int slot = e->diagmat_kind == 1 ? e->diagmat_slot : e->cur_slot;
  if (slot < 0) return GSCF_B200_ERR_INVALID_STATE;
  return copy_out(e, F, ld, e->Fslot(slot), e->fring_pstride(), e->mstride);
}
int gscf_b200_get_density(gscf_b200_engine* e, double* D, int ld) {
  if (!e || !D || ld < e->n) return GSCF_B200_ERR_INVALID_ARG;
  return copy_out(e, D, ld, e->D, e->pstride, e->mstride);
}
int gscf_b200_get_error(gscf_b200_engine* e, double* E, int ld) {
  if (!e || !E || ld < e->n || e->last_error_slot < 0) return GSCF_B200_ERR_INVALID_ARG;
  return copy_out(e, E, ld, e->Eslot(e->last_error_slot), e->ering_pstride(), e->mstride);
}
int gscf_b200_get_history_size(gscf_b200_engine* e) { return e ? (int)e->hist_order.size() : -1; }
// ring entry i (oldest = 0) of the error / problem-matrix history (PulayDiisScfState::errors :60, prev_problem_matrix_ptrs :56)
int gscf_b200_get_error_entry(gscf_b200_engine* e, int entry, double* E, int ld) {
  if (!e || !E || ld < e->n || entry < 0 || entry >= (int)e->hist_order.size()) return GSCF_B200_ERR_INVALID_ARG;
  return copy_out(e, E, ld, e->Eslot(e->hist_order[entry]), e->ering_pstride(), e->mstride);
}
int gscf_b200_get_fock_entry(gscf_b200_engine* e, int entry, double* F, int ld) {
  if (!e || !F || ld < e->n || entry < 0 || entry >= (int)e->hist_order.size()) return GSCF_B200_ERR_INVALID_ARG;
  return copy_out(e, F, ld, e->Fslot(e->hist_order[entry]), e->fring_pstride(), e->mstride);
}
int gscf_b200_get_diis_coefficients(gscf_b200_engine* e, double* c, int* n_coeff) {
  if (!e || !c || !n_coeff) return GSCF_B200_ERR_INVALID_ARG;
  *n_coeff = e->n_coeff;
  for (int p = 0; p < e->P; ++p) {
    const auto& tc = e->trace_coeff[p];
    for (int j = 0; j < kMaxHistory; ++j)
      c[(size_t)p * kMaxHistory + j] = tc.size() >= (size_t)kMaxHistory ? tc[tc.size() - kMaxHistory + j]
                                                                        : std::numeric_limits<double>::quiet_NaN();
  }
  return GSCF_B200_OK;
}
int gscf_b200_get_current_diis_coefficients(gscf_b200_engine* e, double* c, int k) {
  if (!e || !c || k < 0 || k > kMaxHistory) return GSCF_B200_ERR_INVALID_ARG;
  std::vector<double> h((size_t)e->P * kMaxHistory);
  GSCF_CUDA_TRY(cudaMemcpyAsync(h.data(), e->coeff, h.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
  for (int p = 0; p < e->P; ++p)
    for (int j = 0; j < k; ++j) c[(size_t)p * k + j] = h[(size_t)p * kMaxHistory + j];
  return GSCF_B200_OK;
}
int gscf_b200_get_error_overlaps(gscf_b200_engine* e, int problem, int entry, double* ov, int* len) {
  if (!e || !ov || !len || problem < 0 || problem >= e->P) return GSCF_B200_ERR_INVALID_ARG;
  const int k = (int)e->hist_order.size();
  if (entry < 0 || entry >= k) return GSCF_B200_ERR_INVALID_ARG;
  std::vector<double> g((size_t)e->M * e->M);
  GSCF_CUDA_TRY(cudaMemcpyAsync(g.data(), e->gram + (size_t)problem * e->M * e->M, g.size() * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  GSCF_CUDA_TRY(cudaStreamSynchronize(e->stream));
  // <i|i>, <i|i-1>, ... for the errors still in the ring (PulayDiisScf.hh:62-81)
  *len = entry + 1;
  for (int j = 0; j <= entry; ++j) ov[j] = g[(size_t)e->hist_order[entry] * e->M + e->hist_order[entry - j]];
  return GSCF_B200_OK;
}
int gscf_b200_get_trace(gscf_b200_engine* e, int problem, double* energy, double* error_norm, double* coeff,
                        int max_rows, int* n_rows) {
  if (!e || problem < 0 || problem >= e->P || !n_rows) return GSCF_B200_ERR_INVALID_ARG;
  const int rows = (int)e->trace_energy[problem].size();
  *n_rows = rows;
  const int r = std::min(rows, max_rows);
  for (int i = 0; i < r; ++i) {
    if (energy) energy[i] = e->trace_energy[problem][i];
    if (error_norm) error_norm[i] = e->trace_error[problem][i];
    if (coeff) for (int j = 0; j < kMaxHistory; ++j) coeff[(size_t)i * kMaxHistory + j] = e->trace_coeff[problem][(size_t)i * kMaxHistory + j];
  }
  return GSCF_B200_OK;
}
int gscf_b200_get_timings(gscf_b200_engine* e, double* scf_ms, double* fock_ms) {
  if (!e) return GSCF_B200_ERR_INVALID_ARG;
  if (scf_ms) *scf_ms = e->scf_ms;
  if (fock_ms) *fock_ms = e->fock_ms;
  return GSCF_B200_OK;
}
int gscf_b200_get_phase_timings(gscf_b200_engine* e, double* ms, int nn) {
  if (!e || !ms) return GSCF_B200_ERR_INVALID_ARG;
  for (int i = 0; i < nn && i < 16; ++i) ms[i] = e->phase_ms[i];
  for (int i = 16; i < nn && i < 32; ++i) ms[i] = (double)e->phase_count[i - 16];
  return GSCF_B200_OK;
}
int gscf_b200_set_phase_timing(gscf_b200_engine* e, int enabled) {
  if (!e) return GSCF_B200_ERR_INVALID_ARG;
  e->phase_timing = enabled != 0;
  return GSCF_B200_OK;
}

}  // extern "C"
