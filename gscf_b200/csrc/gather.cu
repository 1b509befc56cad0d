// The path's only collective: an all-gather of per-problem result records of an ensemble that is sharded
// over the GPUs of one box by problem index - energies, error norms, iteration counts and convergence flags,
// 32 bytes per problem.  NCCL is bound at run time (dlopen of libnccl.so.2: the copy the host process already
// loaded, e.g. torch's, or the system one), so libgscf_b200.so itself has no link-time dependency on it and a
// single-GPU host never needs it.  The reference has no counterpart (it has no parallelism of any kind).
#include <dlfcn.h>

#include <cstring>
#include <vector>

#include "engine.cuh"

using namespace gscfb;

namespace {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
typedef int (*fn_get_unique_id)(NcclUniqueId*);
typedef int (*fn_comm_init_rank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*fn_comm_destroy)(NcclComm);
typedef int (*fn_all_gather)(const void*, void*, size_t, int /*ncclDataType_t*/, NcclComm, cudaStream_t);
typedef const char* (*fn_error_string)(int);
typedef int (*fn_get_version)(int*);
constexpr int kNcclInt8 = 0;

struct NcclApi {
  void* handle = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_all_gather all_gather = nullptr;
  fn_error_string error_string = nullptr;
  fn_get_version get_version = nullptr;
  std::string origin;
};

const NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle ? &api : nullptr;
  tried = true;
  std::vector<std::string> names;
  if (const char* env = getenv("GSCF_B200_NCCL_LIB")) names.push_back(env);
  names.push_back("libnccl.so.2");
  names.push_back("libnccl.so");
  for (const auto& nm : names) {
    void* h = dlopen(nm.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (!h) continue;
    api.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
    api.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
    api.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
    api.all_gather = (fn_all_gather)dlsym(h, "ncclAllGather");
    api.error_string = (fn_error_string)dlsym(h, "ncclGetErrorString");
    api.get_version = (fn_get_version)dlsym(h, "ncclGetVersion");
    if (api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_gather) {
      api.handle = h;
      api.origin = nm;
      return &api;
    }
    dlclose(h);
  }
  return nullptr;
}

int nccl_fail(const NcclApi* api, const char* what, int rc) {
  set_last_error(std::string(what) + " failed: " + (api && api->error_string ? api->error_string(rc) : "nccl error") +
                 " (" + std::to_string(rc) + ")");
  return GSCF_B200_ERR_CUDA;
}

}  // namespace

struct gscf_b200_comm {
  NcclComm comm = nullptr;
  int world = 1, rank = 0;
};

extern "C" {

int gscf_b200_nccl_version(void) {
  const NcclApi* api = nccl_api();
  int v = 0;
  if (!api || !api->get_version || api->get_version(&v) != 0) return 0;
  return v;
}

int gscf_b200_comm_unique_id(void* id128) {
  if (!id128) return GSCF_B200_ERR_INVALID_ARG;
  const NcclApi* api = nccl_api();
  if (!api) { set_last_error("libnccl.so.2 not found (set GSCF_B200_NCCL_LIB)"); return GSCF_B200_ERR_NOT_IMPLEMENTED; }
  NcclUniqueId id;
  const int rc = api->get_unique_id(&id);
  if (rc != 0) return nccl_fail(api, "ncclGetUniqueId", rc);
  std::memcpy(id128, &id, sizeof(id));
  return GSCF_B200_OK;
}

int gscf_b200_comm_create(const void* id128, int world, int rank, gscf_b200_comm** out) {
  if (!id128 || !out || world < 1 || rank < 0 || rank >= world) return GSCF_B200_ERR_INVALID_ARG;
  *out = nullptr;
  const NcclApi* api = nccl_api();
  if (!api) { set_last_error("libnccl.so.2 not found (set GSCF_B200_NCCL_LIB)"); return GSCF_B200_ERR_NOT_IMPLEMENTED; }
  NcclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  gscf_b200_comm* c = new gscf_b200_comm();
  c->world = world; c->rank = rank;
  const int rc = api->comm_init_rank(&c->comm, world, id, rank);
  if (rc != 0) { delete c; return nccl_fail(api, "ncclCommInitRank", rc); }
  *out = c;
  return GSCF_B200_OK;
}

void gscf_b200_comm_destroy(gscf_b200_comm* c) {
  if (!c) return;
  const NcclApi* api = nccl_api();
  if (api && c->comm) api->comm_destroy(c->comm);
  delete c;
}

int gscf_b200_ensemble_gather(gscf_b200_engine* e, gscf_b200_comm* c, const int* counts, gscf_b200_record* out) {
  if (!e || !c || !counts || !out) return GSCF_B200_ERR_INVALID_ARG;
  const NcclApi* api = nccl_api();
  if (!api) { set_last_error("libnccl.so.2 not found (set GSCF_B200_NCCL_LIB)"); return GSCF_B200_ERR_NOT_IMPLEMENTED; }
  if (counts[c->rank] != e->P) {
    set_last_error("ensemble_gather: counts[rank] differs from the engine's number of problems");
    return GSCF_B200_ERR_INVALID_ARG;
  }
  int cmax = 0;
  for (int r = 0; r < c->world; ++r) {
    if (counts[r] < 0) return GSCF_B200_ERR_INVALID_ARG;
    cmax = counts[r] > cmax ? counts[r] : cmax;
  }
  if (cmax == 0) return GSCF_B200_OK;
  // ncclAllGather wants equal contributions: every rank sends cmax records (its own, zero-padded)
  std::vector<gscf_b200_record> mine((size_t)cmax);
  std::memset(mine.data(), 0, mine.size() * sizeof(gscf_b200_record));
  for (int p = 0; p < e->P; ++p) {
    mine[p].energy_total = e->e1[p] + e->e2[p];
    mine[p].error_norm = e->errnorm[p];
    mine[p].n_iter = e->n_iter[p];
    mine[p].converged = e->converged[p];
    mine[p].status = e->status[p];
    mine[p].reserved = 0;
  }
  const size_t sendbytes = (size_t)cmax * sizeof(gscf_b200_record);
  char* dbuf = nullptr;
  GSCF_CUDA_TRY(cudaMalloc((void**)&dbuf, sendbytes * (size_t)(c->world + 1)));
  int st = GSCF_B200_OK;
  std::vector<gscf_b200_record> all((size_t)cmax * c->world);
  do {
    if (cudaMemcpyAsync(dbuf, mine.data(), sendbytes, cudaMemcpyHostToDevice, e->stream) != cudaSuccess) { st = GSCF_B200_ERR_CUDA; break; }
    const int rc = api->all_gather(dbuf, dbuf + sendbytes, sendbytes, kNcclInt8, c->comm, e->stream);
    if (rc != 0) { st = nccl_fail(api, "ncclAllGather", rc); break; }
    if (cudaMemcpyAsync(all.data(), dbuf + sendbytes, sendbytes * c->world, cudaMemcpyDeviceToHost, e->stream) != cudaSuccess ||
        cudaStreamSynchronize(e->stream) != cudaSuccess) {
      set_last_error("ensemble_gather: CUDA copy / synchronise failed");
      st = GSCF_B200_ERR_CUDA;
    }
  } while (0);
  cudaFree(dbuf);
  if (st != GSCF_B200_OK) return st;
  size_t o = 0;
  for (int r = 0; r < c->world; ++r)
    for (int p = 0; p < counts[r]; ++p) out[o++] = all[(size_t)r * cmax + p];
  return GSCF_B200_OK;
}

}  // extern "C"
