// In-library NCCL communicator for the gallery-sharded evaluation: one
// process per GPU, one communicator per process.  The two collectives of an evaluation --
// all-gather of the same-identity records, all-reduce(sum) of the rank counts -- are issued on
// the caller's compute stream, directly behind the kernels that produce their inputs, with no host
// round trip in between.
//
// libnccl is not linked: it is resolved at demo_comm_init() time with dlopen, preferring the copy
// that is already loaded in the process (PyTorch's bundled NCCL), so the library still loads on a
// machine without NCCL or without a GPU.
#include <dlfcn.h>
#include <nccl.h>  // types and enums only

#include <mutex>

#include "common.cuh"

using namespace demo;

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

std::mutex g_mu;
NcclApi g_api;
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_world = 1;

template <class F>
bool bind(void* h, const char* name, F* fn) {
  *fn = reinterpret_cast<F>(dlsym(h, name));
  return *fn != nullptr;
}

int load_api() {
  if (g_api.handle) return DEMO_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the process already uses (PyTorch's)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("demo_comm: libnccl.so.2 not found (%s)", dlerror());
    return DEMO_ERR_UNSUPPORTED;
  }
  NcclApi a;
  a.handle = h;
  const bool ok = bind(h, "ncclGetUniqueId", &a.GetUniqueId) && bind(h, "ncclCommInitRank", &a.CommInitRank) &&
                  bind(h, "ncclCommDestroy", &a.CommDestroy) && bind(h, "ncclCommGetAsyncError", &a.CommGetAsyncError) &&
                  bind(h, "ncclAllGather", &a.AllGather) && bind(h, "ncclAllReduce", &a.AllReduce) &&
                  bind(h, "ncclBroadcast", &a.Broadcast) && bind(h, "ncclGetErrorString", &a.GetErrorString) &&
                  bind(h, "ncclGetVersion", &a.GetVersion);
  if (!ok) {
    set_error("demo_comm: libnccl.so.2 lacks a required symbol");
    return DEMO_ERR_UNSUPPORTED;
  }
  g_api = a;
  return DEMO_OK;
}

#define DEMO_CHECK_NCCL(expr)                                                              \
  do {                                                                                     \
    ncclResult_t _r = (expr);                                                              \
    if (_r != ncclSuccess) {                                                               \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_api.GetErrorString(_r));   \
      return DEMO_ERR_CUDA;                                                                \
    }                                                                                      \
  } while (0)

int require_comm() {
  if (!g_comm) {
    set_error("demo_comm: no communicator (call demo_comm_init first)");
    return DEMO_ERR_INVALID;
  }
  return DEMO_OK;
}

}  // namespace

extern "C" {

// 1 when libnccl can be resolved in this process (host-only query), else 0.
int demo_comm_available(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  return load_api() == DEMO_OK ? 1 : 0;
}

// NCCL version code of the resolved library (e.g. 22809), 0 when unavailable.
int demo_comm_nccl_version(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (load_api() != DEMO_OK) return 0;
  int v = 0;
  return g_api.GetVersion(&v) == ncclSuccess ? v : 0;
}

// Rank 0 creates the 128-byte id and hands it to the other ranks out of band (the Python host
// broadcasts it through torch.distributed / the rendezvous store).
int demo_comm_unique_id(void* id_out_host) {
  std::lock_guard<std::mutex> lk(g_mu);
  DEMO_REQUIRE(id_out_host, "demo_comm_unique_id: null pointer");
  DEMO_TRY(load_api());
  ncclUniqueId id;
  DEMO_CHECK_NCCL(g_api.GetUniqueId(&id));
  memcpy(id_out_host, &id, sizeof(id));
  return DEMO_OK;
}

// Collective over all ranks; the CUDA device that is current at this call is the rank's device.
int demo_comm_init(int rank, int world, const void* unique_id_host) {
  std::lock_guard<std::mutex> lk(g_mu);
  DEMO_REQUIRE(unique_id_host && world >= 1 && rank >= 0 && rank < world, "demo_comm_init: bad arguments");
  DEMO_TRY(load_api());
  if (g_comm) {
    set_error("demo_comm_init: communicator already initialised (rank %d of %d)", g_rank, g_world);
    return DEMO_ERR_INVALID;
  }
  ncclUniqueId id;
  memcpy(&id, unique_id_host, sizeof(id));
  DEMO_CHECK_NCCL(g_api.CommInitRank(&g_comm, world, id, rank));
  g_rank = rank;
  g_world = world;
  return DEMO_OK;
}

int demo_comm_destroy(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_comm) {
    DEMO_CHECK_NCCL(g_api.CommDestroy(g_comm));
    g_comm = nullptr;
  }
  g_rank = 0;
  g_world = 1;
  return DEMO_OK;
}

// rank / world of the communicator; returns 1 when one exists, 0 otherwise
int demo_comm_info(int* rank, int* world) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (rank) *rank = g_rank;
  if (world) *world = g_world;
  return g_comm ? 1 : 0;
}

// Asynchronous errors of the communicator (ncclCommGetAsyncError); DEMO_OK when healthy.
int demo_comm_check(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  DEMO_TRY(require_comm());
  ncclResult_t async = ncclSuccess;
  DEMO_CHECK_NCCL(g_api.CommGetAsyncError(g_comm, &async));
  if (async != ncclSuccess && async != ncclInProgress) {
    set_error("demo_comm: asynchronous NCCL error: %s", g_api.GetErrorString(async));
    return DEMO_ERR_CUDA;
  }
  return DEMO_OK;
}

// recv[r * bytes_per_rank ...] = send of rank r (device buffers; recv may contain send in place)
int demo_comm_all_gather(const void* send, void* recv, size_t bytes_per_rank, void* stream) {
  DEMO_TRY(require_comm());
  DEMO_REQUIRE(send && recv, "demo_comm_all_gather: null pointer");
  if (bytes_per_rank == 0) return DEMO_OK;
  DEMO_CHECK_NCCL(g_api.AllGather(send, recv, bytes_per_rank, ncclUint8, g_comm, static_cast<cudaStream_t>(stream)));
  return DEMO_OK;
}

// buf[i] = sum over ranks of buf[i] (uint32, in place): the rank counts are additive over gallery shards
int demo_comm_all_reduce_sum_u32(void* buf, size_t count, void* stream) {
  DEMO_TRY(require_comm());
  DEMO_REQUIRE(buf, "demo_comm_all_reduce: null pointer");
  if (count == 0) return DEMO_OK;
  DEMO_CHECK_NCCL(g_api.AllReduce(buf, buf, count, ncclUint32, ncclSum, g_comm, static_cast<cudaStream_t>(stream)));
  return DEMO_OK;
}

int demo_comm_broadcast(void* buf, size_t bytes, int root, void* stream) {
  DEMO_TRY(require_comm());
  DEMO_REQUIRE(buf && root >= 0 && root < g_world, "demo_comm_broadcast: bad arguments");
  if (bytes == 0) return DEMO_OK;
  DEMO_CHECK_NCCL(g_api.Broadcast(buf, buf, bytes, ncclUint8, root, g_comm, static_cast<cudaStream_t>(stream)));
  return DEMO_OK;
}

}  // extern "C"
