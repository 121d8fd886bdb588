// deepv_b200 — the exchange step of Ulysses sequence parallelism.
//
// The denoiser shards the video tokens over the ranks of a group; around every attention the
// ranks swap "my tokens, your heads" for "all tokens, my heads" — an all-to-all with equal
// blocks.  dv_mmdit_forward only sees a function pointer (dv_exchange_fn); this file provides the
// NCCL implementation of it.  NCCL is not linked: the library the host process already loaded
// (torch ships libnccl.so.2) is resolved with dlopen, and the communicator is created here from
// a unique id that the host exchanges over its own process group.
#include <dlfcn.h>

#include <cstring>

#include "../../include/deepv_b200.h"
#include "common.cuh"

using namespace dv;

namespace {

struct NcclUniqueId {
  char internal[128];
};
typedef void* NcclComm;
typedef int (*FnGetUniqueId)(NcclUniqueId*);
typedef int (*FnCommInitRank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*FnCommDestroy)(NcclComm);
typedef int (*FnGroup)(void);
typedef int (*FnSendRecv)(const void*, size_t, int /*ncclDataType_t*/, int, NcclComm, cudaStream_t);
typedef const char* (*FnErrStr)(int);
typedef int (*FnAllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);

struct NcclApi {
  void* h = nullptr;
  FnGetUniqueId get_id = nullptr;
  FnCommInitRank init = nullptr;
  FnCommDestroy destroy = nullptr;
  FnGroup group_start = nullptr, group_end = nullptr;
  FnSendRecv send = nullptr;
  int (*recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  FnErrStr err = nullptr;
  FnAllReduce all_reduce = nullptr;
};

NcclApi g_nccl;

int load_nccl(const char* path) {
  if (g_nccl.h) return 0;
  void* h = nullptr;
  if (path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);  // already mapped by torch in most hosts
  if (!h) {
    set_error("dv_comm: cannot load libnccl.so.2 (%s)", dlerror());
    return DV_ERR_INVALID;
  }
  g_nccl.get_id = reinterpret_cast<FnGetUniqueId>(dlsym(h, "ncclGetUniqueId"));
  g_nccl.init = reinterpret_cast<FnCommInitRank>(dlsym(h, "ncclCommInitRank"));
  g_nccl.destroy = reinterpret_cast<FnCommDestroy>(dlsym(h, "ncclCommDestroy"));
  g_nccl.group_start = reinterpret_cast<FnGroup>(dlsym(h, "ncclGroupStart"));
  g_nccl.group_end = reinterpret_cast<FnGroup>(dlsym(h, "ncclGroupEnd"));
  g_nccl.send = reinterpret_cast<FnSendRecv>(dlsym(h, "ncclSend"));
  g_nccl.recv = reinterpret_cast<decltype(g_nccl.recv)>(dlsym(h, "ncclRecv"));
  g_nccl.err = reinterpret_cast<FnErrStr>(dlsym(h, "ncclGetErrorString"));
  g_nccl.all_reduce = reinterpret_cast<FnAllReduce>(dlsym(h, "ncclAllReduce"));
  if (!g_nccl.get_id || !g_nccl.init || !g_nccl.destroy || !g_nccl.group_start || !g_nccl.group_end ||
      !g_nccl.send || !g_nccl.recv || !g_nccl.all_reduce || !g_nccl.err) {
    set_error("dv_comm: libnccl.so.2 lacks a required symbol");
    return DV_ERR_INVALID;
  }
  g_nccl.h = h;
  return 0;
}

#define DV_NCCL(call)                                                                       \
  do {                                                                                      \
    int _r = (call);                                                                        \
    if (_r != 0) {                                                                          \
      set_error("%s -> NCCL error %d (%s)", #call, _r, g_nccl.err ? g_nccl.err(_r) : "?"); \
      return DV_ERR_CUDA;                                                                   \
    }                                                                                       \
  } while (0)

}  // namespace

struct dv_comm {
  NcclComm comm = nullptr;
  int rank = 0, world = 1;
  int* token = nullptr;  // device word all-reduced by the barrier form of the exchange
};

extern "C" int dv_comm_unique_id(const char* nccl_path, void* id128) {
  DV_REQUIRE(id128, "dv_comm_unique_id: null buffer");
  int rc = load_nccl(nccl_path);
  if (rc) return rc;
  NcclUniqueId id;
  DV_NCCL(g_nccl.get_id(&id));
  memcpy(id128, &id, sizeof(id));
  return DV_OK;
}

extern "C" int dv_comm_create(const char* nccl_path, const void* id128, int rank, int world,
                              dv_comm** out) {
  DV_REQUIRE(id128 && out && world >= 1 && rank >= 0 && rank < world, "dv_comm_create: bad argument");
  int rc = load_nccl(nccl_path);
  if (rc) return rc;
  NcclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  dv_comm* c = new dv_comm();
  c->rank = rank;
  c->world = world;
  int r = g_nccl.init(&c->comm, world, id, rank);
  if (r != 0) {
    set_error("ncclCommInitRank -> NCCL error %d (%s)", r, g_nccl.err ? g_nccl.err(r) : "?");
    delete c;
    return DV_ERR_CUDA;
  }
  if (cudaMalloc(&c->token, 2 * sizeof(int)) != cudaSuccess || cudaMemset(c->token, 0, 2 * sizeof(int)) != cudaSuccess) {
    set_error("dv_comm_create: cannot allocate the barrier word");
    dv_comm_destroy(c);   // frees the word (if any) and the communicator
    return DV_ERR_CUDA;
  }
  *out = c;
  return DV_OK;
}

extern "C" void dv_comm_destroy(dv_comm* c) {
  if (!c) return;
  if (c->token) cudaFree(c->token);
  if (c->comm && g_nccl.destroy) g_nccl.destroy(c->comm);
  delete c;
}

// dv_exchange_fn over NCCL: block j of `send` goes to rank j, block i of `recv` comes from rank i.
extern "C" int dv_comm_exchange(void* user, const void* send_dev, void* recv_dev,
                                long long bytes_per_peer, void* stream) {
  dv_comm* c = reinterpret_cast<dv_comm*>(user);
  DV_REQUIRE(c && c->comm && bytes_per_peer >= 0, "dv_comm_exchange: bad argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (bytes_per_peer == 0) {
    // barrier form (peer-memory Ulysses: the payload already went over NVLink as plain stores):
    // a one-word all-reduce orders every rank's preceding kernels before every rank's following ones
    DV_NCCL(g_nccl.all_reduce(c->token, c->token + 1, 1, 2 /*ncclInt32*/, 0 /*ncclSum*/, c->comm, st));
    return DV_OK;
  }
  DV_REQUIRE(send_dev && recv_dev, "dv_comm_exchange: null buffer");
  const char* s = reinterpret_cast<const char*>(send_dev);
  char* r = reinterpret_cast<char*>(recv_dev);
  DV_NCCL(g_nccl.group_start());
  int first_err = 0;   // the group is always closed, or later collectives would hang instead of failing
  for (int p = 0; p < c->world && first_err == 0; ++p) {
    first_err = g_nccl.send(s + static_cast<size_t>(p) * bytes_per_peer, static_cast<size_t>(bytes_per_peer),
                            0 /*ncclInt8*/, p, c->comm, st);
    if (first_err == 0)
      first_err = g_nccl.recv(r + static_cast<size_t>(p) * bytes_per_peer, static_cast<size_t>(bytes_per_peer), 0,
                              p, c->comm, st);
  }
  const int end_err = g_nccl.group_end();
  if (first_err != 0 || end_err != 0) {
    const int e = first_err != 0 ? first_err : end_err;
    set_error("dv_comm_exchange: NCCL error %d (%s)", e, g_nccl.err(e));
    return DV_ERR_CUDA;
  }
  return DV_OK;
}
