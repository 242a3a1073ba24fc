// NCCL communicator owned by the library: one rank per process / GPU.
//
// The scorer needs two collectives on its own stream, in its own launch order:
//   * the per-layer all-gather of the prefix K/V slices when the shared prefix is split over the
//     ranks (engine.cu: layer_pack -> all-gather -> layer_unpack), and
//   * ONE all-reduce(MIN) over the packed 64-bit winner key (`allreduce_min_packed`).
// NCCL is resolved at run time from the libnccl.so.2 that is already mapped into the process (the
// Python host imports torch first, so this is torch's bundled copy); the library itself does not
// link against it and single-GPU use never touches it. The unique id is created on rank 0 and
// handed to the other ranks by the host (torch.distributed broadcast, 128 bytes).
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/bimodal_b200.h"

struct bma_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, size = 1;
};

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.handle != nullptr ? &api : nullptr;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (h == nullptr) {
    api.error = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
    return nullptr;
  }
  bool ok = true;
  auto sym = [&](const char* name) {
    void* p = dlsym(h, name);
    if (p == nullptr) { ok = false; api.error = std::string("NCCL symbol missing: ") + name; }
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  if (!ok) return nullptr;
  api.handle = h;
  return &api;
}

thread_local std::string g_comm_error;

int comm_fail(const char* what, ncclResult_t r) {
  NcclApi* a = nccl();
  g_comm_error = std::string(what) + ": " + (a != nullptr && a->GetErrorString != nullptr ? a->GetErrorString(r) : "NCCL error");
  fprintf(stderr, "[bimodal_b200] %s\n", g_comm_error.c_str());
  return 1000 + static_cast<int>(r);
}

}  // namespace

// used by engine.cu
int comm_rank(const bma_comm* c) { return c != nullptr ? c->rank : 0; }
int comm_size(const bma_comm* c) { return c != nullptr ? c->size : 1; }
int comm_allgather_bytes(bma_comm* c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t st) {
  NcclApi* a = nccl();
  if (a == nullptr || c == nullptr || c->comm == nullptr) return 1;
  ncclResult_t r = a->AllGather(send, recv, bytes_per_rank, ncclInt8, c->comm, st);
  return r == ncclSuccess ? 0 : comm_fail("ncclAllGather", r);
}

extern "C" {

const char* bma_comm_last_error(void) { return g_comm_error.c_str(); }

int bma_comm_unique_id(void* out128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  NcclApi* a = nccl();
  if (a == nullptr) { g_comm_error = "NCCL is not available in this process"; return 1; }
  ncclUniqueId id;
  ncclResult_t r = a->GetUniqueId(&id);
  if (r != ncclSuccess) return comm_fail("ncclGetUniqueId", r);
  memcpy(out128, &id, sizeof(id));
  return 0;
}

bma_comm* bma_comm_create(const void* id128, int size, int rank) {
  NcclApi* a = nccl();
  if (a == nullptr || id128 == nullptr || size < 1 || rank < 0 || rank >= size) {
    g_comm_error = a == nullptr ? "NCCL is not available in this process" : "comm_create: bad arguments";
    return nullptr;
  }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  bma_comm* c = new bma_comm();
  c->rank = rank; c->size = size;
  ncclResult_t r = a->CommInitRank(&c->comm, size, id, rank);
  if (r != ncclSuccess) {
    comm_fail("ncclCommInitRank", r);
    delete c;
    return nullptr;
  }
  return c;
}

void bma_comm_destroy(bma_comm* c) {
  if (c == nullptr) return;
  NcclApi* a = nccl();
  if (a != nullptr && c->comm != nullptr) a->CommDestroy(c->comm);
  delete c;
}

int bma_comm_rank(const bma_comm* c) { return comm_rank(c); }
int bma_comm_size(const bma_comm* c) { return comm_size(c); }

// In-place MIN over `n` packed 64-bit keys (orderable(loss) << 31 | candidate index): the lowest loss
// wins, the lowest index among equal losses -- torch.argmin's answer (reference :550-551).
int bma_comm_allreduce_min_packed(bma_comm* c, unsigned long long* keys, int n, void* stream) {
  NcclApi* a = nccl();
  if (a == nullptr || c == nullptr || c->comm == nullptr) { g_comm_error = "allreduce_min_packed: no communicator"; return 1; }
  ncclResult_t r = a->AllReduce(keys, keys, static_cast<size_t>(n), ncclUint64, ncclMin, c->comm,
                                static_cast<cudaStream_t>(stream));
  return r == ncclSuccess ? 0 : comm_fail("ncclAllReduce(min, u64)", r);
}

// Winner key (MIN) and early-stop flag (MAX) of one step in ONE fused NCCL launch.
int bma_comm_reduce_winner(bma_comm* c, unsigned long long* key, int* all_match, void* stream) {
  NcclApi* a = nccl();
  if (a == nullptr || c == nullptr || c->comm == nullptr) { g_comm_error = "reduce_winner: no communicator"; return 1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (all_match == nullptr) return bma_comm_allreduce_min_packed(c, key, 1, stream);
  ncclResult_t r = a->GroupStart();
  if (r == ncclSuccess) r = a->AllReduce(key, key, 1, ncclUint64, ncclMin, c->comm, st);
  if (r == ncclSuccess) r = a->AllReduce(all_match, all_match, 1, ncclInt32, ncclMax, c->comm, st);
  ncclResult_t e = a->GroupEnd();
  if (r == ncclSuccess) r = e;
  return r == ncclSuccess ? 0 : comm_fail("ncclAllReduce(winner)", r);
}

}  // extern "C"
