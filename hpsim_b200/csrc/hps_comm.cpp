// hps_comm.cpp -- see hps_comm.h
#include "hps_comm.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>

namespace hps {
namespace {

struct Api {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

Api* api() {
  static Api a;
  if (a.handle || !a.err.empty()) return &a;
  const char* names[] = {getenv("HPSIM_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    a.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (a.handle) break;
  }
  if (!a.handle) { a.err = std::string("cannot dlopen NCCL: ") + dlerror(); return &a; }
#define SYM(field, name)                                         \
  *(void**)(&a.field) = dlsym(a.handle, name);                   \
  if (!a.field) { a.err = std::string("NCCL symbol missing: ") + name; return &a; }
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllGather, "ncclAllGather");
  SYM(AllReduce, "ncclAllReduce");
  SYM(Broadcast, "ncclBroadcast");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return &a;
}

}  // namespace

int comm_unique_id(void* out128) {
  Api* a = api();
  if (!a->err.empty()) return 1;
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (a->GetUniqueId(&id) != ncclSuccess) return 1;
  std::memcpy(out128, &id, 128);
  return 0;
}

int comm_init(Comm& c, int rank, int nranks, const void* id128, std::string* err) {
  Api* a = api();
  if (!a->err.empty()) { if (err) *err = a->err; return 1; }
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  ncclComm_t comm;
  ncclResult_t r = a->CommInitRank(&comm, nranks, id, rank);
  if (r != ncclSuccess) { if (err) *err = std::string("ncclCommInitRank: ") + a->GetErrorString(r); return 1; }
  c.rank = rank; c.nranks = nranks; c.nccl = comm;
  return 0;
}

void comm_destroy(Comm& c) {
  if (c.nccl) api()->CommDestroy((ncclComm_t)c.nccl);
  c.nccl = nullptr; c.rank = 0; c.nranks = 1;
}

int comm_allgather(const Comm& c, const void* send, void* recv, size_t bytes, cudaStream_t s) {
  return api()->AllGather(send, recv, bytes, ncclChar, (ncclComm_t)c.nccl, s) != ncclSuccess;
}
int comm_allreduce_sum_f64(const Comm& c, double* buf, size_t count, cudaStream_t s) {
  return api()->AllReduce(buf, buf, count, ncclDouble, ncclSum, (ncclComm_t)c.nccl, s) != ncclSuccess;
}
int comm_bcast_double(const Comm& c, double* v, int root, cudaStream_t s) {
  double* d = nullptr;
  if (cudaMalloc(&d, sizeof(double)) != cudaSuccess) return 1;
  int rc = 0;
  if (c.rank == root) rc |= cudaMemcpyAsync(d, v, sizeof(double), cudaMemcpyHostToDevice, s) != cudaSuccess;
  rc |= api()->Broadcast(d, d, 1, ncclDouble, root, (ncclComm_t)c.nccl, s) != ncclSuccess;
  rc |= cudaMemcpyAsync(v, d, sizeof(double), cudaMemcpyDeviceToHost, s) != cudaSuccess;
  rc |= cudaStreamSynchronize(s) != cudaSuccess;
  cudaFree(d);
  return rc;
}

}  // namespace hps
