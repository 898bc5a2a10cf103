// hps_comm.h -- NCCL plumbing for particle-sharded beams.  The reference is
// single-GPU and has no counterpart.  NCCL is opened with dlopen(RTLD_LOCAL) so that the library
// never clashes with another NCCL copy living in the same process (e.g. the one bundled with torch).
#ifndef HPS_COMM_H
#define HPS_COMM_H

#include <cuda_runtime.h>
#include <string>

namespace hps {

struct Comm {
  int rank = 0, nranks = 1;
  void* nccl = nullptr;  // ncclComm_t
};

int comm_unique_id(void* out128);
int comm_init(Comm&, int rank, int nranks, const void* id128, std::string* err);
void comm_destroy(Comm&);
// in-place collectives on device buffers, enqueued on `stream`
int comm_allgather(const Comm&, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t);
int comm_allreduce_sum_f64(const Comm&, double* buf, size_t count, cudaStream_t);
// host scalar broadcast (blocking; used by rare getters only)
int comm_bcast_double(const Comm&, double* v, int root, cudaStream_t);

}  // namespace hps
#endif
