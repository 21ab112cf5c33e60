// comm.cuh — NCCL plumbing for the partitioned PDLP (one host thread per GPU: in one process or in several).
// NCCL is resolved with dlopen at first use so that the single-GPU path (and the R package) has no
// link-time NCCL dependency.  Used: all-gather (the x-bar and y blocks every iteration, scaling vectors at
// setup), allreduce (scalar residual partials), grouped send/recv (one-off exchange that builds each rank's
// column block of A from the ranks' row blocks).
#pragma once
#include "common.cuh"
#include <nccl.h>

namespace elp {

struct Comm {
    bool active = false;
    int nranks = 1, rank = 0;
    ncclComm_t comm = nullptr;
};
Comm& comm();                         // the calling thread's communicator (comm.cu)

void comm_unique_id(void* id128);
void comm_init(int nranks, int rank, const void* id128);
void comm_destroy();
// in-place allreduce on `stream`; no-op when the communicator is inactive
void comm_allreduce_sum(double* buf, size_t count, cudaStream_t stream);
void comm_allreduce_max(double* buf, size_t count, cudaStream_t stream);
// in-place all-gather: rank r's `count` doubles live at buf + r*count before the call, everyone's after it
void comm_allgather(double* buf, size_t count, cudaStream_t stream);
void comm_allgather_bytes(void* buf, size_t bytes_per_rank, cudaStream_t stream);
void comm_group_start();
void comm_group_end();
void comm_send(const void* buf, size_t bytes, int peer, cudaStream_t stream);
void comm_recv(void* buf, size_t bytes, int peer, cudaStream_t stream);

}  // namespace elp
