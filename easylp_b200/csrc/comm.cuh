// comm.cuh — NCCL plumbing for the row-partitioned PDLP (one process per GPU).
// NCCL is resolved with dlopen at first use so that the single-GPU path (and the R package) has no
// link-time NCCL dependency.  Only allreduce is needed: the partial A'y (sum), the scalar residual
// partials (sum, packed in the tail of the same buffer) and the Ruiz column maxima (max).
#pragma once
#include "common.cuh"
#include <nccl.h>

namespace elp {

struct Comm {
    bool active = false;
    int nranks = 1, rank = 0;
    ncclComm_t comm = nullptr;
};
Comm& comm();                         // process-wide communicator (comm.cu)

void comm_unique_id(void* id128);
void comm_init(int nranks, int rank, const void* id128);
void comm_destroy();
// in-place allreduce on `stream`; no-op when the communicator is inactive
void comm_allreduce_sum(double* buf, size_t count, cudaStream_t stream);
void comm_allreduce_max(double* buf, size_t count, cudaStream_t stream);

}  // namespace elp
