// comm.h - process-wide communicator state of libgfk (one process per GPU, NCCL over NVLink).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gfk {
int comm_size();  // 1 when no communicator was initialised
int comm_rank();
// all-gather of one 16-byte (value, index) record per rank: recv[r] <- rank r's *send
int comm_allgather16(const void* send, void* recv, cudaStream_t st);
}  // namespace gfk
