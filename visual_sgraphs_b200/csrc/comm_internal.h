// comm_internal.h — the communicator behind vsg_comm_* (include/vsg_cuda.h): NCCL, loaded at run time.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/vsg_cuda.h"

namespace vsg {
int comm_rank(const vsg_comm *c);
int comm_size(const vsg_comm *c);
// byte-wise collectives / point-to-point on `stream` (device buffers)
vsg_status comm_all_gather(vsg_comm *c, const void *send_dev, void *recv_dev, size_t bytes_per_rank, cudaStream_t stream);
vsg_status comm_send(vsg_comm *c, const void *buf_dev, size_t bytes, int peer, cudaStream_t stream);
vsg_status comm_recv(vsg_comm *c, void *buf_dev, size_t bytes, int peer, cudaStream_t stream);
}  // namespace vsg
