// mem_pool.h -- process-wide cache of device and pinned-host blocks.
//
// A registration needs a few large, identically sized buffers every time (the distance-transform
// working volume, the DT grid, the 0.6 GB heap spill slab, pinned result/task arrays).  cudaMalloc /
// cudaFree / cudaMallocHost take the driver lock and cost from a millisecond up to hundreds of
// milliseconds when another process on the host queries the driver; on a 180 GB part there is no
// reason to give the blocks back between handles.  Freed blocks are kept (up to GOICP_POOL_MAX_MB,
// default 8192 MB of device memory) and handed out again best-fit; goicp_trim_memory() returns them
// to the driver.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace goicp {

cudaError_t pool_alloc(void** p, size_t bytes);          // device memory on the current device
void pool_free(void* p);
cudaError_t pool_alloc_host(void** p, size_t bytes);     // pinned host memory
void pool_free_host(void* p);
void pool_trim();                                        // give every cached block back to the driver

} // namespace goicp
