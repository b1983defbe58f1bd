// Shared host-side helpers of libothello_b200 (defined in env_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oth {
int cuda_status(cudaError_t e);         // cudaError -> OTH_* code, records the message
int sm_count();                         // SMs of the current device (148 on B200)
int grid_for(int64_t threads, int block);  // blocks: multiples of the SM count, capped at 8 per SM
}  // namespace oth
