// Shared host-side plumbing of libb200radiomics: error state and CUDA call checking.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include <string>

#include "../../include/b200radiomics.h"

namespace rb {
std::string& last_error_ref();
int fail(int code, const char* fmt, ...);
}  // namespace rb

#define RB_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return rb::fail(_e == cudaErrorMemoryAllocation ? RB_ERR_NOMEM : RB_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                      cudaGetErrorString(_e), __FILE__, __LINE__);                             \
  } while (0)

#define RB_LAUNCH_CHECK() RB_CUDA(cudaGetLastError())
