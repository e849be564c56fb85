// Shared host helpers: error reporting and device selection for the C ABI (include/slslam_b200.h).
#pragma once
#include <cuda_runtime.h>

#include "../../include/slslam_b200.h"

namespace slslam {

void set_last_error(const char* msg);
// Selects `device` (or keeps the current one when < 0) and checks it is an sm_100 part.
int ensure_device(int device);

#define CUDA_TRY(x)                                                   \
  do {                                                                \
    cudaError_t e_ = (x);                                             \
    if (e_ != cudaSuccess) {                                          \
      ::slslam::set_last_error(cudaGetErrorString(e_));               \
      cudaGetLastError();                                             \
      return SLSLAM_ERR_CUDA;                                         \
    }                                                                 \
  } while (0)

#define CUDA_TRY_OR(x, cleanup)                                       \
  do {                                                                \
    cudaError_t e_ = (x);                                             \
    if (e_ != cudaSuccess) {                                          \
      ::slslam::set_last_error(cudaGetErrorString(e_));               \
      cudaGetLastError();                                             \
      cleanup;                                                        \
    }                                                                 \
  } while (0)

}  // namespace slslam
