#include "common_host.h"

#include <cstdio>
#include <cstring>

namespace slslam {

static thread_local char g_last_error[256] = "";

void set_last_error(const char* msg) {
  snprintf(g_last_error, sizeof(g_last_error), "%s", msg ? msg : "");
}

int ensure_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    set_last_error(e != cudaSuccess ? cudaGetErrorString(e) : "no CUDA device");
    cudaGetLastError();
    return SLSLAM_ERR_CUDA;   // no CPU fallback by design
  }
  if (device >= 0) {
    if (device >= count) { set_last_error("device index out of range"); return SLSLAM_ERR_CUDA; }
    CUDA_TRY(cudaSetDevice(device));
  }
  int dev = 0, major = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) { set_last_error("device is not sm_100 (this library ships sm_100a code only)"); return SLSLAM_ERR_CUDA; }
  return SLSLAM_OK;
}

}  // namespace slslam

extern "C" {

int slslam_version(void) { return 100; }

const char* slslam_strerror(int code) {
  switch (code) {
    case SLSLAM_OK: return "ok";
    case SLSLAM_ERR_INVALID: return "invalid argument";
    case SLSLAM_ERR_UNSUPPORTED: return "problem exceeds a kernel limit";
    case SLSLAM_ERR_CUDA: return "CUDA error or no sm_100 device (no CPU fallback)";
    case SLSLAM_ERR_NUMERICAL: return "non-finite input";
    default: return "unknown error";
  }
}

const char* slslam_last_error(void) { return slslam::g_last_error; }

int slslam_device_count(void) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess) { cudaGetLastError(); return 0; }
  int usable = 0;
  for (int d = 0; d < count; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++usable;
  }
  return usable;
}

}  // extern "C"
