#include "common_host.h"

#include <algorithm>
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

// Measured fp64 FMA throughput of the device: every thread runs 16 independent DFMA chains (enough ILP to cover the
// pipe latency at any occupancy), enough CTAs to fill every SM.  The denominator of the "fp64 pipe" reading aid in
// bench.py: the LBA / PO kernels are fp64-issue and latency bound, not HBM bound.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
  double a[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = seed + 1e-3 * (threadIdx.x + 17 * k);
  const double m = 1.0 + 1e-9 * seed, c = 1e-12;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = fma(a[k], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += a[k];
  if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;     // never true: keeps the chains alive
}

}  // namespace slslam

extern "C" {

int slslam_measure_fp64_peak(int32_t device, double* tflops_out, double* sm_clock_mhz_out) {
  slslam::set_last_error("");
  if (!tflops_out) return SLSLAM_ERR_INVALID;
  int rc = slslam::ensure_device(device);
  if (rc != SLSLAM_OK) return rc;
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  double* d_out = nullptr;
  const int ctas = sms * 8, iters = 20000;
  CUDA_TRY(cudaMalloc((void**)&d_out, (size_t)ctas * 256 * 8));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    slslam::dfma_peak_kernel<<<ctas, 256>>>(d_out, iters, 1.0 + rep);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaGetLastError(); break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16.0 * (double)iters * 256.0 * ctas;
    if (rep > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d_out);
  *tflops_out = best;
  if (sm_clock_mhz_out) *sm_clock_mhz_out = khz / 1e3;
  return best > 0.0 ? SLSLAM_OK : SLSLAM_ERR_CUDA;
}

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
