// Host side of the RANSAC scoring entry point (include/slslam_b200.h: slslam_ransac_score).
#include <algorithm>
#include <cstring>
#include <vector>

#include "common_host.h"
#include "ransac_kernel.cuh"

namespace slslam {
namespace {
struct RansacWs {
  int device = -1;
  char* d = nullptr; size_t d_cap = 0;
  char* h = nullptr; size_t h_cap = 0;
};
thread_local RansacWs g_rws;

int rws_ensure(int dev, size_t bytes) {
  RansacWs& w = g_rws;
  if (w.device != dev) {
    if (w.d) cudaFree(w.d);
    if (w.h) cudaFreeHost(w.h);
    w = RansacWs();
    w.device = dev;
  }
  if (bytes > w.d_cap) {
    if (w.d) cudaFree(w.d);
    if (w.h) cudaFreeHost(w.h);
    w.d = nullptr; w.h = nullptr; w.d_cap = w.h_cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    CUDA_TRY(cudaMalloc((void**)&w.d, want));
    CUDA_TRY(cudaMallocHost((void**)&w.h, want));
    w.d_cap = w.h_cap = want;
  }
  return SLSLAM_OK;
}
}  // namespace
}  // namespace slslam

using namespace slslam;

extern "C" int slslam_ransac_score(int32_t n_hyp, const double* poses, int32_t n_lines, const double* lines, const double* obs,
                                   double baseline, double thr, int32_t* scores, uint8_t* inlier, float* errors) {
  if (n_hyp < 0 || n_lines < 0 || !scores) return SLSLAM_ERR_INVALID;
  if (n_hyp > 0 && !poses) return SLSLAM_ERR_INVALID;
  if (n_lines > 0 && (!lines || !obs)) return SLSLAM_ERR_INVALID;
  if (n_hyp > 65535) return SLSLAM_ERR_UNSUPPORTED;
  int rc = ensure_device(-1);
  if (rc != SLSLAM_OK) return rc;
  if (n_hyp == 0) return SLSLAM_OK;
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t H = (size_t)n_hyp, K = (size_t)n_lines;
  size_t off = 0;
  auto reserve = [&](size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; };
  const size_t o_pose = reserve(96 * H), o_line = reserve(48 * K), o_obs = reserve(64 * K);
  const size_t upload = off;
  const size_t o_score = reserve(4 * H), o_in = reserve(inlier ? H * K : 0), o_err = reserve(errors ? 4 * H * K : 0);
  rc = rws_ensure(dev, off);
  if (rc != SLSLAM_OK) return rc;
  char* d = g_rws.d;
  char* h = g_rws.h;
  memcpy(h + o_pose, poses, 96 * H);
  if (K) { memcpy(h + o_line, lines, 48 * K); memcpy(h + o_obs, obs, 64 * K); }
  CUDA_TRY(cudaMemcpyAsync(d, h, upload, cudaMemcpyHostToDevice, nullptr));
  CUDA_TRY(cudaMemsetAsync(d + o_score, 0, 4 * H, nullptr));
  const dim3 grid((unsigned)std::max<size_t>(1, (K + 255) / 256), (unsigned)H);
  ransac_score_kernel<<<grid, 256>>>(n_hyp, (const double*)(d + o_pose), n_lines, (const double*)(d + o_line),
                                     (const double*)(d + o_obs), baseline, thr, (int*)(d + o_score),
                                     inlier ? (unsigned char*)(d + o_in) : nullptr, errors ? (float*)(d + o_err) : nullptr);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(h + o_score, d + o_score, off - o_score, cudaMemcpyDeviceToHost, nullptr));
  CUDA_TRY(cudaStreamSynchronize(nullptr));
  memcpy(scores, h + o_score, 4 * H);
  if (inlier) memcpy(inlier, h + o_in, H * K);
  if (errors) memcpy(errors, h + o_err, 4 * H * K);
  return SLSLAM_OK;
}
