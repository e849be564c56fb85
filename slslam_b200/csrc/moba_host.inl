// Part of lba_host.cu (included there): host side of the motion-only BA fast path (moba_kernel.cuh).
// Not a translation unit of its own: it uses the calling thread's cached Workspace and timing record of lba_host.cu.
namespace slslam {

// Motion-only BA (reference src/slam.cpp:578-675): exactly one used, non-constant camera and every observed line
// constant (constants are sticky per block, reference src/lba_problem.cpp:88-91).  Such a window needs no plan at all:
// the caller's arrays go to the device as they are and lba_motion_only_kernel (moba_kernel.cuh) does the rest.
static bool moba_candidate(const slslam_lba_desc& d, int* free_cam, int* nfree) {
  const int C = d.num_cameras, L = d.num_lines, N = d.num_observations;
  if (N <= 0 || C <= 0 || L <= 0 || C > MAX_CAMS) return false;
  char cam_used[MAX_CAMS] = {0}, cam_const[MAX_CAMS] = {0};
  int cnt[MAX_CAMS] = {0};
  std::vector<char> line_const((size_t)L, 0);
  for (int i = 0; i < N; ++i) {
    const int c = d.camera_index[i];
    cam_used[c] = 1; ++cnt[c];
    if (d.fixed_index[2 * i]) cam_const[c] = 1;
    if (d.fixed_index[2 * i + 1]) line_const[d.line_index[i]] = 1;
  }
  for (int i = 0; i < N; ++i) if (!line_const[d.line_index[i]]) return false;
  int fc = -1;
  for (int c = 0; c < C; ++c) {
    if (cam_used[c] && !cam_const[c]) { if (fc >= 0) return false; fc = c; }
  }
  if (fc < 0 || cnt[fc] > MOBA_MAX_FREE_OBS) return false;
  *free_cam = fc; *nfree = cnt[fc];
  return true;
}

// n motion-only problems: one staging buffer, one H2D copy, one launch (a CTA per problem), one D2H copy.
static int moba_solve_batch(int n, const slslam_lba_desc* descs, double* const* params_inout, slslam_summary* summaries_out,
                            const int* free_cam, const int* nfree) {
  const double t0 = now_ms();
  int rc = ensure_device(-1);
  if (rc != SLSLAM_OK) return rc;
  int device = 0;
  cudaGetDevice(&device);
  size_t off = 0;
  auto reserve = [&](size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; };
  const size_t o_hdr = reserve(sizeof(MobaHdr) * n);
  std::vector<size_t> o_ci(n), o_li(n), o_ob(n), o_p(n), o_po(n), np(n);
  int max_free = 0;
  for (int i = 0; i < n; ++i) {
    const size_t N = (size_t)descs[i].num_observations;
    np[i] = (size_t)6 * descs[i].num_cameras + (size_t)4 * descs[i].num_lines;
    o_ci[i] = reserve(4 * N); o_li[i] = reserve(4 * N); o_ob[i] = reserve(64 * N); o_p[i] = reserve(8 * np[i]);
    max_free = std::max(max_free, nfree[i]);
  }
  const size_t upload = off;
  size_t res = 0;
  for (int i = 0; i < n; ++i) { o_po[i] = res; res += (np[i] + 1) & ~(size_t)1; }
  const size_t o_pout = reserve(res * 8), o_summ = reserve(sizeof(slslam_summary) * n);
  const size_t result_bytes = res * 8 + sizeof(slslam_summary) * n + 256;
  rc = g_ws.ensure(device, off, upload, result_bytes);
  if (rc != SLSLAM_OK) return rc;
  char* host = g_ws.h_pin;
  char* dev = g_ws.d_pool;
  for (int i = 0; i < n; ++i) {
    const slslam_lba_desc& d = descs[i];
    const size_t N = (size_t)d.num_observations;
    MobaHdr h; memset(&h, 0, sizeof(h));
    h.C = d.num_cameras; h.L = d.num_lines; h.N = d.num_observations; h.free_cam = free_cam[i];
    h.max_iters = d.max_iterations; h.robust = d.robust ? 1 : 0;
    h.huber_a = d.huber_delta > 0 ? d.huber_delta : 1.0 / 406.05;
    h.baseline = d.baseline >= 0 ? d.baseline : 0.12;
    h.ftol = d.function_tolerance > 0 ? d.function_tolerance : 1e-6;
    h.gtol = d.gradient_tolerance > 0 ? d.gradient_tolerance : 1e-10;
    h.ptol = d.parameter_tolerance > 0 ? d.parameter_tolerance : 1e-8;
    h.radius0 = d.initial_trust_region_radius > 0 ? d.initial_trust_region_radius : 1e4;
    h.cam_idx = (const int*)(dev + o_ci[i]); h.line_idx = (const int*)(dev + o_li[i]);
    h.obs = (const double*)(dev + o_ob[i]); h.params_in = (const double*)(dev + o_p[i]);
    h.params_out = (double*)(dev + o_pout) + o_po[i];
    h.summary = (slslam_summary*)(dev + o_summ) + i;
    h.trace = nullptr;
    memcpy(host + o_hdr + sizeof(MobaHdr) * i, &h, sizeof(h));
    memcpy(host + o_ci[i], d.camera_index, 4 * N);
    memcpy(host + o_li[i], d.line_index, 4 * N);
    memcpy(host + o_ob[i], d.observations, 64 * N);
    memcpy(host + o_p[i], params_inout[i], 8 * np[i]);
  }
  const double t1 = now_ms();
  const size_t smem = ((size_t)MOBA_FIXED_DOUBLES + (size_t)MOBA_OBS_STRIDE * std::max(max_free, 1)) * 8;
  {
    static std::mutex attr_mutex;
    static bool attr_set[16] = {false};
    std::lock_guard<std::mutex> lk(attr_mutex);
    if (device < 0 || device >= 16 || !attr_set[device]) {
      int smem_optin = 0;
      cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
      CUDA_TRY(cudaFuncSetAttribute(lba_motion_only_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));
      if (device >= 0 && device < 16) attr_set[device] = true;
    }
  }
  CUDA_TRY(cudaEventRecord(g_ws.ev[0], nullptr));
  CUDA_TRY(cudaMemcpyAsync(dev, host, upload, cudaMemcpyHostToDevice, nullptr));
  CUDA_TRY(cudaEventRecord(g_ws.ev[1], nullptr));
  lba_motion_only_kernel<<<n, MOBA_NT, smem>>>((const MobaHdr*)(dev + o_hdr));
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(g_ws.ev[2], nullptr));
  double* h_par = (double*)g_ws.h_res;
  slslam_summary* h_summ = (slslam_summary*)(h_par + res);
  CUDA_TRY(cudaMemcpyAsync(h_par, dev + o_pout, res * 8, cudaMemcpyDeviceToHost, nullptr));
  CUDA_TRY(cudaMemcpyAsync(h_summ, dev + o_summ, sizeof(slslam_summary) * n, cudaMemcpyDeviceToHost, nullptr));
  CUDA_TRY(cudaEventRecord(g_ws.ev[3], nullptr));
  CUDA_TRY(cudaStreamSynchronize(nullptr));
  const double t2 = now_ms();
  for (int i = 0; i < n; ++i) {
    memcpy(params_inout[i], h_par + o_po[i], np[i] * 8);
    if (summaries_out) summaries_out[i] = h_summ[i];
  }
  const double t3 = now_ms();
  float ms = 0.f;
  for (int k = 0; k < 3; ++k) { cudaEventElapsedTime(&ms, g_ws.ev[k], g_ws.ev[k + 1]); g_timing[5 + k] = ms; }
  g_timing[0] = 0.0; g_timing[1] = t1 - t0; g_timing[2] = t2 - t1; g_timing[3] = t3 - t2; g_timing[4] = t3 - t0;
  return SLSLAM_OK;
}

}  // namespace slslam

