// Part of lba_host.cu (included there): host side of the general LBA kernel (wide_kernel.cuh) for windows beyond the
// tiled kernel's limits -- more than MAX_CAMS camera blocks, more than MAX_FREE_CAMS free cameras, or a line observed
// more than 32 times: the shapes the reference's --ba_window_size 20 / 40 runs produce (src/slam.cpp:1376-1382).
// Not a translation unit of its own: it uses the calling thread's cached Workspace of lba_host.cu.
namespace slslam {

struct WidePlan {
  int C = 0, L = 0, N = 0, Cf = 0;
  std::vector<int> cam_free, line_free, line_start, order, cam_start, cam_obs, pos;
};

// Sticky constants, reduced indices, line-sorted order, per-camera lists, (line, free camera) -> observation table.
// 0 = ok, SLSLAM_ERR_* otherwise.  `needs_wide` is set when the tiled kernel cannot take the window.
static int wide_plan(const slslam_lba_desc& d, WidePlan& p, bool* needs_wide) {
  const int C = d.num_cameras, L = d.num_lines, N = d.num_observations;
  p.C = C; p.L = L; p.N = N;
  std::vector<char> cam_used(C, 0), cam_const(C, 0), line_const(L, 0);
  std::vector<int> cnt(L, 0);
  for (int i = 0; i < N; ++i) {
    const int c = d.camera_index[i], l = d.line_index[i];
    if (c < 0 || c >= C || l < 0 || l >= L) return SLSLAM_ERR_INVALID;
    cam_used[c] = 1; ++cnt[l];
    if (d.fixed_index[2 * i]) cam_const[c] = 1;
    if (d.fixed_index[2 * i + 1]) line_const[l] = 1;
  }
  p.cam_free.assign(C, -1); p.Cf = 0;
  for (int c = 0; c < C; ++c) if (cam_used[c] && !cam_const[c]) p.cam_free[c] = p.Cf++;
  p.line_free.assign(L, 0);
  int maxcnt = 0;
  for (int l = 0; l < L; ++l) { p.line_free[l] = (cnt[l] > 0 && !line_const[l]) ? 1 : 0; maxcnt = std::max(maxcnt, cnt[l]); }
  if (needs_wide) *needs_wide = C > MAX_CAMS || p.Cf > MAX_FREE_CAMS || maxcnt > 32;
  p.line_start.assign(L + 1, 0);
  for (int l = 0; l < L; ++l) p.line_start[l + 1] = p.line_start[l] + cnt[l];
  std::vector<int> fill(p.line_start.begin(), p.line_start.end() - 1);
  p.order.resize(N);
  for (int i = 0; i < N; ++i) p.order[fill[d.line_index[i]]++] = i;          // stable: the caller's order inside a line
  return SLSLAM_OK;
}

static int wide_plan_tables(const slslam_lba_desc& d, WidePlan& p) {
  const int N = p.N, L = p.L, Cf = p.Cf;
  if (Cf > WIDE_MAX_FREE) { set_last_error("more free cameras than the general kernel takes (slslam_lba_get_limits: max_free_cameras_general)"); return SLSLAM_ERR_UNSUPPORTED; }
  p.cam_start.assign(Cf + 1, 0);
  std::vector<int> ccount(Cf, 0);
  for (int s = 0; s < N; ++s) { const int f = p.cam_free[d.camera_index[p.order[s]]]; if (f >= 0) ++ccount[f]; }
  for (int f = 0; f < Cf; ++f) p.cam_start[f + 1] = p.cam_start[f] + ccount[f];
  p.cam_obs.assign(p.cam_start[Cf], 0);
  std::vector<int> cfill(p.cam_start.begin(), p.cam_start.end() - 1);
  p.pos.assign((size_t)L * std::max(Cf, 1), -1);
  for (int s = 0; s < N; ++s) {
    const int i = p.order[s], f = p.cam_free[d.camera_index[i]], l = d.line_index[i];
    if (f < 0) continue;
    p.cam_obs[cfill[f]++] = s;
    if (!p.line_free[l]) continue;
    int& slot = p.pos[(size_t)l * Cf + f];
    if (slot >= 0) { set_last_error("a camera observes one line twice: not supported by the general kernel"); return SLSLAM_ERR_UNSUPPORTED; }
    slot = s;
  }
  return SLSLAM_OK;
}

static int wide_solve_batch(int n, const slslam_lba_desc* descs, double* const* params_inout, slslam_summary* summaries_out,
                            std::vector<WidePlan>& plans) {
  const double t0 = now_ms();
  int rc = ensure_device(-1);
  if (rc != SLSLAM_OK) return rc;
  int device = 0;
  cudaGetDevice(&device);
  for (int i = 0; i < n; ++i) { rc = wide_plan_tables(descs[i], plans[i]); if (rc != SLSLAM_OK) return rc; }
  size_t off = 0;
  auto reserve = [&](size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; };
  struct Off { size_t cam_s, line_s, obs_s, lstart, cfree, lfree, cstart, cobs, pos, par; };
  std::vector<Off> o(n);
  const size_t o_hdr = reserve(sizeof(WideHdr) * n);
  std::vector<size_t> np(n), o_po(n);
  for (int i = 0; i < n; ++i) {
    const WidePlan& p = plans[i];
    const size_t N = (size_t)p.N, L = (size_t)p.L, C = (size_t)p.C, Cf = (size_t)p.Cf;
    np[i] = 6 * C + 4 * L;
    o[i].cam_s = reserve(4 * N); o[i].line_s = reserve(4 * N); o[i].obs_s = reserve(64 * N); o[i].lstart = reserve(4 * (L + 1));
    o[i].cfree = reserve(4 * C + 4); o[i].lfree = reserve(4 * L + 4); o[i].cstart = reserve(4 * (Cf + 1)); o[i].cobs = reserve(4 * p.cam_obs.size() + 4);
    o[i].pos = reserve(4 * p.pos.size() + 4); o[i].par = reserve(8 * np[i] + 8);
  }
  const size_t upload = off;
  size_t res = 0;
  for (int i = 0; i < n; ++i) { o_po[i] = res; res += (np[i] + 1) & ~(size_t)1; }
  const size_t o_pout = reserve(res * 8 + 8), o_summ = reserve(sizeof(slslam_summary) * n);
  struct Scr { size_t P, camx, camxt, camR, camRt, linex, linext, ltrig, ltrigt, cscale, lscale, r, Jc, Jl, Z, lineLU, S, gc, zu, hd, yc, ub, ab, part, flag; };
  std::vector<Scr> sc(n);
  for (int i = 0; i < n; ++i) {
    const WidePlan& p = plans[i];
    const size_t N = (size_t)p.N, L = (size_t)p.L, C = (size_t)p.C, nn = 6 * (size_t)p.Cf;
    Scr& s = sc[i];
    s.camx = reserve(48 * C + 8); s.camxt = reserve(48 * C + 8); s.camR = reserve(8 * CAM_STRIDE * C + 8); s.camRt = reserve(8 * CAM_STRIDE * C + 8);
    s.linex = reserve(32 * L + 8); s.linext = reserve(32 * L + 8); s.cscale = reserve(8 * nn + 8); s.lscale = reserve(32 * L + 8);
    s.r = reserve(32 * N + 8); s.Jc = reserve(192 * N + 8); s.Jl = reserve(128 * N + 8); s.Z = reserve(192 * N + 8); s.lineLU = reserve(176 * L + 8);
    s.S = reserve(8 * nn * nn + 8); s.P = reserve(8 * nn * nn + 8); s.gc = reserve(8 * nn + 8); s.zu = reserve(8 * nn + 8); s.hd = reserve(8 * nn + 8);
    s.yc = reserve(8 * nn + 8); s.ub = reserve(8 * nn + 8); s.ab = reserve(8 * nn + 8);
    s.ltrig = reserve(64 * L + 8); s.ltrigt = reserve(64 * L + 8);
    s.part = reserve(8 * 2 * WIDE_MAX_G * WIDE_NPART); s.flag = reserve(8);
  }
  const size_t o_bar = reserve((size_t)n * 128);          // one barrier counter per window, on its own line
  const size_t result_bytes = res * 8 + sizeof(slslam_summary) * n + 256;
  rc = g_ws.ensure(device, off, upload, result_bytes);
  if (rc != SLSLAM_OK) return rc;
  char* host = g_ws.h_pin;
  char* dev = g_ws.d_pool;
  for (int i = 0; i < n; ++i) {
    const slslam_lba_desc& d = descs[i];
    const WidePlan& p = plans[i];
    WideHdr h; memset(&h, 0, sizeof(h));
    h.C = p.C; h.Cf = p.Cf; h.L = p.L; h.N = p.N; h.n = 6 * p.Cf; h.max_iters = d.max_iterations; h.robust = d.robust ? 1 : 0;
    h.huber_a = d.huber_delta > 0 ? d.huber_delta : 1.0 / 406.05;
    h.baseline = d.baseline >= 0 ? d.baseline : 0.12;
    h.ftol = d.function_tolerance > 0 ? d.function_tolerance : 1e-6;
    h.gtol = d.gradient_tolerance > 0 ? d.gradient_tolerance : 1e-10;
    h.ptol = d.parameter_tolerance > 0 ? d.parameter_tolerance : 1e-8;
    h.radius0 = d.initial_trust_region_radius > 0 ? d.initial_trust_region_radius : 1e4;
    h.cam_s = (const int*)(dev + o[i].cam_s); h.line_s = (const int*)(dev + o[i].line_s); h.obs_s = (const double*)(dev + o[i].obs_s);
    h.line_start = (const int*)(dev + o[i].lstart); h.cam_free = (const int*)(dev + o[i].cfree); h.line_free = (const int*)(dev + o[i].lfree);
    h.cam_start = (const int*)(dev + o[i].cstart); h.cam_obs = (const int*)(dev + o[i].cobs); h.pos = (const int*)(dev + o[i].pos);
    h.params_in = (const double*)(dev + o[i].par); h.params_out = (double*)(dev + o_pout) + o_po[i];
    h.summary = (slslam_summary*)(dev + o_summ) + i; h.trace = nullptr;
    const Scr& s = sc[i];
    h.camx[0] = (double*)(dev + s.camx); h.camx[1] = (double*)(dev + s.camxt); h.camR[0] = (double*)(dev + s.camR); h.camR[1] = (double*)(dev + s.camRt);
    h.linex[0] = (double*)(dev + s.linex); h.linex[1] = (double*)(dev + s.linext); h.ltrig[0] = (double*)(dev + s.ltrig); h.ltrig[1] = (double*)(dev + s.ltrigt);
    h.cscale = (double*)(dev + s.cscale); h.lscale = (double*)(dev + s.lscale);
    h.part = (double*)(dev + s.part); h.flag = (double*)(dev + s.flag); h.bar = (unsigned int*)(dev + o_bar + (size_t)i * 128);
    h.r = (double*)(dev + s.r); h.Jc = (double*)(dev + s.Jc); h.Jl = (double*)(dev + s.Jl); h.Z = (double*)(dev + s.Z); h.lineLU = (double*)(dev + s.lineLU);
    h.S = (double*)(dev + s.S); h.P = (double*)(dev + s.P); h.gc = (double*)(dev + s.gc); h.zu = (double*)(dev + s.zu); h.hd = (double*)(dev + s.hd);
    h.yc = (double*)(dev + s.yc); h.ub = (double*)(dev + s.ub); h.ab = (double*)(dev + s.ab);
    memcpy(host + o_hdr + sizeof(WideHdr) * i, &h, sizeof(h));
    int* cs = (int*)(host + o[i].cam_s); int* ls = (int*)(host + o[i].line_s); double* os = (double*)(host + o[i].obs_s);
    for (int k = 0; k < p.N; ++k) {
      const int src = p.order[k];
      cs[k] = d.camera_index[src]; ls[k] = d.line_index[src];
      memcpy(os + 8 * (size_t)k, d.observations + 8 * (size_t)src, 64);
    }
    memcpy(host + o[i].lstart, p.line_start.data(), 4 * p.line_start.size());
    if (p.C) memcpy(host + o[i].cfree, p.cam_free.data(), 4 * (size_t)p.C);
    if (p.L) memcpy(host + o[i].lfree, p.line_free.data(), 4 * (size_t)p.L);
    memcpy(host + o[i].cstart, p.cam_start.data(), 4 * p.cam_start.size());
    if (!p.cam_obs.empty()) memcpy(host + o[i].cobs, p.cam_obs.data(), 4 * p.cam_obs.size());
    if (!p.pos.empty()) memcpy(host + o[i].pos, p.pos.data(), 4 * p.pos.size());
    memcpy(host + o[i].par, params_inout[i], 8 * np[i]);
  }
  const double t1 = now_ms();
  const size_t smem = (size_t)(WIDE_NPART * WIDE_NT + 36 + 16 + (WIDE_MAX_FREE * (WIDE_MAX_FREE + 1) / 2 + 1) / 2 +
                               WIDE_TAIL_MAX * (WIDE_TAIL_MAX + 1) / 2 * 36 + (WIDE_TAIL_MAX - 1) * 36 + 12 * WIDE_TAIL_MAX) * 8;
  {
    static bool attr_set[16] = {false};
    if (device >= 16 || !attr_set[device]) {
      CUDA_TRY(cudaFuncSetAttribute(lba_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      if (device < 16) attr_set[device] = true;
    }
  }
  // group size: as many CTAs per window as stay co-resident with every window of the call (cooperative launch: the
  // group barrier spins), at most WIDE_MAX_G; more windows than SMs run in waves of one CTA each
  static int wide_cap[16] = {0};
  if (device >= 16 || wide_cap[device] == 0) {
    int per_sm = 0, sms = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lba_wide_kernel, WIDE_NT, smem));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (per_sm < 1) { set_last_error("the general LBA kernel does not fit on this device"); return SLSLAM_ERR_UNSUPPORTED; }
    if (device < 16) wide_cap[device] = per_sm * sms;
    else wide_cap[0] = per_sm * sms;
  }
  const int cap = wide_cap[device < 16 ? device : 0];
  const int G = std::max(1, std::min((int)WIDE_MAX_G, cap / std::max(1, n)));
  const int groups = std::min(n, cap / G);
  CUDA_TRY(cudaEventRecord(g_ws.ev[0], nullptr));
  CUDA_TRY(cudaMemcpyAsync(dev, host, upload, cudaMemcpyHostToDevice, nullptr));
  CUDA_TRY(cudaMemsetAsync(dev + o_bar, 0, (size_t)n * 128, nullptr));
  CUDA_TRY(cudaEventRecord(g_ws.ev[1], nullptr));
  {
    const WideHdr* d_hdrs = (const WideHdr*)(dev + o_hdr);
    int nwin = n, gsz = G;
    void* args[] = {(void*)&d_hdrs, (void*)&nwin, (void*)&gsz};
    CUDA_TRY(cudaLaunchCooperativeKernel((const void*)lba_wide_kernel, dim3((unsigned)(groups * G)), dim3(WIDE_NT), args, smem, nullptr));
  }
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
  g_timing[0] = t1 - t0; g_timing[1] = 0.0; g_timing[2] = t2 - t1; g_timing[3] = t3 - t2; g_timing[4] = t3 - t0;
  return SLSLAM_OK;
}

}  // namespace slslam
