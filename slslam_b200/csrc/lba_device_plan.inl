// Part of lba_host.cu (included there, inside namespace slslam): creation of a device-planned batch.
// Not a translation unit of its own: it uses the file-local state of lba_host.cu (Workspace, timings, HostPool).
// ---------------------------------------------------------------------------------------------------------------
// Device-planned batch: the caller's arrays are copied to the device as they are (through pinned staging, or straight
// from the caller's memory when that is already page-locked) and lba_plan_kernel builds the plan there.  The host does
// no per-observation work.  Returns SLSLAM_PLAN_FALLBACK when the host planner has to take over (a camera observing a
// line twice, a batch whose shared-memory shape needs the group-size search).
// ---------------------------------------------------------------------------------------------------------------
constexpr int SLSLAM_PLAN_FALLBACK = -1000;

static int validate_desc_light(const slslam_lba_desc& d) {
  if (d.num_cameras < 0 || d.num_lines < 0 || d.num_observations < 0 || d.max_iterations < 0) return SLSLAM_ERR_INVALID;
  if (d.num_observations > 0 && (!d.camera_index || !d.line_index || !d.fixed_index || !d.observations)) return SLSLAM_ERR_INVALID;
  if (d.num_cameras > MAX_CAMS) return SLSLAM_ERR_UNSUPPORTED;
  return SLSLAM_OK;
}

static bool is_page_locked(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// `device_inputs`: the arrays the descs point to, and `params`, are DEVICE memory (e.g. a buffer NCCL just received):
// nothing is staged or copied, the plan kernel reads the caller's arrays where they are, the solve kernel updates the
// parameters in place and writes the summaries to `summ_dev` when given.  Only the headers (< 1 KB per window) go up.
static int batch_create_device_plan(int32_t n, const slslam_lba_desc* descs, const double* const* params, int32_t device,
                                    int32_t cluster_size, Workspace* ws, cudaStream_t stream, slslam_lba_batch** out,
                                    bool device_inputs = false, slslam_summary* summ_dev = nullptr, bool deferred = false) {
  if (!out) return SLSLAM_ERR_INVALID;
  *out = nullptr;
  if (n <= 0 || !descs || !params) return SLSLAM_ERR_INVALID;
  for (int i = 0; i < n; ++i) {
    const int rc = validate_desc_light(descs[i]);
    if (rc != SLSLAM_OK) return rc;
    if (!params[i]) return SLSLAM_ERR_INVALID;
    if (device_inputs) continue;
    const int np = 6 * descs[i].num_cameras + 4 * descs[i].num_lines;
    for (int k = 0; k < np; ++k) if (!std::isfinite(params[i][k])) return SLSLAM_ERR_NUMERICAL;
  }
  const double t_begin = now_ms();
  int rc = ensure_device(device);
  if (rc != SLSLAM_OK) {
    // no device: argument errors still take precedence over the missing GPU (the index checks otherwise run on the device)
    if (!device_inputs) for (int i = 0; i < n; ++i) { const int v = validate_desc(descs[i]); if (v != SLSLAM_OK) return v; }
    return rc;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  long long max_obs = 0;
  int Lmax = 0, Cmax = 1;
  for (int i = 0; i < n; ++i) {
    max_obs = std::max<long long>(max_obs, descs[i].num_observations);
    Lmax = std::max(Lmax, descs[i].num_lines); Cmax = std::max(Cmax, descs[i].num_cameras);
  }
  const int cap = resident_ctas(dev);
  int smem_optin = 0;
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const int CS = pick_group_size(dev, n, max_obs, cluster_size, min_group_size_for_lines(Lmax, smem_optin));
  if (CS > cap) return SLSLAM_PLAN_FALLBACK;
  // the plan kernel keeps 13 + min(C, 24) bytes per line of a window in shared memory
  const size_t plan_smem = (13 + (size_t)std::min(Cmax, (int)MAX_FREE_CAMS)) * (size_t)Lmax + 16;
  if (plan_smem > (size_t)smem_optin - 4096) return SLSLAM_PLAN_FALLBACK;
  {
    static std::mutex attr_mutex;
    static bool attr_set[16] = {false};
    std::lock_guard<std::mutex> lk(attr_mutex);
    if (dev < 0 || dev >= 16 || !attr_set[dev]) {
      CUDA_TRY(cudaFuncSetAttribute(lba_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - 4096));
      if (dev >= 0 && dev < 16) attr_set[dev] = true;
    }
  }

  slslam_lba_batch* b = new (std::nothrow) slslam_lba_batch();
  if (!b) return SLSLAM_ERR_INVALID;
  b->device = dev; b->n = n; b->borrowed = ws != nullptr; b->ws = ws; b->device_planned = true; b->CS = CS;
  if (ws) b->plans.swap(ws->plans);
  b->plans.resize(n);
  b->dp.resize(n); b->dp_info.resize(n);

  // ---- pool layout: [uploaded: PlanIn | WinHdr | parameters | raw arrays] [device only: plan outputs, scratch, results, group scratch] ----
  size_t off = 0;
  auto reserve = [&](size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; };
  const size_t o_pin_in = reserve(sizeof(PlanIn) * n), o_hdr = reserve(sizeof(WinHdr) * n);
  b->param_off.resize(n); b->trace_off.resize(n); b->nparams.resize(n);
  size_t tp = 0, tt = 0;
  for (int i = 0; i < n; ++i) {
    b->nparams[i] = 6 * descs[i].num_cameras + 4 * descs[i].num_lines;
    b->param_off[i] = tp; tp += (size_t)((b->nparams[i] + 1) & ~1);
    b->trace_off[i] = tt; tt += (size_t)std::max(1, descs[i].max_iterations) * SLSLAM_TRACE_WIDTH;
  }
  b->total_params = tp; b->total_trace = tt;
  const size_t o_par = device_inputs ? off : reserve(tp * 8);
  std::vector<size_t> o_ci(n), o_li(n), o_fi(n), o_raw(n);
  std::vector<char> direct(n, 0);   // observations copied straight from page-locked caller memory
  if (!device_inputs) {
    for (int i = 0; i < n; ++i) {
      const size_t N = (size_t)descs[i].num_observations;
      o_ci[i] = reserve(4 * N); o_li[i] = reserve(4 * N); o_fi[i] = reserve(8 * N);
    }
    // observations last, so that the ones that are DMA'd directly leave no hole in the staged prefix
    // (page-locked index / flag arrays are staged all the same: 24 small DMAs ahead of the planner were slower than one
    // cache-hot staged copy, 1.08 against 1.01 ms per 8-window call)
    for (int i = 0; i < n; ++i) {
      const size_t N = (size_t)descs[i].num_observations;
      direct[i] = (N * 64 >= 65536 && is_page_locked(descs[i].observations)) ? 1 : 0;
    }
    for (int i = 0; i < n; ++i) if (!direct[i]) o_raw[i] = reserve(64 * (size_t)descs[i].num_observations);
  }
  const size_t upload = off;
  if (!device_inputs) for (int i = 0; i < n; ++i) if (direct[i]) o_raw[i] = reserve(64 * (size_t)descs[i].num_observations);
  struct Scratch { size_t cnt, start, fill, lconst, order, slotl; };
  std::vector<Scratch> sc(n);
  std::vector<size_t> o_vg(n), o_vr(n), o_sg(n), o_z(n);
  const int Cf_cap = std::min(Cmax, (int)MAX_FREE_CAMS);
  const int nkeys_cap = Cf_cap * (Cf_cap + 1) / 2, vpad_cap = (lba_vlen(Cf_cap) + 31) & ~31;
  // Z staging in global memory only when the Z rows of a CTA may not fit in shared memory
  bool want_zg = false;
  for (int i = 0; i < n; ++i) {
    const size_t N = (size_t)descs[i].num_observations, L = (size_t)descs[i].num_lines;
    auto& d = b->dp[i];
    d.slot_cap = (int)(2 * N + 32 * (size_t)MAX_G); d.item_cap = (int)(N * 31 / 2 + 1);
    d.obs = reserve((size_t)d.slot_cap * 64); d.meta = reserve((size_t)d.slot_cap * 8); d.gid = reserve(4 * L + 4);
    d.items = reserve((size_t)d.item_cap * 4 + 4); d.koff = reserve((size_t)MAX_G * (nkeys_cap + 1) * 4 + 4);
    d.hdr = o_hdr + sizeof(WinHdr) * i;
    Scratch& s = sc[i];
    s.cnt = reserve(4 * L + 4); s.start = reserve(4 * L + 8); s.fill = reserve(4 * L + 4); s.lconst = reserve(4 * L + 4);
    s.order = reserve(4 * N + 4);
    s.slotl = reserve((size_t)d.slot_cap * 4);
    const size_t slots_est = N / (size_t)CS * 5 / 4 + 96;
    if (49152 + L / (size_t)CS * 400 + slots_est * ZST * 8 > (size_t)smem_optin) want_zg = true;
  }
  const size_t o_info = reserve(sizeof(PlanInfo) * n);
  const size_t o_pout = device_inputs ? off : reserve(tp * 8);
  const size_t o_summ = reserve(sizeof(slslam_summary) * n), o_trace = reserve(tt * 8),
               o_phase = reserve(sizeof(long long) * NPHASE * n);
  for (int i = 0; i < n; ++i) {
    o_vg[i] = reserve((size_t)CS * vpad_cap * 8); o_vr[i] = reserve((size_t)vpad_cap * 8); o_sg[i] = reserve((size_t)CS * 8 * 8);
  }
  const size_t o_bar = reserve((size_t)n * 128);
  b->bar_bytes = (size_t)n * 128;
  for (int i = 0; i < n; ++i) o_z[i] = want_zg ? reserve((size_t)b->dp[i].slot_cap * ZST * 8) : 0;
  const size_t result_bytes = tp * 8 + sizeof(slslam_summary) * n + sizeof(PlanInfo) * n + 512;
  char* host = nullptr;
  std::vector<char> host_vec;
  PlanInfo* h_info = nullptr;
  if (ws) {
    rc = ws->ensure(dev, off, upload, result_bytes);
    if (rc != SLSLAM_OK) { slslam_lba_batch_destroy(b); return rc; }
    b->d_pool = ws->d_pool; host = ws->h_pin; b->h_params = (double*)ws->h_res;
    h_info = (PlanInfo*)(ws->h_res + ((tp * 8 + sizeof(slslam_summary) * n + 255) & ~(size_t)255));
  } else {
    CUDA_TRY_OR(cudaMalloc((void**)&b->d_pool, off), { delete b; return SLSLAM_ERR_CUDA; });
    host_vec.resize(upload);
    host = host_vec.data();
    CUDA_TRY_OR(cudaMallocHost((void**)&b->h_params, std::max<size_t>(tp, 1) * 8), { slslam_lba_batch_destroy(b); return SLSLAM_ERR_CUDA; });
    h_info = b->dp_info.data();
  }
  char* dp = b->d_pool;
  b->d_hdrs = (WinHdr*)(dp + o_hdr);
  b->d_params_in = (double*)(dp + o_par); b->d_params_out = (double*)(dp + o_pout);
  b->d_trace = (double*)(dp + o_trace); b->d_summ = (slslam_summary*)(dp + o_summ);
  b->d_phase = (long long*)(dp + o_phase); b->d_bar = (unsigned int*)(dp + o_bar);
  // observations that sit in page-locked caller memory are DMA'd straight from there, FIRST: those copies (most of the
  // bytes) then run while this thread stages the small arrays below
  // -- on the workspace's copy stream, so that they also overlap the planner (which reads the index arrays only; the
  // kernel that gathers the observations into slot order waits for them)
  cudaError_t e = cudaSuccess;
  if (ws) cudaEventRecord(ws->ev[0], stream);
  size_t direct_bytes = 0;
  cudaStream_t obs_stream = (ws && ws->copy_stream) ? ws->copy_stream : stream;
  if (obs_stream != stream) { cudaEventRecord(ws->ev_obs, stream); cudaStreamWaitEvent(obs_stream, ws->ev_obs, 0); }   // the pool may still be in use on `stream`
  for (int i = 0; i < n && e == cudaSuccess; ++i) {
    if (!direct[i]) continue;
    const size_t bytes = 64 * (size_t)descs[i].num_observations;
    e = cudaMemcpyAsync(dp + o_raw[i], descs[i].observations, bytes, cudaMemcpyHostToDevice, obs_stream);
    direct_bytes += bytes;
  }
  const bool obs_on_copy_stream = obs_stream != stream && direct_bytes > 0;
  if (obs_on_copy_stream && e == cudaSuccess) e = cudaEventRecord(ws->ev_obs, obs_stream);
  if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); cudaGetLastError(); slslam_lba_batch_destroy(b); return SLSLAM_ERR_CUDA; }
  // ---- staging: headers and the caller's arrays as they are (one thread, one buffer, one copy) ----
  for (int i = 0; i < n; ++i) {
    const slslam_lba_desc& d = descs[i];
    const size_t N = (size_t)d.num_observations;
    WindowPlan& wp = b->plans[i];
    wp.C = d.num_cameras; wp.L = d.num_lines; wp.N = d.num_observations; wp.max_iters = d.max_iterations; wp.robust = d.robust ? 1 : 0;
    wp.huber_a = d.huber_delta > 0 ? d.huber_delta : 1.0 / 406.05;
    wp.baseline = d.baseline >= 0 ? d.baseline : 0.12;
    wp.ftol = d.function_tolerance > 0 ? d.function_tolerance : 1e-6;
    wp.gtol = d.gradient_tolerance > 0 ? d.gradient_tolerance : 1e-10;
    wp.ptol = d.parameter_tolerance > 0 ? d.parameter_tolerance : 1e-8;
    wp.radius0 = d.initial_trust_region_radius > 0 ? d.initial_trust_region_radius : 1e4;
    const auto& q = b->dp[i];
    WinHdr h; memset(&h, 0, sizeof(h));
    h.C = wp.C; h.L = wp.L; h.max_iters = wp.max_iters; h.robust = wp.robust;
    h.huber_a = wp.huber_a; h.baseline = wp.baseline; h.ftol = wp.ftol; h.gtol = wp.gtol; h.ptol = wp.ptol; h.radius0 = wp.radius0;
    h.obs = (const double*)(dp + q.obs); h.meta = (const int2*)(dp + q.meta); h.line_gid = (const int*)(dp + q.gid);
    h.items = (const uint32_t*)(dp + q.items); h.key_off = (const int*)(dp + q.koff);
    h.params_in = b->d_params_in + b->param_off[i]; h.params_out = b->d_params_out + b->param_off[i];
    if (device_inputs) { h.params_in = params[i]; h.params_out = const_cast<double*>(params[i]); }   // in place: every read of a block precedes the first group barrier, every write follows the last
    h.Zg = want_zg ? (double*)(dp + o_z[i]) : nullptr;
    h.Vg = (double*)(dp + o_vg[i]); h.Vr = (double*)(dp + o_vr[i]); h.scalg = (double*)(dp + o_sg[i]);
    h.bar = (unsigned int*)(dp + o_bar + (size_t)i * 128);
    h.summary = summ_dev ? summ_dev + i : b->d_summ + i;
    h.trace = ws ? nullptr : b->d_trace + b->trace_off[i];
    h.phase_cycles = ws ? nullptr : b->d_phase + (size_t)NPHASE * i;
    memcpy(host + o_hdr + sizeof(WinHdr) * i, &h, sizeof(h));
    const Scratch& s = sc[i];
    PlanIn pi; memset(&pi, 0, sizeof(pi));
    pi.C = wp.C; pi.L = wp.L; pi.N = wp.N; pi.CS = CS; pi.slot_cap = q.slot_cap; pi.item_cap = q.item_cap;
    pi.cam_idx = (const int*)(dp + o_ci[i]); pi.line_idx = (const int*)(dp + o_li[i]); pi.fixed = (const int*)(dp + o_fi[i]);
    pi.obs_raw = (const double*)(dp + o_raw[i]);
    if (device_inputs) { pi.cam_idx = d.camera_index; pi.line_idx = d.line_index; pi.fixed = d.fixed_index; pi.obs_raw = d.observations; }
    pi.obs = (double*)(dp + q.obs); pi.meta = (int2*)(dp + q.meta); pi.line_gid = (int*)(dp + q.gid);
    pi.items = (uint32_t*)(dp + q.items); pi.key_off = (int*)(dp + q.koff);
    pi.hdr = (WinHdr*)(dp + q.hdr); pi.info = (PlanInfo*)(dp + o_info) + i;
    pi.line_cnt = (int*)(dp + s.cnt); pi.line_start = (int*)(dp + s.start); pi.fill = (int*)(dp + s.fill); pi.lconst = (int*)(dp + s.lconst);
    pi.order = (int*)(dp + s.order);
    pi.slot_line = (int*)(dp + s.slotl);
    memcpy(host + o_pin_in + sizeof(PlanIn) * i, &pi, sizeof(pi));
    if (device_inputs) continue;
    memcpy(host + o_par + b->param_off[i] * 8, params[i], (size_t)b->nparams[i] * 8);
    memcpy(host + o_ci[i], d.camera_index, 4 * N);
    memcpy(host + o_li[i], d.line_index, 4 * N);
    memcpy(host + o_fi[i], d.fixed_index, 8 * N);
  }
  b->inplace = device_inputs;
  b->upload_bytes = upload + direct_bytes;
  // The staged prefix (headers, parameters, index and flag arrays of every window) goes up first and the planner, which
  // reads nothing else, starts behind it.  Pageable observations follow in pieces of <= 1 MB: piece k is copied into the
  // page-locked staging buffer while the DMA of piece k-1 runs (on the copy stream when there is one, so the pieces also
  // overlap the planner), and a piece that has just been written is still in the cache the DMA reads from (measured on the
  // bench host: 49 GB/s for a freshly written buffer, 13-28 GB/s otherwise, profiles/r2_h2d_staging.txt).
  size_t prefix = upload;
  if (!device_inputs) for (int i = 0; i < n; ++i) if (!direct[i] && descs[i].num_observations > 0) prefix = std::min(prefix, o_raw[i]);
  e = cudaMemcpyAsync(dp, host, prefix, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) {
    lba_plan_kernel<<<n, PLAN_NT, plan_smem, stream>>>((const PlanIn*)(dp + o_pin_in));
    e = cudaGetLastError();
  }
  bool staged_obs = false;
  if (!device_inputs) {
    constexpr size_t PIECE = (size_t)1 << 20;
    for (int i = 0; i < n && e == cudaSuccess; ++i) {
      if (direct[i]) continue;
      const size_t bytes = 64 * (size_t)descs[i].num_observations;
      const char* src = reinterpret_cast<const char*>(descs[i].observations);
      for (size_t o = 0; o < bytes && e == cudaSuccess; o += PIECE) {
        const size_t m = std::min(PIECE, bytes - o);
        memcpy(host + o_raw[i] + o, src + o, m);
        e = cudaMemcpyAsync(dp + o_raw[i] + o, host + o_raw[i] + o, m, cudaMemcpyHostToDevice, obs_stream);
        staged_obs = true;
      }
    }
    if (e == cudaSuccess && staged_obs && obs_stream != stream) e = cudaEventRecord(ws->ev_obs, obs_stream);
  }
  const double t_staged = now_ms();
  if (e == cudaSuccess) {
    if ((obs_on_copy_stream || (staged_obs && obs_stream != stream))) e = cudaStreamWaitEvent(stream, ws->ev_obs, 0);
    if (e == cudaSuccess) {
      const int gx = (int)std::max<long long>(1, std::min<long long>(32, (max_obs * 4 + 2047) / 2048));
      lba_gather_obs_kernel<<<dim3((unsigned)gx, (unsigned)n), 256, 0, stream>>>((const PlanIn*)(dp + o_pin_in));
      e = cudaGetLastError();
    }
  }
  b->d_info = (PlanInfo*)(dp + o_info); b->Cmax = Cmax; b->smem_optin = smem_optin; b->want_zg = want_zg;
  if (deferred && e == cudaSuccess) {
    // `deferred`: nothing is read back here.  The solve kernel is launched right behind the planner with the maximal
    // shared-memory size and derives every window's layout itself; the planner's flags are looked at after the solve
    // (device_plan_check below), together with the results -- one synchronisation per call instead of two.
    b->deferred = true;
    b->lay = SmemLayout(); memset(&b->lay, 0, sizeof(b->lay));
    b->lay.G = CS; b->lay.smem_limit = smem_optin; b->lay.total = 0;
    b->smem_bytes = (size_t)smem_optin;
    for (int i = 0; i < n; ++i) b->plans[i].has_unobserved_blocks = true;      // unknown yet: keep the pre-copy of the parameters
    rc = set_solve_kernel_smem_limit(dev, smem_optin);
    if (rc != SLSLAM_OK) { slslam_lba_batch_destroy(b); return rc; }
    b->max_active = balanced_wave(n, cap / CS);
    if (ws) { g_timing[0] = now_ms() - t_staged; g_timing[1] = t_staged - t_begin; }
    *out = b;
    return SLSLAM_OK;
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(h_info, dp + o_info, sizeof(PlanInfo) * n, cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); cudaGetLastError(); slslam_lba_batch_destroy(b); return SLSLAM_ERR_CUDA; }
  int Cfmax = 0, mlines = 1, mslots = 32, mitems = 0, flags = 0;
  for (int i = 0; i < n; ++i) {
    const PlanInfo& pi = h_info[i];
    b->dp_info[i] = pi;
    flags |= pi.error;
    Cfmax = std::max(Cfmax, pi.Cf); mlines = std::max(mlines, pi.max_lines_cta); mslots = std::max(mslots, pi.max_slots_cta);
    mitems = std::max(mitems, pi.max_items_cta);
    b->plans[i].Cf = pi.Cf; b->plans[i].nkeys = pi.Cf * (pi.Cf + 1) / 2; b->plans[i].has_unobserved_blocks = pi.has_unobserved != 0;
    b->plans[i].max_lines_cta = pi.max_lines_cta; b->plans[i].max_slots_cta = pi.max_slots_cta; b->plans[i].max_items_cta = pi.max_items_cta;
  }
  if (flags & PLAN_ERR_INDEX) { slslam_lba_batch_destroy(b); return SLSLAM_ERR_INVALID; }
  if (flags & (PLAN_DUPLICATE_CAMERA | PLAN_ERR_CAPACITY)) { slslam_lba_batch_destroy(b); return SLSLAM_PLAN_FALLBACK; }
  if (flags & PLAN_ERR_LIMIT) {
    // too many free cameras / observations per line are final; "too many slots per CTA" may go away with a larger group
    slslam_lba_batch_destroy(b);
    return SLSLAM_PLAN_FALLBACK;
  }
  b->lay = lba_layout(Cmax, Cfmax, mlines, mslots, CS, (size_t)smem_optin, mitems);
  b->smem_bytes = (size_t)b->lay.total * 8;
  if (b->smem_bytes > (size_t)smem_optin || (!b->lay.z_in_smem && !want_zg)) { slslam_lba_batch_destroy(b); return SLSLAM_PLAN_FALLBACK; }
  rc = set_solve_kernel_smem_limit(dev, smem_optin);
  if (rc != SLSLAM_OK) { slslam_lba_batch_destroy(b); return rc; }
  b->max_active = balanced_wave(n, cap / CS);
  // "plan" here = H2D + device plan kernel + read-back of the sizes (the host waits for it); "stage" = the pinned staging
  if (ws) { g_timing[0] = now_ms() - t_staged; g_timing[1] = t_staged - t_begin; }
  *out = b;
  return SLSLAM_OK;
}



// After the solve of a `deferred` batch: the planner's records (already copied to `info`, host memory) say whether every
// window was planned and fitted -- the same checks batch_create_device_plan makes before the launch otherwise.
static int device_plan_check_deferred(const slslam_lba_batch* b, const PlanInfo* info) {
  int flags = 0;
  for (int i = 0; i < b->n; ++i) flags |= info[i].error;
  if (flags & PLAN_ERR_INDEX) return SLSLAM_ERR_INVALID;
  if (flags) return SLSLAM_PLAN_FALLBACK;
  for (int i = 0; i < b->n; ++i) {
    const SmemLayout l = lba_layout(b->plans[i].C, info[i].Cf, info[i].max_lines_cta, info[i].max_slots_cta, b->CS, (size_t)b->smem_optin,
                                    info[i].max_items_cta);
    if ((size_t)l.total * 8 > (size_t)b->smem_optin || (!l.z_in_smem && !b->want_zg)) return SLSLAM_PLAN_FALLBACK;
  }
  return SLSLAM_OK;
}
