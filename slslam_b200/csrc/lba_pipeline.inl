// Part of lba_host.cu (included there, inside extern "C"): the slslam_lba_pipeline_* entry points.
// Not a translation unit of its own: it calls the file-local batch_create_impl of lba_host.cu.
// ---- pipelined host-buffer entry points: submit() hands a batch to one of `depth` slots (own device pool, pinned
// staging, stream and -- with SLSLAM_PIPELINE_ASYNC_HOST -- own host thread, which plans, stages and enqueues it) and
// returns; wait() blocks on that batch and writes the results back.  The host work and the H2D copy of batch k+1 run
// while the device solves batch k; with the host threads two batches are planned / staged side by side. ----
struct slslam_lba_pipeline {
  int device = 0, depth = 2, flags = 0;
  struct Slot {
    slslam::Workspace ws;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    slslam_lba_batch* b = nullptr;
    std::vector<slslam_lba_desc> descs;
    std::vector<double*> params;
    slslam_summary* summ = nullptr;
    int64_t ticket = -1;
    // host thread of the slot (SLSLAM_PIPELINE_ASYNC_HOST): state 0 idle, 1 posted, 2 enqueued (rc / err valid)
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    int state = 0, rc = SLSLAM_OK;
    bool quit = false;
    std::string err;
  };
  std::vector<std::unique_ptr<Slot>> slots;
  int64_t next_ticket = 0;
};

// plan + stage + H2D + launch + D2H enqueue of the slot's posted batch, on the slot's stream
static int pipeline_enqueue(slslam_lba_pipeline* p, slslam_lba_pipeline::Slot& s) {
  cudaSetDevice(p->device);
  const int n = (int)s.descs.size();
  slslam_lba_batch* b = nullptr;
  int rc = batch_create_impl(n, s.descs.data(), (const double* const*)s.params.data(), -1, 0, &s.ws, s.stream, &b);
  if (rc != SLSLAM_OK) return rc;
  rc = slslam_lba_batch_solve(b, s.stream);
  if (rc == SLSLAM_OK) {
    slslam_summary* h_summ = (slslam_summary*)(b->h_params + b->total_params);
    cudaError_t e = cudaMemcpyAsync(b->h_params, b->d_params_out, b->total_params * 8, cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_summ, b->d_summ, sizeof(slslam_summary) * n, cudaMemcpyDeviceToHost, s.stream);
    if (e == cudaSuccess) e = cudaEventRecord(s.done, s.stream);
    if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); cudaGetLastError(); rc = SLSLAM_ERR_CUDA; }
  }
  if (rc != SLSLAM_OK) {
    cudaStreamSynchronize(s.stream);
    slslam_lba_batch_destroy(b);
    return rc;
  }
  s.b = b;
  return SLSLAM_OK;
}

static void pipeline_worker(slslam_lba_pipeline* p, slslam_lba_pipeline::Slot* s) {
  for (;;) {
    {
      std::unique_lock<std::mutex> lk(s->m);
      s->cv.wait(lk, [&] { return s->state == 1 || s->quit; });
      if (s->quit) return;
    }
    const int rc = pipeline_enqueue(p, *s);
    {
      std::lock_guard<std::mutex> lk(s->m);
      s->rc = rc;
      s->err = rc == SLSLAM_OK ? "" : slslam_last_error();
      s->state = 2;
    }
    s->cv.notify_all();
  }
}

// Blocks until the slot's batch (if any) has left the device, writes its results to the caller's arrays, frees the slot.
static int pipeline_finish(slslam_lba_pipeline* p, slslam_lba_pipeline::Slot& s) {
  if (s.ticket < 0) return SLSLAM_OK;
  int rc = SLSLAM_OK;
  if (p->flags & SLSLAM_PIPELINE_ASYNC_HOST) {
    std::unique_lock<std::mutex> lk(s.m);
    s.cv.wait(lk, [&] { return s.state == 2; });
    rc = s.rc;
    if (rc != SLSLAM_OK) set_last_error(s.err.c_str());
    s.state = 0;
  }
  slslam_lba_batch* b = s.b;
  if (rc == SLSLAM_OK && b) {
    cudaError_t e = cudaEventSynchronize(s.done);
    if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); cudaGetLastError(); rc = SLSLAM_ERR_CUDA; }
  }
  if (rc == SLSLAM_OK && b) {
    const slslam_summary* h_summ = (const slslam_summary*)(b->h_params + b->total_params);
    for (int i = 0; i < b->n; ++i) {
      memcpy(s.params[i], b->h_params + b->param_off[i], (size_t)b->nparams[i] * 8);
      if (s.summ) s.summ[i] = h_summ[i];
    }
  }
  if (b) slslam_lba_batch_destroy(b);
  s.b = nullptr; s.ticket = -1;
  return rc;
}

int slslam_lba_pipeline_create(int32_t device, int32_t depth, int32_t flags, slslam_lba_pipeline** out) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!out) return SLSLAM_ERR_INVALID;
  *out = nullptr;
  if (depth < 1 || depth > 8) return SLSLAM_ERR_INVALID;
  int rc = ensure_device(device);
  if (rc != SLSLAM_OK) return rc;
  slslam_lba_pipeline* p = new (std::nothrow) slslam_lba_pipeline();
  if (!p) return SLSLAM_ERR_INVALID;
  cudaGetDevice(&p->device);
  p->depth = depth; p->flags = flags;
  for (int k = 0; k < depth; ++k) p->slots.emplace_back(new slslam_lba_pipeline::Slot());
  for (auto& sp : p->slots) {
    auto& s = *sp;
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) {
      set_last_error(cudaGetErrorString(cudaGetLastError()));
      slslam_lba_pipeline_destroy(p);
      return SLSLAM_ERR_CUDA;
    }
    if (flags & SLSLAM_PIPELINE_ASYNC_HOST) s.th = std::thread(pipeline_worker, p, &s);
  }
  *out = p;
  return SLSLAM_OK;
}

int slslam_lba_pipeline_submit(slslam_lba_pipeline* p, int32_t n, const slslam_lba_desc* descs, double* const* params_inout,
                               slslam_summary* summaries_out, int64_t* ticket_out) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!p || n <= 0 || !descs || !params_inout) return SLSLAM_ERR_INVALID;
  // argument errors are reported here, before anything is queued (the same checks run again inside the enqueue)
  for (int i = 0; i < n; ++i) {
    const int rc = validate_desc(descs[i]);
    if (rc != SLSLAM_OK) return rc;
    if (!params_inout[i]) return SLSLAM_ERR_INVALID;
  }
  cudaSetDevice(p->device);
  auto& s = *p->slots[(size_t)(p->next_ticket % p->depth)];
  int rc = pipeline_finish(p, s);        // the slot's previous batch (submitted `depth` calls ago) must have drained
  if (rc != SLSLAM_OK) return rc;
  s.descs.assign(descs, descs + n);
  s.params.assign(params_inout, params_inout + n);
  s.summ = summaries_out;
  if (p->flags & SLSLAM_PIPELINE_ASYNC_HOST) {
    { std::lock_guard<std::mutex> lk(s.m); s.state = 1; }
    s.cv.notify_all();
  } else {
    rc = pipeline_enqueue(p, s);
    if (rc != SLSLAM_OK) return rc;
  }
  s.ticket = p->next_ticket++;
  if (ticket_out) *ticket_out = s.ticket;
  return SLSLAM_OK;
}

int slslam_lba_pipeline_wait(slslam_lba_pipeline* p, int64_t ticket) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!p) return SLSLAM_ERR_INVALID;
  cudaSetDevice(p->device);
  int rc = SLSLAM_OK;
  for (auto& sp : p->slots) {
    if (sp->ticket < 0) continue;
    if (ticket < 0 || sp->ticket == ticket) {
      const int r = pipeline_finish(p, *sp);
      if (r != SLSLAM_OK) rc = r;
    }
  }
  return rc;
}

void slslam_lba_pipeline_destroy(slslam_lba_pipeline* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  for (auto& sp : p->slots) {
    auto& s = *sp;
    if (s.th.joinable()) {
      {
        std::unique_lock<std::mutex> lk(s.m);
        s.cv.wait(lk, [&] { return s.state != 1; });    // let a posted batch finish its enqueue
        s.quit = true;
      }
      s.cv.notify_all();
      s.th.join();
    }
    if (s.b) { cudaStreamSynchronize(s.stream); slslam_lba_batch_destroy(s.b); s.b = nullptr; }   // abandoned: results dropped
    if (s.stream) cudaStreamDestroy(s.stream);
    if (s.done) cudaEventDestroy(s.done);
    s.ws.release();
  }
  delete p;
}

