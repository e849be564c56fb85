// Host side of the LBA path behind the C ABI (include/slslam_b200.h): argument checks, staging of the caller's arrays,
// launches (device planner -> solve kernel, or the motion-only kernel), download; plus the HOST planner, which builds the
// same per-solve "symbolic" plan as lba_plan_kernel.cuh (observations grouped by line and packed into 32-lane tiles,
// lines partitioned over the CTAs of a window's group, camera-pair list for the Schur blocks) and serves as fallback
// and as parity reference of the device planner.  Parts: lba_device_plan.inl, moba_host.inl, lba_pipeline.inl.
// Replaces what LBAProblem::build + ceres::Solve do on the host (reference src/lba_problem.cpp:54-93).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "common_host.h"
#include "lba_kernel.cuh"
#include "moba_kernel.cuh"
#include "lba_plan_kernel.cuh"
#include "wide_kernel.cuh"

namespace slslam {

struct WindowPlan {
  int C = 0, Cf = 0, L = 0, N = 0, nkeys = 0;
  int max_iters = 0, robust = 1;
  double huber_a, baseline, ftol, gtol, ptol, radius0;
  std::vector<int> slot_src;      // [slots] index of the caller's observation in this slot, or -1 (padding)
  std::vector<int2> meta;         // [slots]
  std::vector<int> line_gid;      // device lines
  std::vector<uint32_t> items;
  std::vector<int> key_off;       // [CS*(nkeys+1)]
  int cta_slot_off[MAX_G + 1], cta_line_off[MAX_G + 1];
  signed char cam_free[MAX_CAMS];
  int max_lines_cta = 0, max_slots_cta = 0, max_items_cta = 0;
  bool has_unobserved_blocks = false;
  // planning scratch, kept so that a cached plan object does not allocate on reuse
  std::vector<char> s_cam_used, s_cam_const, s_line_const;
  std::vector<int> s_line_cnt, s_line_start, s_order, s_fill, s_first_slot, s_kcount, s_sorted_off;
  std::vector<signed char> s_ocf;
  std::vector<unsigned char> s_sorted_pos, s_sorted_cnt;
};

static int validate_desc(const slslam_lba_desc& d) {
  if (d.num_cameras < 0 || d.num_lines < 0 || d.num_observations < 0 || d.max_iterations < 0) return SLSLAM_ERR_INVALID;
  if (d.num_observations > 0 && (!d.camera_index || !d.line_index || !d.fixed_index || !d.observations)) return SLSLAM_ERR_INVALID;
  if (d.num_cameras > MAX_CAMS) return SLSLAM_ERR_UNSUPPORTED;
  for (int i = 0; i < d.num_observations; ++i) {
    if (d.camera_index[i] < 0 || d.camera_index[i] >= d.num_cameras) return SLSLAM_ERR_INVALID;
    if (d.line_index[i] < 0 || d.line_index[i] >= d.num_lines) return SLSLAM_ERR_INVALID;
  }
  return SLSLAM_OK;
}

// Builds the device plan of one window for a cluster of CS CTAs.
static int build_plan(const slslam_lba_desc& d, int CS, WindowPlan& p) {
  const int C = d.num_cameras, L = d.num_lines, N = d.num_observations;
  p.C = C; p.L = L; p.N = N; p.max_iters = d.max_iterations; p.robust = d.robust ? 1 : 0;
  p.huber_a = d.huber_delta > 0 ? d.huber_delta : 1.0 / 406.05;     // reference lba_problem.cpp:78-80
  p.baseline = d.baseline >= 0 ? d.baseline : 0.12;                 // reference lba_problem.h:101
  p.ftol = d.function_tolerance > 0 ? d.function_tolerance : 1e-6;  // Ceres 1.7.0 defaults
  p.gtol = d.gradient_tolerance > 0 ? d.gradient_tolerance : 1e-10;
  p.ptol = d.parameter_tolerance > 0 ? d.parameter_tolerance : 1e-8;
  p.radius0 = d.initial_trust_region_radius > 0 ? d.initial_trust_region_radius : 1e4;
  // sticky constants per block (reference lba_problem.cpp:88-91); unobserved blocks are never touched
  p.has_unobserved_blocks = false;
  std::vector<char>&cam_used = p.s_cam_used, &cam_const = p.s_cam_const, &line_const = p.s_line_const;
  cam_used.assign(C, 0); cam_const.assign(C, 0); line_const.assign(L, 0);
  std::vector<int>& line_cnt = p.s_line_cnt;
  line_cnt.assign(L, 0);
  for (int i = 0; i < N; ++i) {
    cam_used[d.camera_index[i]] = 1;
    ++line_cnt[d.line_index[i]];
    if (d.fixed_index[2 * i]) cam_const[d.camera_index[i]] = 1;
    if (d.fixed_index[2 * i + 1]) line_const[d.line_index[i]] = 1;
  }
  p.Cf = 0;
  for (int c = 0; c < MAX_CAMS; ++c) p.cam_free[c] = -1;
  for (int c = 0; c < C; ++c) {
    if (cam_used[c] && !cam_const[c]) p.cam_free[c] = (signed char)p.Cf++;
    if (!cam_used[c]) p.has_unobserved_blocks = true;
  }
  if (p.Cf > MAX_FREE_CAMS) return SLSLAM_ERR_UNSUPPORTED;
  p.nkeys = p.Cf * (p.Cf + 1) / 2;
  // group observations by line (stable counting sort: keeps the caller's order inside a line)
  std::vector<int>& line_start = p.s_line_start;
  line_start.assign(L + 1, 0);
  for (int l = 0; l < L; ++l) {
    line_start[l + 1] = line_start[l] + line_cnt[l];
    if (line_cnt[l] > 32) return SLSLAM_ERR_UNSUPPORTED;
    if (line_cnt[l] == 0) p.has_unobserved_blocks = true;
  }
  std::vector<int>&order = p.s_order, &fill = p.s_fill;
  order.resize(N); fill.assign(line_start.begin(), line_start.end() - 1);
  for (int i = 0; i < N; ++i) order[fill[d.line_index[i]]++] = i;
  std::vector<int>& dl = p.line_gid;   // device lines in increasing id
  dl.clear(); dl.reserve(L);
  for (int l = 0; l < L; ++l) if (line_cnt[l] > 0) dl.push_back(l);
  // partition the device lines over the CTAs, balancing observation counts
  const int nd = (int)dl.size();
  p.cta_line_off[0] = 0;
  {
    int li = 0; long long seen = 0;
    for (int r = 0; r < CS; ++r) {
      const long long target = ((long long)N * (r + 1) + CS - 1) / CS;
      while (li < nd && (seen < target || r == CS - 1)) { seen += line_cnt[dl[li]]; ++li; }
      p.cta_line_off[r + 1] = li;
    }
    for (int r = CS + 1; r <= MAX_G; ++r) p.cta_line_off[r] = nd;
  }
  p.key_off.assign((size_t)CS * (p.nkeys + 1), 0);
  p.slot_src.clear(); p.meta.clear(); p.items.clear();
  p.slot_src.reserve((size_t)N + 32 * (size_t)CS + N / 4); p.meta.reserve((size_t)N + 32 * (size_t)CS + N / 4);
  p.max_lines_cta = 1; p.max_slots_cta = 32; p.max_items_cta = 0;
  p.cta_slot_off[0] = 0;
  // reduced camera index of every observation in line-grouped order (-1: constant camera or constant line => no pair)
  std::vector<signed char>& ocf = p.s_ocf;
  ocf.resize(N);
  for (int l = 0; l < L; ++l)
    for (int a = line_start[l]; a < line_start[l + 1]; ++a)
      ocf[a] = line_const[l] ? (signed char)-1 : p.cam_free[d.camera_index[order[a]]];
  std::vector<int>&first_slot = p.s_first_slot, &kcount = p.s_kcount, &sorted_off = p.s_sorted_off;
  first_slot.assign(nd, 0); kcount.assign(p.nkeys + 1, 0); sorted_off.assign(nd, 0);
  std::vector<unsigned char>&sorted_pos = p.s_sorted_pos, &sorted_cnt = p.s_sorted_cnt;
  sorted_pos.resize((size_t)N + 1); sorted_cnt.assign(nd, 0);
  int pos_off = 0;
  for (int r = 0; r < CS; ++r) {
    const int lb = p.cta_line_off[r], le = p.cta_line_off[r + 1];
    p.max_lines_cta = std::max(p.max_lines_cta, le - lb);
    const size_t slot_base = p.meta.size();
    int lane = 0;
    int tile_rounds[MAX_CAMS];          // next free accumulator round per camera in the current tile
    for (int c = 0; c < MAX_CAMS; ++c) tile_rounds[c] = 0;
    auto pad_tile = [&]() {
      while (lane != 0 && lane < 32) {
        int2 m; m.x = 0 | (lane << 8) | (1 << 14); m.y = 0;
        p.meta.push_back(m);
        p.slot_src.push_back(-1);
        ++lane;
      }
      lane = 0;
      for (int c = 0; c < C; ++c) tile_rounds[c] = 0;
    };
    std::fill(kcount.begin(), kcount.end(), 0);
    for (int li = lb; li < le; ++li) {
      const int l = dl[li], k = line_cnt[l], ls0 = line_start[l];
      if (lane + k > 32) pad_tile();
      const int seg_start = lane;
      first_slot[li] = (int)(p.meta.size() - slot_base);
      const int lflag = line_const[l] ? F_LINE_FIXED : 0;
      for (int a = 0; a < k; ++a) {
        const int i = order[ls0 + a];
        const int cam = d.camera_index[i];
        int flags = F_VALID | lflag;
        if (cam_const[cam]) flags |= F_CAM_FIXED;
        if (a == 0) flags |= F_HEAD;
        const int round = tile_rounds[cam]++;   // lanes sharing a camera inside a tile get distinct rounds
        int2 m;
        m.x = cam | (seg_start << 8) | (k << 14) | (flags << 24);
        m.y = (li - lb) | (round << 20);
        p.meta.push_back(m);
        p.slot_src.push_back(i);
        ++lane;
      }
      if (lane == 32) { lane = 0; for (int c = 0; c < C; ++c) tile_rounds[c] = 0; }
      // count the Schur pairs of this line: ordered pairs (a,b) of free-camera observations with cf_a >= cf_b
      // (cameras are distinct within a line).  The line's free observations are sorted by reduced camera index first
      // (insertion sort; the reference packs them in keyframe order, so they usually are already), which makes the
      // pair loops branch-free: every (i, j < i) of the sorted list is a pair.  The same-observation term (i, i), the
      // whole of a diagonal block, is folded into the camera accumulators by the linearisation sweep instead.
      {
        int m = 0;
        for (int a = 0; a < k; ++a) {
          const int ca = ocf[ls0 + a];
          if (ca < 0) continue;
          int q = m++;
          while (q > 0 && ocf[ls0 + sorted_pos[pos_off + q - 1]] > ca) { sorted_pos[pos_off + q] = sorted_pos[pos_off + q - 1]; --q; }
          sorted_pos[pos_off + q] = (unsigned char)a;
        }
        sorted_cnt[li] = (unsigned char)m; sorted_off[li] = pos_off;
        for (int i = 0; i < m; ++i) {
          const int ca = ocf[ls0 + sorted_pos[pos_off + i]], base = ca * (ca + 1) / 2;
          for (int j = 0; j < i; ++j) {
            const int cb = ocf[ls0 + sorted_pos[pos_off + j]];
            // A camera observing the line twice (never produced by the reference; the device planner hands such windows
            // to this one): the two observations' cross term lands on the DIAGONAL block, which must stay symmetric,
            // -(Z_i Z_j^T + Z_j Z_i^T) -- both orderings are emitted (the reduced solve reads the lower triangle).
            kcount[base + cb] += (cb == ca) ? 2 : 1;
          }
        }
        pos_off += m;
      }
    }
    pad_tile();
    const int nslots = (int)(p.meta.size() - slot_base);
    if (nslots > 65535) return SLSLAM_ERR_UNSUPPORTED;
    p.max_slots_cta = std::max(p.max_slots_cta, nslots);
    p.cta_slot_off[r + 1] = (int)p.meta.size();
    // offsets of this CTA's pair blocks, then the fill pass (same loop order => lines ascending inside a block)
    int* ko = &p.key_off[(size_t)r * (p.nkeys + 1)];
    int run = (int)p.items.size();
    for (int key = 0; key < p.nkeys; ++key) { ko[key] = run; run += kcount[key]; kcount[key] = ko[key]; }
    ko[p.nkeys] = run;
    p.max_items_cta = std::max(p.max_items_cta, run - ko[0]);
    p.items.resize((size_t)run);
    for (int li = lb; li < le; ++li) {
      const int l = dl[li], k = line_cnt[l], ls0 = line_start[l];
      const uint32_t fs = (uint32_t)first_slot[li];
      const int m = sorted_cnt[li];
      const unsigned char* sp = &sorted_pos[sorted_off[li]];
      (void)k;
      for (int i = 0; i < m; ++i) {
        const int ca = ocf[ls0 + sp[i]], base = ca * (ca + 1) / 2;
        const uint32_t si = fs + sp[i];
        for (int j = 0; j < i; ++j) {
          const int cb = ocf[ls0 + sp[j]];
          const uint32_t sj = fs + sp[j];
          p.items[(size_t)kcount[base + cb]++] = si | (sj << 16);
          if (cb == ca) p.items[(size_t)kcount[base + cb]++] = sj | (si << 16);
        }
      }
    }
  }
  for (int r = CS + 1; r <= MAX_G; ++r) p.cta_slot_off[r] = (int)p.meta.size();
  return SLSLAM_OK;
}

}  // namespace slslam

using namespace slslam;
namespace slslam { struct Workspace; }

struct slslam_lba_batch {
  int n = 0, device = 0, CS = 1;
  std::vector<WindowPlan> plans;
  std::vector<size_t> param_off, trace_off;   // in doubles
  std::vector<int> nparams;
  size_t total_params = 0, total_trace = 0;
  SmemLayout lay;
  size_t smem_bytes = 0;
  // device
  char* d_pool = nullptr;
  WinHdr* d_hdrs = nullptr;
  double *d_params_in = nullptr, *d_params_out = nullptr, *d_trace = nullptr;
  slslam_summary* d_summ = nullptr;
  long long* d_phase = nullptr;
  // pinned host staging
  double* h_params = nullptr;
  size_t upload_bytes = 0;
  int max_active = 0;   // windows of this shape the device keeps resident at once (a larger batch runs in waves)
  unsigned int* d_bar = nullptr; size_t bar_bytes = 0;   // group barrier counters, zeroed before every launch
  bool inplace = false;    // device-resident inputs: parameters are read and written where the caller keeps them
  bool deferred = false;   // launched without reading the planner's sizes back: its flags are checked after the solve
  slslam::PlanInfo* d_info = nullptr; int Cmax = 1, smem_optin = 0; bool want_zg = false;
  bool borrowed = false;   // device pool and pinned staging belong to a Workspace (the calling thread's or a pipeline slot's)
  slslam::Workspace* ws = nullptr;
  // device-side plan (lba_plan_kernel.cuh): where its outputs live in the pool, for the planner parity check
  bool device_planned = false;
  struct DevPlanOff { size_t obs, meta, gid, items, koff, hdr; int slot_cap, item_cap; };
  std::vector<DevPlanOff> dp;
  std::vector<slslam::PlanInfo> dp_info;
};

namespace slslam {
// Grow-only device pool + pinned staging cached per host thread, so that the one-shot entry points
// (slslam_lba_solve / slslam_lba_solve_batch: plan + H2D + solve + D2H per call, the way the reference calls
// ceres::Solve once per keyframe) do not pay cudaMalloc / cudaMallocHost / cudaFree on every call.
struct Workspace {
  int device = -1;
  char* d_pool = nullptr; size_t d_cap = 0;
  char* h_pin = nullptr; size_t h_cap = 0;     // upload staging: write-combined, written once by one thread, read only by the DMA engine
  char* h_res = nullptr; size_t r_cap = 0;     // results: ordinary pinned memory (the CPU reads it)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t copy_stream = nullptr;          // page-locked observation arrays are DMA'd on their own stream, beside staging + planner
  cudaEvent_t ev_obs = nullptr;
  std::vector<WindowPlan> plans;               // plan objects of the previous call, reused for their capacity
  int ensure(int dev, size_t d_bytes, size_t h_bytes, size_t r_bytes) {
    if (device != dev) { release(); device = dev; }
    for (int k = 0; k < 4; ++k) if (!ev[k]) CUDA_TRY(cudaEventCreate(&ev[k]));
    if (!copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    if (!ev_obs) CUDA_TRY(cudaEventCreateWithFlags(&ev_obs, cudaEventDisableTiming));
    if (d_bytes > d_cap) {
      if (d_pool) cudaFree(d_pool);
      d_pool = nullptr; d_cap = 0;
      const size_t want = d_bytes + d_bytes / 4;
      CUDA_TRY(cudaMalloc((void**)&d_pool, want));
      d_cap = want;
    }
    if (h_bytes > h_cap) {
      if (h_pin) cudaFreeHost(h_pin);
      h_pin = nullptr; h_cap = 0;
      const size_t want = h_bytes + h_bytes / 4;
      CUDA_TRY(cudaHostAlloc((void**)&h_pin, want, getenv("SLSLAM_NO_WC_STAGING") ? cudaHostAllocDefault : cudaHostAllocWriteCombined));
      h_cap = want;
    }
    if (r_bytes > r_cap) {
      if (h_res) cudaFreeHost(h_res);
      h_res = nullptr; r_cap = 0;
      const size_t want = r_bytes + r_bytes / 4;
      CUDA_TRY(cudaMallocHost((void**)&h_res, want));
      r_cap = want;
    }
    return SLSLAM_OK;
  }
  void release() {
    if (d_pool) cudaFree(d_pool);
    if (h_pin) cudaFreeHost(h_pin);
    if (h_res) cudaFreeHost(h_res);
    d_pool = nullptr; h_pin = nullptr; h_res = nullptr; d_cap = h_cap = r_cap = 0;
    for (int k = 0; k < 4; ++k) { if (ev[k]) cudaEventDestroy(ev[k]); ev[k] = nullptr; }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (ev_obs) cudaEventDestroy(ev_obs);
    copy_stream = nullptr; ev_obs = nullptr;
  }
};
static thread_local Workspace g_ws;
// host wall-clock split of this thread's last one-shot solve, ms: plan | stage + H2D enqueue | launch + wait + D2H | copy-out | total
static thread_local double g_timing[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // + device ms: H2D | kernel | D2H
static inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace slslam

namespace slslam {

// CTAs of the solve kernel that can be resident at once on `device` (1 per SM: 255 registers x 256 threads).
static int resident_ctas(int device) {
  static int table[16];
  static bool known[16];
  if (device < 0 || device >= 16) return 0;
  if (!known[device]) {
    int sms = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lba_solve_kernel, LBA_NT, 0) != cudaSuccess) { cudaGetLastError(); per_sm = 0; }
    table[device] = sms * std::max(per_sm, 0); known[device] = true;
  }
  return table[device];
}

// CTAs per window.  One LM iteration of a window on G CTAs costs about fixed + per_obs * N / G (profiles/: ~45 k +
// 156 N / G cycles on B200) and all CTAs of a launch must be co-resident.  A batch that does not fit in one launch runs
// in waves; the number of waves and the group size are chosen together so that the waves are equally full (29 + 3
// windows cost two full waves; 16 + 16 with larger groups is faster).  `cs_min`: lower bound from shared memory (the
// per-line state of a CTA must fit); at least three 32-lane tiles per CTA; beyond ~48 CTAs the exchange grows faster
// than the sweeps shrink (profiles/r1_cluster_vs_group.txt).
static int pick_group_size_for(int cap, int nwin, long long max_obs, int requested, int cs_min) {
  if (requested > 0) return std::min(requested, (int)MAX_G);
  cap = std::max(1, cap);
  cs_min = std::max(1, std::min(cs_min, std::min(cap, (int)MAX_G)));
  const long long tiles = (max_obs + 27) / 28;
  const int cs_max = std::max(cs_min, (int)std::min<long long>(48, std::max<long long>(1, tiles / 3)));
  const int w0 = (nwin + cap / cs_min - 1) / std::max(1, cap / cs_min);
  int best = cs_min;
  double best_t = 1e300;
  for (int w = std::max(1, w0); w <= w0 + 3; ++w) {
    const int per = (nwin + w - 1) / w;
    const int cs = std::max(cs_min, std::min(cs_max, cap / std::max(1, per)));
    if ((long long)cs * per > cap) continue;
    const double t = w * (45000.0 + 156.0 * (double)max_obs / cs);
    if (t < best_t - 1e-9) { best_t = t; best = cs; }
  }
  return best;
}

static int pick_group_size(int device, int nwin, long long max_obs, int requested, int cs_min = 1) {
  return pick_group_size_for(resident_ctas(device), nwin, max_obs, requested, cs_min);
}

// Shared-memory lower bound on the group size: 50 doubles of per-line state per line beside ~48 KB of fixed state.
static int min_group_size_for_lines(int Lmax, int smem_optin) {
  const long long per_line = 50 * 8, room = std::max(16384, smem_optin - 49152);
  return (int)std::max<long long>(1, std::min<long long>(MAX_G, ((long long)Lmax * per_line * 11 / 10 + room - 1) / room));
}

// Windows per wave for `n` windows when `fit` groups are resident at once: equally full waves.
static int balanced_wave(int n, int fit) {
  fit = std::max(1, fit);
  const int waves = (n + fit - 1) / fit;
  return std::max(1, (n + waves - 1) / waves);
}

// fn(i) for i in [0, n) on up to 8 host threads (the plans of different windows are independent work).  The workers
// are created once and parked on a condition variable: creating and joining 8 threads per solve cost ~0.1 ms of a
// 2 ms call.  One job at a time (callers are serialised by a mutex); the calling thread takes a share of the work.
class HostPool {
 public:
  static HostPool& get() { static HostPool* p = new HostPool(); return *p; }   // never destroyed: no join at exit
  template <class F>
  void run(int n, F& fn) {
    const int nthreads = std::min(n, max_threads_);
    if (nthreads <= 1) { for (int i = 0; i < n; ++i) fn(i); return; }
    std::lock_guard<std::mutex> serial(job_mutex_);
    ensure_workers(nthreads - 1);
    {
      std::lock_guard<std::mutex> lk(m_);
      call_ = [](void* f, int i) { (*static_cast<F*>(f))(i); };
      fn_ = &fn; n_ = n; next_.store(0); pending_ = nthreads - 1; active_ = nthreads - 1; ++generation_;
    }
    cv_.notify_all();
    for (int i = next_.fetch_add(1); i < n; i = next_.fetch_add(1)) fn(i);
    std::unique_lock<std::mutex> lk(m_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
  }
 private:
  HostPool() {
    const char* e = getenv("SLSLAM_HOST_THREADS");
    const int v = e ? atoi(e) : 8;
    max_threads_ = v < 1 ? 1 : (v > 64 ? 64 : v);
  }
  void ensure_workers(int k) {
    while ((int)workers_.size() < k) {
      const int id = (int)workers_.size();
      workers_.emplace_back([this, id]() { worker(id); });
      workers_.back().detach();
    }
  }
  void worker(int id) {
    unsigned long long seen = 0;
    for (;;) {
      void (*call)(void*, int); void* fn; int n;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return generation_ != seen && id < active_; });
        seen = generation_; call = call_; fn = fn_; n = n_;
      }
      for (int i = next_.fetch_add(1); i < n; i = next_.fetch_add(1)) call(fn, i);
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) done_cv_.notify_one();
      }
    }
  }
  std::mutex m_, job_mutex_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> workers_;
  void (*call_)(void*, int) = nullptr;
  void* fn_ = nullptr;
  int n_ = 0, pending_ = 0, active_ = 0, max_threads_ = 8;
  std::atomic<int> next_{0};
  unsigned long long generation_ = 0;
};

template <class F>
static void parallel_for(int n, F fn) { HostPool::get().run(n, fn); }

// The dynamic shared memory opt-in of the solve kernel is raised once per device to the maximum and never lowered:
// batches of different shapes may be enqueued from several host threads (pipeline slots), and a per-batch value would
// race with another thread's launch.
static int set_solve_kernel_smem_limit(int device, int smem_optin) {
  static std::mutex attr_mutex;
  static bool attr_set[16] = {false};
  std::lock_guard<std::mutex> lk(attr_mutex);
  if (device < 0 || device >= 16 || !attr_set[device]) {
    CUDA_TRY(cudaFuncSetAttribute(lba_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin));
    if (device >= 0 && device < 16) attr_set[device] = true;
  }
  return SLSLAM_OK;
}

static int build_plans(int n, const slslam_lba_desc* descs, int CS, std::vector<WindowPlan>& plans) {
  plans.resize(n);   // existing plan objects keep their vector capacity
  std::vector<int> rcs(n, SLSLAM_OK);
  parallel_for(n, [&](int i) { rcs[i] = build_plan(descs[i], CS, plans[i]); });
  for (int i = 0; i < n; ++i) if (rcs[i] != SLSLAM_OK) return rcs[i];
  return SLSLAM_OK;
}

static int batch_create_host_plan(int32_t n, const slslam_lba_desc* descs, const double* const* params, int32_t device,
                                  int32_t cluster_size, Workspace* ws, cudaStream_t stream, slslam_lba_batch** out) {
  if (!out) return SLSLAM_ERR_INVALID;
  *out = nullptr;
  if (n <= 0 || !descs || !params) return SLSLAM_ERR_INVALID;
  for (int i = 0; i < n; ++i) {
    const int rc = validate_desc(descs[i]);
    if (rc != SLSLAM_OK) return rc;
    if (!params[i]) return SLSLAM_ERR_INVALID;
    const int np = 6 * descs[i].num_cameras + 4 * descs[i].num_lines;
    for (int k = 0; k < np; ++k) if (!std::isfinite(params[i][k])) return SLSLAM_ERR_NUMERICAL;
  }
  const double t_begin = now_ms();
  int rc = ensure_device(device);
  if (rc != SLSLAM_OK) return rc;
  slslam_lba_batch* b = new (std::nothrow) slslam_lba_batch();
  if (!b) return SLSLAM_ERR_INVALID;
  cudaGetDevice(&b->device);
  b->n = n;
  b->borrowed = ws != nullptr;
  b->ws = ws;
  if (ws) b->plans.swap(ws->plans);
  long long max_obs = 0;
  for (int i = 0; i < n; ++i) max_obs = std::max<long long>(max_obs, descs[i].num_observations);
  const int cap = resident_ctas(b->device);
  int smem_optin = 0;
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, b->device);
  bool placed = false;
  int Lmax = 0;
  for (int i = 0; i < n; ++i) Lmax = std::max(Lmax, descs[i].num_lines);
  int CS = pick_group_size(b->device, n, max_obs, cluster_size, min_group_size_for_lines(Lmax, smem_optin));
  for (int attempt = 0; attempt < 8 && !placed; ++attempt) {
    rc = build_plans(n, descs, CS, b->plans);
    if (rc != SLSLAM_OK) { delete b; return rc; }
    int Cmax = 1, Cfmax = 0, mlines = 1, mslots = 32, mitems = 0;
    for (int i = 0; i < n; ++i) {
      Cmax = std::max(Cmax, b->plans[i].C); Cfmax = std::max(Cfmax, b->plans[i].Cf);
      mlines = std::max(mlines, b->plans[i].max_lines_cta); mslots = std::max(mslots, b->plans[i].max_slots_cta);
      mitems = std::max(mitems, b->plans[i].max_items_cta);
    }
    b->lay = lba_layout(Cmax, Cfmax, mlines, mslots, CS, (size_t)smem_optin, mitems);
    b->smem_bytes = (size_t)b->lay.total * 8;
    b->CS = CS;
    if (b->smem_bytes > (size_t)smem_optin) {
      // the fixed part alone does not fit: too many lines per CTA.  More CTAs per window (fewer windows per wave).
      if (cluster_size > 0 || CS >= MAX_G) break;
      CS = std::min((int)MAX_G, CS * 2);
      continue;
    }
    { const int arc = set_solve_kernel_smem_limit(b->device, smem_optin); if (arc != SLSLAM_OK) { delete b; return arc; } }
    if (CS > cap) { set_last_error("more CTAs per window than the device keeps resident"); delete b; return SLSLAM_ERR_CUDA; }
    b->max_active = balanced_wave(n, cap / CS);
    placed = true;
  }
  if (!placed) {
    set_last_error("no CTA group shape fits this batch in shared memory");
    delete b;
    return SLSLAM_ERR_UNSUPPORTED;
  }

  const double t_planned = now_ms();
  // ---- pooled device allocation: [uploaded: headers | initial parameters | plans] [device only: results | Z staging] ----
  size_t off = 0;
  auto reserve = [&](size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; };
  const size_t o_hdr = reserve(sizeof(WinHdr) * n);
  b->param_off.resize(n); b->trace_off.resize(n); b->nparams.resize(n);
  size_t tp = 0, tt = 0;
  for (int i = 0; i < n; ++i) {
    b->nparams[i] = 6 * b->plans[i].C + 4 * b->plans[i].L;
    b->param_off[i] = tp; tp += (size_t)((b->nparams[i] + 1) & ~1);
    b->trace_off[i] = tt; tt += (size_t)std::max(1, b->plans[i].max_iters) * SLSLAM_TRACE_WIDTH;
  }
  b->total_params = tp; b->total_trace = tt;
  const size_t o_pin = reserve(tp * 8);
  std::vector<size_t> o_obs(n), o_meta(n), o_gid(n), o_items(n), o_koff(n), o_z(n);
  for (int i = 0; i < n; ++i) {
    const WindowPlan& p = b->plans[i];
    o_obs[i] = reserve(p.slot_src.size() * 64); o_meta[i] = reserve(p.meta.size() * sizeof(int2));
    o_gid[i] = reserve(p.line_gid.size() * 4 + 4); o_items[i] = reserve(p.items.size() * 4 + 4);
    o_koff[i] = reserve(p.key_off.size() * 4 + 4);
  }
  const size_t upload = off;
  const size_t o_pout = reserve(tp * 8), o_summ = reserve(sizeof(slslam_summary) * n), o_trace = reserve(tt * 8),
               o_phase = reserve(sizeof(long long) * NPHASE * n);
  std::vector<size_t> o_vg(n), o_vr(n), o_sg(n);
  for (int i = 0; i < n; ++i) {
    const int vpad = (lba_vlen(b->plans[i].Cf) + 31) & ~31;
    o_vg[i] = reserve((size_t)b->CS * vpad * 8); o_vr[i] = reserve((size_t)vpad * 8); o_sg[i] = reserve((size_t)b->CS * 8 * 8);
  }
  const size_t o_bar = reserve((size_t)n * 128);   // one counter per window, 128 B apart
  b->bar_bytes = (size_t)n * 128;
  for (int i = 0; i < n; ++i) o_z[i] = b->lay.z_in_smem ? 0 : reserve(b->plans[i].meta.size() * ZST * 8);
  const size_t result_bytes = tp * 8 + sizeof(slslam_summary) * n + 256;
  char* host = nullptr;
  std::vector<char> host_vec;
  if (ws) {
    rc = ws->ensure(b->device, off, upload, result_bytes);
    if (rc != SLSLAM_OK) { delete b; return rc; }
    b->d_pool = ws->d_pool;
    host = ws->h_pin;
    b->h_params = (double*)ws->h_res;
  } else {
    CUDA_TRY_OR(cudaMalloc((void**)&b->d_pool, off), { delete b; return SLSLAM_ERR_CUDA; });
    host_vec.resize(upload);
    host = host_vec.data();
    CUDA_TRY_OR(cudaMallocHost((void**)&b->h_params, std::max<size_t>(tp, 1) * 8), { slslam_lba_batch_destroy(b); return SLSLAM_ERR_CUDA; });
  }
  b->d_hdrs = (WinHdr*)(b->d_pool + o_hdr);
  b->d_params_in = (double*)(b->d_pool + o_pin); b->d_params_out = (double*)(b->d_pool + o_pout);
  b->d_trace = (double*)(b->d_pool + o_trace); b->d_summ = (slslam_summary*)(b->d_pool + o_summ);
  b->d_phase = (long long*)(b->d_pool + o_phase);
  b->d_bar = (unsigned int*)(b->d_pool + o_bar);
  // Staging is done by ONE thread and sent as ONE copy: a pinned buffer written by several cores is read by the DMA
  // engine at 8.5 GB/s instead of 48 GB/s (measured on the B200 host, scripts/h2d_staging_probe.py), and interleaving
  // per-window copies with the staging of the next window slowed the staging more than the overlap saved.
  // (Measured again inside the pipeline, where the copy hides behind the previous kernel: parallel staging still lost,
  // 48.7 k vs 65.7 k LM iterations/s.  Each pipeline slot therefore stages with one thread -- its own.)
  auto stage_window = [&](int i) {
    const WindowPlan& p = b->plans[i];
    WinHdr h; memset(&h, 0, sizeof(h));
    h.C = p.C; h.Cf = p.Cf; h.L = p.L; h.n = 6 * p.Cf; h.nkeys = p.nkeys; h.vlen = lba_vlen(p.Cf);
    h.max_iters = p.max_iters; h.robust = p.robust;
    h.huber_a = p.huber_a; h.baseline = p.baseline; h.ftol = p.ftol; h.gtol = p.gtol; h.ptol = p.ptol; h.radius0 = p.radius0;
    h.obs = (const double*)(b->d_pool + o_obs[i]); h.meta = (const int2*)(b->d_pool + o_meta[i]);
    h.line_gid = (const int*)(b->d_pool + o_gid[i]); h.items = (const uint32_t*)(b->d_pool + o_items[i]);
    h.key_off = (const int*)(b->d_pool + o_koff[i]);
    h.params_in = b->d_params_in + b->param_off[i]; h.params_out = b->d_params_out + b->param_off[i];
    h.Zg = b->lay.z_in_smem ? nullptr : (double*)(b->d_pool + o_z[i]);
    h.Vg = (double*)(b->d_pool + o_vg[i]); h.Vr = (double*)(b->d_pool + o_vr[i]); h.scalg = (double*)(b->d_pool + o_sg[i]);
    h.bar = (unsigned int*)(b->d_pool + o_bar + (size_t)i * 128); h.vpad = (lba_vlen(p.Cf) + 31) & ~31;
    h.summary = b->d_summ + i;
    h.trace = ws ? nullptr : b->d_trace + b->trace_off[i];          // the one-shot path never reads the trace back
    h.phase_cycles = ws ? nullptr : b->d_phase + (size_t)NPHASE * i;
    memcpy(h.cta_slot_off, p.cta_slot_off, sizeof(h.cta_slot_off));
    memcpy(h.cta_line_off, p.cta_line_off, sizeof(h.cta_line_off));
    memcpy(h.cam_free, p.cam_free, sizeof(h.cam_free));
    memcpy(host + o_hdr + sizeof(WinHdr) * i, &h, sizeof(h));
    // observations are gathered from the caller's array straight into the (pinned) staging buffer, slot order
    double* so = (double*)(host + o_obs[i]);
    const double* src = descs[i].observations;
    const size_t ns = p.slot_src.size();
    for (size_t k = 0; k < ns; ++k) {
      const int j = p.slot_src[k];
      if (j >= 0) memcpy(so + 8 * k, src + 8 * (size_t)j, 64); else memset(so + 8 * k, 0, 64);
    }
    memcpy(host + o_meta[i], p.meta.data(), p.meta.size() * sizeof(int2));
    memcpy(host + o_gid[i], p.line_gid.data(), p.line_gid.size() * 4);
    memcpy(host + o_items[i], p.items.data(), p.items.size() * 4);
    memcpy(host + o_koff[i], p.key_off.data(), p.key_off.size() * 4);
    memcpy(host + o_pin + b->param_off[i] * 8, params[i], (size_t)b->nparams[i] * 8);
  };
  for (int i = 0; i < n; ++i) stage_window(i);
  b->upload_bytes = upload;
  if (ws) {
    cudaEventRecord(ws->ev[0], stream);
    CUDA_TRY_OR(cudaMemcpyAsync(b->d_pool, host, upload, cudaMemcpyHostToDevice, stream), { slslam_lba_batch_destroy(b); return SLSLAM_ERR_CUDA; });
  } else {
    CUDA_TRY_OR(cudaMemcpy(b->d_pool, host, upload, cudaMemcpyHostToDevice), { slslam_lba_batch_destroy(b); return SLSLAM_ERR_CUDA; });
  }
  if (ws) { g_timing[0] = t_planned - t_begin; g_timing[1] = now_ms() - t_planned; }
  *out = b;
  return SLSLAM_OK;
}


#include "lba_device_plan.inl"

static int batch_create_impl(int32_t n, const slslam_lba_desc* descs, const double* const* params, int32_t device,
                             int32_t cluster_size, Workspace* ws, cudaStream_t stream, slslam_lba_batch** out) {
  if (!getenv("SLSLAM_HOST_PLAN")) {
    const int rc = batch_create_device_plan(n, descs, params, device, cluster_size, ws, stream, out);
    if (rc != SLSLAM_PLAN_FALLBACK) return rc;
  }
  return batch_create_host_plan(n, descs, params, device, cluster_size, ws, stream, out);
}

}  // namespace slslam

#include "moba_host.inl"
#include "wide_host.inl"

extern "C" {

void slslam_lba_get_limits(slslam_lba_limits* out) {
  if (!out) return;
  out->max_cameras = MAX_CAMS; out->max_free_cameras = MAX_FREE_CAMS; out->max_observations_per_line = 32;
  out->max_cluster_size = MAX_G;
  out->max_free_cameras_general = WIDE_MAX_FREE;
}

int slslam_lba_plan_check(int32_t n, const slslam_lba_desc* descs, const double* const* params, int32_t cluster_size,
                          int32_t* detail) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  // Planner parity: the device-built plan (lba_plan_kernel.cuh) against the host planner (build_plan) for the same group
  // size, array by array, bit for bit.  Returns 0 when identical, 1 when the device planner asked for the host
  // fallback, 2 on a mismatch (detail[0] = window, detail[1] = field, detail[2] = index), < 0 on errors.
  if (detail) detail[0] = detail[1] = detail[2] = -1;
  slslam_lba_batch* b = nullptr;
  int rc = batch_create_device_plan(n, descs, params, -1, cluster_size, nullptr, nullptr, &b);
  if (rc == SLSLAM_PLAN_FALLBACK) return 1;
  if (rc != SLSLAM_OK) return rc;
  std::vector<WindowPlan> plans;
  rc = build_plans(n, descs, b->CS, plans);
  if (rc != SLSLAM_OK) { slslam_lba_batch_destroy(b); return rc; }
  int result = 0;
  auto fail = [&](int w, int field, long long idx) { if (result == 0) { result = 2; if (detail) { detail[0] = w; detail[1] = field; detail[2] = (int)idx; } } };
  for (int i = 0; i < n && result == 0; ++i) {
    const WindowPlan& hp = plans[i];
    const PlanInfo& pi = b->dp_info[i];
    const auto& q = b->dp[i];
    WinHdr h;
    cudaMemcpy(&h, b->d_pool + q.hdr, sizeof(h), cudaMemcpyDeviceToHost);
    if (pi.Cf != hp.Cf || h.Cf != hp.Cf || h.nkeys != hp.nkeys || h.n != 6 * hp.Cf || h.vlen != lba_vlen(hp.Cf)) fail(i, 1, 0);
    if (pi.max_lines_cta != hp.max_lines_cta) fail(i, 2, 0);
    if (pi.max_slots_cta != hp.max_slots_cta) fail(i, 3, 0);
    if (pi.max_items_cta != hp.max_items_cta) fail(i, 4, 0);
    if ((pi.has_unobserved != 0) != hp.has_unobserved_blocks) fail(i, 5, 0);
    for (int r = 0; r <= MAX_G; ++r) {
      if (h.cta_slot_off[r] != hp.cta_slot_off[r]) fail(i, 6, r);
      if (h.cta_line_off[r] != hp.cta_line_off[r]) fail(i, 7, r);
    }
    for (int c = 0; c < MAX_CAMS; ++c) if (h.cam_free[c] != hp.cam_free[c]) fail(i, 8, c);
    if (result) break;
    const size_t ns = hp.meta.size();
    if ((size_t)pi.nslots != ns) { fail(i, 9, 0); break; }
    std::vector<int2> meta(ns);
    std::vector<double> obs(ns * 8);
    cudaMemcpy(meta.data(), b->d_pool + q.meta, ns * sizeof(int2), cudaMemcpyDeviceToHost);
    cudaMemcpy(obs.data(), b->d_pool + q.obs, ns * 64, cudaMemcpyDeviceToHost);
    for (size_t k = 0; k < ns; ++k) {
      if (meta[k].x != hp.meta[k].x || meta[k].y != hp.meta[k].y) { fail(i, 10, (long long)k); break; }
      const int j = hp.slot_src[k];
      for (int e = 0; e < 8; ++e) {
        const double want = j >= 0 ? descs[i].observations[8 * (size_t)j + e] : 0.0;
        if (memcmp(&want, &obs[8 * k + e], 8) != 0) { fail(i, 11, (long long)k); break; }
      }
      if (result) break;
    }
    if (result) break;
    std::vector<int> gid(hp.line_gid.size());
    cudaMemcpy(gid.data(), b->d_pool + q.gid, gid.size() * 4, cudaMemcpyDeviceToHost);
    for (size_t k = 0; k < gid.size(); ++k) if (gid[k] != hp.line_gid[k]) { fail(i, 12, (long long)k); break; }
    if ((size_t)pi.nitems != hp.items.size()) { fail(i, 13, 0); break; }
    std::vector<uint32_t> items(hp.items.size());
    cudaMemcpy(items.data(), b->d_pool + q.items, items.size() * 4, cudaMemcpyDeviceToHost);
    for (size_t k = 0; k < items.size(); ++k) if (items[k] != hp.items[k]) { fail(i, 14, (long long)k); break; }
    std::vector<int> koff(hp.key_off.size());
    cudaMemcpy(koff.data(), b->d_pool + q.koff, koff.size() * 4, cudaMemcpyDeviceToHost);
    for (size_t k = 0; k < koff.size(); ++k) if (koff[k] != hp.key_off[k]) { fail(i, 15, (long long)k); break; }
  }
  slslam_lba_batch_destroy(b);
  return result;
}

int slslam_lba_launch_shape(int32_t num_windows, int32_t max_observations, int32_t max_lines, int32_t resident_ctas,
                            int32_t smem_bytes_per_cta, int32_t* ctas_per_window, int32_t* windows_per_wave) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  // The planner's choice of group size and wave size as a pure function (no device): what batch creation uses with
  // resident_ctas = SMs x resident CTAs per SM and smem_bytes_per_cta = the opt-in maximum.
  if (num_windows <= 0 || max_observations < 0 || max_lines < 0 || resident_ctas <= 0 || smem_bytes_per_cta <= 0) return SLSLAM_ERR_INVALID;
  const int cs = pick_group_size_for(resident_ctas, num_windows, max_observations, 0, min_group_size_for_lines(max_lines, smem_bytes_per_cta));
  if (cs > resident_ctas) return SLSLAM_ERR_UNSUPPORTED;
  if (ctas_per_window) *ctas_per_window = cs;
  if (windows_per_wave) *windows_per_wave = balanced_wave(num_windows, resident_ctas / cs);
  return SLSLAM_OK;
}

int slslam_lba_batch_create(int32_t n, const slslam_lba_desc* descs, const double* const* params, int32_t device,
                            int32_t cluster_size, slslam_lba_batch** out) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  return batch_create_impl(n, descs, params, device, cluster_size, nullptr, nullptr, out);
}

int slslam_lba_batch_upload_params(slslam_lba_batch* b, const double* const* params, void* cuda_stream) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!b || !params) return SLSLAM_ERR_INVALID;
  cudaSetDevice(b->device);
  for (int i = 0; i < b->n; ++i) {
    if (!params[i]) return SLSLAM_ERR_INVALID;
    memcpy(b->h_params + b->param_off[i], params[i], (size_t)b->nparams[i] * 8);
  }
  CUDA_TRY(cudaMemcpyAsync(b->d_params_in, b->h_params, b->total_params * 8, cudaMemcpyHostToDevice, (cudaStream_t)cuda_stream));
  return SLSLAM_OK;
}

int slslam_lba_batch_solve(slslam_lba_batch* b, void* cuda_stream) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!b) return SLSLAM_ERR_INVALID;
  cudaSetDevice(b->device);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  bool need_copy = false;
  for (const auto& p : b->plans) need_copy = need_copy || p.has_unobserved_blocks;
  if (need_copy && !b->inplace) CUDA_TRY(cudaMemcpyAsync(b->d_params_out, b->d_params_in, b->total_params * 8, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemsetAsync(b->d_bar, 0, b->bar_bytes, st));
  // cooperative launches (every CTA of a group must be resident: the group barrier spins), `max_active` windows per wave
  SmemLayout lay = b->lay;
  for (int w0 = 0; w0 < b->n; w0 += b->max_active) {
    const int nw = std::min(b->max_active, b->n - w0);
    const WinHdr* hdrs = b->d_hdrs + w0;
    void* args[2] = {(void*)&hdrs, (void*)&lay};
    CUDA_TRY(cudaLaunchCooperativeKernel((const void*)lba_solve_kernel, dim3((unsigned)(nw * b->CS)), dim3(LBA_NT), args, b->smem_bytes, st));
  }
  return SLSLAM_OK;
}

int slslam_lba_batch_download(slslam_lba_batch* b, void* cuda_stream, double* const* params_out,
                              slslam_summary* summaries_out, double* const* trace_out) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!b) return SLSLAM_ERR_INVALID;
  cudaSetDevice(b->device);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (params_out) CUDA_TRY(cudaMemcpyAsync(b->h_params, b->d_params_out, b->total_params * 8, cudaMemcpyDeviceToHost, st));
  std::vector<double> tr;
  if (trace_out) { tr.resize(b->total_trace); CUDA_TRY(cudaMemcpyAsync(tr.data(), b->d_trace, b->total_trace * 8, cudaMemcpyDeviceToHost, st)); }
  if (summaries_out) CUDA_TRY(cudaMemcpyAsync(summaries_out, b->d_summ, sizeof(slslam_summary) * b->n, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  for (int i = 0; i < b->n; ++i) {
    if (params_out && params_out[i]) memcpy(params_out[i], b->h_params + b->param_off[i], (size_t)b->nparams[i] * 8);
    if (trace_out && trace_out[i]) memcpy(trace_out[i], tr.data() + b->trace_off[i], (size_t)b->plans[i].max_iters * SLSLAM_TRACE_WIDTH * 8);
  }
  return SLSLAM_OK;
}

int slslam_lba_batch_max_active_clusters(const slslam_lba_batch* b, int32_t* max_active) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!b || !max_active) return SLSLAM_ERR_INVALID;
  *max_active = b->max_active;
  return SLSLAM_OK;
}

int slslam_lba_batch_info(const slslam_lba_batch* b, int32_t* cluster_size, int32_t* threads_per_cta,
                          int32_t* smem_bytes_per_cta, int32_t* z_in_smem) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!b) return SLSLAM_ERR_INVALID;
  if (cluster_size) *cluster_size = b->CS;
  if (threads_per_cta) *threads_per_cta = LBA_NT;
  if (smem_bytes_per_cta) *smem_bytes_per_cta = (int32_t)b->smem_bytes;
  if (z_in_smem) *z_in_smem = b->lay.z_in_smem;
  return SLSLAM_OK;
}

int slslam_lba_batch_plan_cycles(const slslam_lba_batch* b, int32_t window, int32_t* cycles8) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!b || !cycles8 || window < 0 || window >= b->n || !b->device_planned) return SLSLAM_ERR_INVALID;
  for (int k = 0; k < 8; ++k) cycles8[k] = b->dp_info[(size_t)window].phase_cycles[k];
  return SLSLAM_OK;
}

int slslam_lba_batch_phase_cycles(slslam_lba_batch* b, void* cuda_stream, int32_t window, int64_t* cycles_out, int32_t n) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!b || !cycles_out || window < 0 || window >= b->n) return SLSLAM_ERR_INVALID;
  cudaSetDevice(b->device);
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)cuda_stream));
  long long tmp[NPHASE];
  CUDA_TRY(cudaMemcpy(tmp, b->d_phase + (size_t)NPHASE * window, sizeof(tmp), cudaMemcpyDeviceToHost));
  for (int k = 0; k < n && k < NPHASE; ++k) cycles_out[k] = tmp[k];
  return SLSLAM_OK;
}

int slslam_lba_batch_transfer_bytes(const slslam_lba_batch* b, int64_t* h2d_bytes, int64_t* d2h_bytes) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!b) return SLSLAM_ERR_INVALID;
  if (h2d_bytes) *h2d_bytes = (int64_t)b->upload_bytes;
  if (d2h_bytes) *d2h_bytes = (int64_t)(b->total_params * 8 + sizeof(slslam_summary) * (size_t)b->n);
  return SLSLAM_OK;
}

void slslam_lba_batch_destroy(slslam_lba_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->borrowed && b->ws) b->ws->plans.swap(b->plans);
  if (!b->borrowed) {
    if (b->d_pool) cudaFree(b->d_pool);
    if (b->h_params) cudaFreeHost(b->h_params);
  }
  delete b;
}

int slslam_lba_route(const slslam_lba_desc* desc) {
  slslam::set_last_error("");
  if (!desc) return SLSLAM_ERR_INVALID;
  const slslam_lba_desc& d = *desc;
  if (d.num_cameras < 0 || d.num_lines < 0 || d.num_observations < 0 ||
      (d.num_observations > 0 && (!d.camera_index || !d.line_index || !d.fixed_index))) return SLSLAM_ERR_INVALID;
  int fc = 0, nf = 0;
  if (d.observations && validate_desc(d) == SLSLAM_OK && d.max_iterations > 0 && moba_candidate(d, &fc, &nf)) return SLSLAM_ROUTE_MOTION_ONLY;
  WidePlan wp;
  bool wide = false;
  int rc = wide_plan(d, wp, &wide);
  if (rc != SLSLAM_OK) return rc;
  if (!wide) return SLSLAM_ROUTE_TILED;
  rc = wide_plan_tables(d, wp);
  return rc == SLSLAM_OK ? SLSLAM_ROUTE_GENERAL : rc;
}

int slslam_lba_solve_batch(int32_t n, const slslam_lba_desc* descs, double* const* params_inout, slslam_summary* summaries_out) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (n <= 0 || !descs || !params_inout) return SLSLAM_ERR_INVALID;
  // motion-only BA (one free camera, constant lines) in every window: the dedicated kernel, no plan
  if (!getenv("SLSLAM_NO_MOBA_FASTPATH")) {
    std::vector<int> fc((size_t)n), nf((size_t)n);
    bool all = true;
    for (int i = 0; i < n && all; ++i) {
      if (validate_desc(descs[i]) != SLSLAM_OK || !params_inout[i] || descs[i].max_iterations <= 0) { all = false; break; }
      const int np = 6 * descs[i].num_cameras + 4 * descs[i].num_lines;
      for (int k = 0; k < np && all; ++k) if (!std::isfinite(params_inout[i][k])) all = false;
      all = all && moba_candidate(descs[i], &fc[i], &nf[i]);
    }
    if (all) return moba_solve_batch(n, descs, params_inout, summaries_out, fc.data(), nf.data());
  }
  // windows beyond the tiled kernel's limits (more than 32 camera blocks / 24 free cameras / 32 observations of a line:
  // the reference's --ba_window_size 20 / 40 shapes) go to the general kernel, the whole batch with them
  {
    bool any_wide = false, plannable = true;
    std::vector<WidePlan> wplans((size_t)n);
    for (int i = 0; i < n && plannable; ++i) {
      const slslam_lba_desc& d = descs[i];
      if (d.num_cameras < 0 || d.num_lines < 0 || d.num_observations < 0 || d.max_iterations < 0 || !params_inout[i] ||
          (d.num_observations > 0 && (!d.camera_index || !d.line_index || !d.fixed_index || !d.observations))) { plannable = false; break; }
      if (d.num_cameras <= MAX_CAMS) {
        // cheap screen first: only windows with many cameras or long lines can need the general kernel
        bool maybe = false;
        if (d.num_cameras > MAX_FREE_CAMS || d.num_observations > 32 * std::max(1, d.num_lines)) maybe = true;
        if (!maybe) {
          // a line with more than 32 observations needs more observations than 32 in total
          if (d.num_observations > 32) {
            std::vector<int> cnt((size_t)std::max(1, d.num_lines), 0);
            for (int k = 0; k < d.num_observations && !maybe; ++k) {
              const int l = d.line_index[k];
              if (l >= 0 && l < d.num_lines && ++cnt[l] > 32) maybe = true;
            }
          }
        }
        if (!maybe) continue;
      }
      bool w = false;
      const int prc = wide_plan(d, wplans[i], &w);
      if (prc != SLSLAM_OK) return prc;
      any_wide = any_wide || w;
    }
    if (plannable && any_wide) {
      for (int i = 0; i < n; ++i) {
        const int np = 6 * descs[i].num_cameras + 4 * descs[i].num_lines;
        for (int k = 0; k < np; ++k) if (!std::isfinite(params_inout[i][k])) return SLSLAM_ERR_NUMERICAL;
        if (wplans[i].N != descs[i].num_observations || wplans[i].order.empty()) {      // windows the screen skipped
          bool w = false;
          const int prc = wide_plan(descs[i], wplans[i], &w);
          if (prc != SLSLAM_OK) return prc;
        }
      }
      return wide_solve_batch(n, descs, params_inout, summaries_out, wplans);
    }
  }
  // plan -> pinned staging -> H2D -> one cluster launch -> D2H, all on the calling thread's cached workspace
  slslam_lba_batch* b = nullptr;
  const double t0 = now_ms();
  // First attempt: planner and solve kernel enqueued back to back, ONE synchronisation (the planner's flags come back
  // with the results).  A batch the planner could not handle that way is solved again through the checked path.
  int rc = SLSLAM_PLAN_FALLBACK;
  if (!getenv("SLSLAM_HOST_PLAN") && !getenv("SLSLAM_NO_DEFERRED_PLAN"))
    rc = batch_create_device_plan(n, descs, (const double* const*)params_inout, -1, 0, &g_ws, nullptr, &b, false, nullptr, true);
  if (rc == SLSLAM_PLAN_FALLBACK) rc = batch_create_impl(n, descs, (const double* const*)params_inout, -1, 0, &g_ws, nullptr, &b);
  if (rc != SLSLAM_OK) return rc;
  const double t1 = now_ms();
  double t2 = t1;
  cudaEventRecord(g_ws.ev[1], nullptr);
  rc = slslam_lba_batch_solve(b, nullptr);
  if (rc == SLSLAM_OK) {
    slslam_summary* h_summ = (slslam_summary*)(b->h_params + b->total_params);
    PlanInfo* h_info = (PlanInfo*)(g_ws.h_res + ((b->total_params * 8 + sizeof(slslam_summary) * n + 255) & ~(size_t)255));
    cudaEventRecord(g_ws.ev[2], nullptr);
    cudaError_t e = cudaMemcpyAsync(b->h_params, b->d_params_out, b->total_params * 8, cudaMemcpyDeviceToHost, nullptr);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_summ, b->d_summ, sizeof(slslam_summary) * n, cudaMemcpyDeviceToHost, nullptr);
    if (e == cudaSuccess && b->deferred) e = cudaMemcpyAsync(h_info, b->d_info, sizeof(PlanInfo) * n, cudaMemcpyDeviceToHost, nullptr);
    if (e == cudaSuccess) e = cudaEventRecord(g_ws.ev[3], nullptr);
    if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
    if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); cudaGetLastError(); rc = SLSLAM_ERR_CUDA; }
    if (rc == SLSLAM_OK && b->deferred) {
      const int chk = device_plan_check_deferred(b, h_info);
      if (chk == SLSLAM_PLAN_FALLBACK) {
        // (rare: a camera observing a line twice, a shape that needs the group-size search) nothing was solved: again,
        // through the path that checks the plan before it launches
        slslam_lba_batch_destroy(b);
        b = nullptr;
        rc = batch_create_impl(n, descs, (const double* const*)params_inout, -1, 0, &g_ws, nullptr, &b);
        if (rc != SLSLAM_OK) return rc;
        rc = slslam_lba_batch_solve(b, nullptr);
        if (rc == SLSLAM_OK) {
          h_summ = (slslam_summary*)(b->h_params + b->total_params);
          e = cudaMemcpyAsync(b->h_params, b->d_params_out, b->total_params * 8, cudaMemcpyDeviceToHost, nullptr);
          if (e == cudaSuccess) e = cudaMemcpyAsync(h_summ, b->d_summ, sizeof(slslam_summary) * n, cudaMemcpyDeviceToHost, nullptr);
          if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
          if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); cudaGetLastError(); rc = SLSLAM_ERR_CUDA; }
        }
      } else if (chk != SLSLAM_OK) {
        rc = chk;
      }
    }
    t2 = now_ms();
    if (rc == SLSLAM_OK) {
      float ms = 0.f;
      for (int k = 0; k < 3; ++k) { cudaEventElapsedTime(&ms, g_ws.ev[k], g_ws.ev[k + 1]); g_timing[5 + k] = ms; }
    }
    if (rc == SLSLAM_OK) {
      // parameters are only overwritten once the whole batch has succeeded
      for (int i = 0; i < n; ++i) {
        memcpy(params_inout[i], b->h_params + b->param_off[i], (size_t)b->nparams[i] * 8);
        if (summaries_out) summaries_out[i] = h_summ[i];
      }
    }
  }
  slslam_lba_batch_destroy(b);
  const double t3 = now_ms();
  g_timing[2] = t2 - t1; g_timing[3] = t3 - t2; g_timing[4] = t3 - t0;
  return rc;
}

int slslam_lba_solve_batch_device(int32_t n, const slslam_lba_desc* descs, double* const* params_dev_inout,
                                  slslam_summary* summaries_dev_out, slslam_summary* summaries_host_out, void* cuda_stream) {
  slslam::set_last_error("");
  if (n <= 0 || !descs || !params_dev_inout) return SLSLAM_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  slslam_lba_batch* b = nullptr;
  const double t0 = now_ms();
  // With host summaries requested the call waits for the solve anyway: the planner's flags are then read back together
  // with them (one synchronisation); without, the plan is checked before the launch so that errors are still reported.
  const bool deferred = summaries_host_out != nullptr && !getenv("SLSLAM_NO_DEFERRED_PLAN");
  static const char* kNeedsHostPlan = "window shape needs the host planner (a camera observing a line twice, or a group-size search), which cannot read device-resident inputs";
  int rc = batch_create_device_plan(n, descs, (const double* const*)params_dev_inout, -1, 0, &g_ws, st, &b, true, summaries_dev_out, deferred);
  if (rc == SLSLAM_PLAN_FALLBACK) { set_last_error(kNeedsHostPlan); return SLSLAM_ERR_UNSUPPORTED; }
  if (rc != SLSLAM_OK) return rc;
  const double t1 = now_ms();
  rc = slslam_lba_batch_solve(b, st);
  if (rc == SLSLAM_OK && summaries_host_out) {
    // the summaries are the only thing read back; without this the call returns as soon as the solve is enqueued
    slslam_summary* h_summ = (slslam_summary*)(b->h_params);
    PlanInfo* h_info = (PlanInfo*)(g_ws.h_res + ((b->total_params * 8 + sizeof(slslam_summary) * n + 255) & ~(size_t)255));
    cudaError_t e = cudaMemcpyAsync(h_summ, summaries_dev_out ? summaries_dev_out : b->d_summ, sizeof(slslam_summary) * n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && b->deferred) e = cudaMemcpyAsync(h_info, b->d_info, sizeof(PlanInfo) * n, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); cudaGetLastError(); rc = SLSLAM_ERR_CUDA; }
    else if (b->deferred && (rc = device_plan_check_deferred(b, h_info)) != SLSLAM_OK) {
      if (rc == SLSLAM_PLAN_FALLBACK) { set_last_error(kNeedsHostPlan); rc = SLSLAM_ERR_UNSUPPORTED; }
    }
    else memcpy(summaries_host_out, h_summ, sizeof(slslam_summary) * n);
  }
  slslam_lba_batch_destroy(b);
  g_timing[2] = now_ms() - t1; g_timing[3] = 0; g_timing[4] = now_ms() - t0;
  return rc;
}

#include "lba_pipeline.inl"

void slslam_lba_last_timings(double* ms5) {   // 8 values, see the header
  if (ms5) for (int k = 0; k < 8; ++k) ms5[k] = g_timing[k];
}

int slslam_lba_solve(const slslam_lba_desc* desc, double* params_inout, slslam_summary* summary_out) {
  if (!desc || !params_inout) return SLSLAM_ERR_INVALID;
  double* pp[1] = {params_inout};
  return slslam_lba_solve_batch(1, desc, pp, summary_out);
}

int slslam_lba_evaluate(const slslam_lba_desc* desc, const double* params, double* residuals, double* jac_camera,
                        double* jac_line, double* cost_out) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!desc || !params || !residuals) return SLSLAM_ERR_INVALID;
  int rc = validate_desc(*desc);
  if (rc != SLSLAM_OK) return rc;
  rc = ensure_device(-1);
  if (rc != SLSLAM_OK) return rc;
  const int N = desc->num_observations, C = desc->num_cameras, L = desc->num_lines;
  if (N == 0) { if (cost_out) *cost_out = 0.0; return SLSLAM_OK; }
  const size_t np = (size_t)6 * C + 4 * L;
  char* pool = nullptr;
  size_t off = 0;
  auto reserve = [&](size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; };
  const size_t o_ci = reserve(4 * (size_t)N), o_li = reserve(4 * (size_t)N), o_ob = reserve(64 * (size_t)N), o_p = reserve(8 * np);
  const size_t o_r = reserve(32 * (size_t)N), o_jc = reserve(192 * (size_t)N), o_jl = reserve(128 * (size_t)N), o_c = reserve(8);
  CUDA_TRY(cudaMalloc((void**)&pool, off));
  rc = SLSLAM_OK;
#define EV_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_last_error(cudaGetErrorString(e_)); rc = SLSLAM_ERR_CUDA; goto done; } } while (0)
  EV_TRY(cudaMemcpy(pool + o_ci, desc->camera_index, 4 * (size_t)N, cudaMemcpyHostToDevice));
  EV_TRY(cudaMemcpy(pool + o_li, desc->line_index, 4 * (size_t)N, cudaMemcpyHostToDevice));
  EV_TRY(cudaMemcpy(pool + o_ob, desc->observations, 64 * (size_t)N, cudaMemcpyHostToDevice));
  EV_TRY(cudaMemcpy(pool + o_p, params, 8 * np, cudaMemcpyHostToDevice));
  EV_TRY(cudaMemset(pool + o_c, 0, 8));
  lba_evaluate_kernel<<<(N + 127) / 128, 128>>>(N, C, (const int*)(pool + o_ci), (const int*)(pool + o_li), (const double*)(pool + o_ob),
                                               (const double*)(pool + o_p), desc->baseline >= 0 ? desc->baseline : 0.12,
                                               desc->huber_delta > 0 ? desc->huber_delta : 1.0 / 406.05, desc->robust,
                                               (double*)(pool + o_r), (double*)(pool + o_jc), (double*)(pool + o_jl), (double*)(pool + o_c));
  EV_TRY(cudaGetLastError());
  EV_TRY(cudaMemcpy(residuals, pool + o_r, 32 * (size_t)N, cudaMemcpyDeviceToHost));
  if (jac_camera) EV_TRY(cudaMemcpy(jac_camera, pool + o_jc, 192 * (size_t)N, cudaMemcpyDeviceToHost));
  if (jac_line) EV_TRY(cudaMemcpy(jac_line, pool + o_jl, 128 * (size_t)N, cudaMemcpyDeviceToHost));
  if (cost_out) EV_TRY(cudaMemcpy(cost_out, pool + o_c, 8, cudaMemcpyDeviceToHost));
done:
#undef EV_TRY
  cudaFree(pool);
  return rc;
}

}  // extern "C"
