// Host side of the device-resident map (include/slslam_b200.h: slslam_map_*): allocation, appends, and the per-keyframe
// bundle adjustment -- window assembly kernels, the LBA solve on the assembled (device-resident) window through
// slslam_lba_solve_batch_device, write-back kernels.  Replaces the host loops of SLAM::bundle_adjustment around
// ceres::Solve (reference src/slam.cpp:799-920 and 957-972); the keyframe graph, metric_embedding and the choice of the
// window keyframes (slam.cpp:1317-1382) stay with the caller, who passes the window's keyframes with their graph-distance
// ranks and their re-anchored poses.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>

#include "common_host.h"
#include "map_kernels.cuh"

using namespace slslam;

struct slslam_map {
  int device = 0;
  MapDev d;
  char* pool = nullptr;
  char* h_pin = nullptr; size_t h_cap = 0;       // staging of the small per-call arrays
  int* h_sizes = nullptr;                        // pinned: L, N
  std::vector<int2> kf_range;                    // per keyframe id: (first observation, count), count < 0: unknown id
  std::vector<char> lm_known;
  int num_obs = 0;
  int cand_cap = 0;
  double* w_params0 = nullptr;                   // the assembled parameters before the solve (diagnostics / parity)
  int* d_small = nullptr; size_t small_cap = 0;  // device copy of the per-call arrays
  int last_C = 0, last_L = 0, last_N = 0;
  std::vector<int> last_cam_kf;
  slslam_map_timings tm;
};

namespace {
size_t up256(size_t b) { return (b + 255) & ~(size_t)255; }
}

extern "C" {

int slslam_map_create(int32_t device, int32_t max_keyframes, int32_t max_landmarks, int32_t max_observations, slslam_map** out) {
  set_last_error("");
  if (!out || max_keyframes <= 0 || max_landmarks <= 0 || max_observations <= 0) return SLSLAM_ERR_INVALID;
  *out = nullptr;
  int rc = ensure_device(device);
  if (rc != SLSLAM_OK) return rc;
  slslam_map* m = new (std::nothrow) slslam_map();
  if (!m) return SLSLAM_ERR_INVALID;
  cudaGetDevice(&m->device);
  memset(&m->d, 0, sizeof(m->d));
  m->d.max_kf = max_keyframes; m->d.max_lm = max_landmarks; m->d.max_obs = max_observations;
  // the window can hold every observation of the map (a window never has more candidates than that)
  m->cand_cap = max_observations;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t r = off; off += up256(bytes); return r; };
  const size_t K = (size_t)max_keyframes, Lm = (size_t)max_landmarks, O = (size_t)max_observations, Cc = (size_t)m->cand_cap;
  const size_t o_T = take(96 * K), o_cam = take(4 * K), o_free = take(4 * K), o_line = take(48 * Lm), o_init = take(4 * Lm),
               o_cnt = take(4 * Lm), o_lidx = take(4 * Lm), o_olm = take(4 * O), o_okf = take(4 * O), o_oxy = take(64 * O),
               o_keep = take(4 * Cc), o_kpos = take(4 * Cc), o_sizes = take(64), o_wcam = take(4 * Cc), o_wline = take(4 * Cc),
               o_wfix = take(8 * Cc), o_wobs = take(64 * Cc), o_wpar = take(8 * (6 * K + 4 * Lm)), o_wpar0 = take(8 * (6 * K + 4 * Lm)),
               o_wlm = take(4 * Lm);
  if (cudaMalloc((void**)&m->pool, off) != cudaSuccess) { set_last_error("cudaMalloc of the map pool failed"); cudaGetLastError(); delete m; return SLSLAM_ERR_CUDA; }
  cudaMemset(m->pool, 0, off);
  char* B = m->pool;
  m->d.kf_T = (double*)(B + o_T); m->d.kf_cam = (int*)(B + o_cam); m->d.kf_free = (int*)(B + o_free);
  m->d.lm_line = (double*)(B + o_line); m->d.lm_init_kf = (int*)(B + o_init); m->d.lm_count = (int*)(B + o_cnt);
  m->d.lm_line_index = (int*)(B + o_lidx); m->d.obs_lm = (int*)(B + o_olm); m->d.obs_kf = (int*)(B + o_okf);
  m->d.obs_xy = (double*)(B + o_oxy); m->d.keep = (int*)(B + o_keep); m->d.keep_pos = (int*)(B + o_kpos);
  m->d.sizes = (int*)(B + o_sizes); m->d.w_cam = (int*)(B + o_wcam); m->d.w_line = (int*)(B + o_wline);
  m->d.w_fixed = (int*)(B + o_wfix); m->d.w_obs = (double*)(B + o_wobs); m->d.w_params = (double*)(B + o_wpar);
  m->w_params0 = (double*)(B + o_wpar0); m->d.w_line_lm = (int*)(B + o_wlm);
  cudaMemset(m->d.kf_cam, 0xff, 4 * K);
  cudaMemset(m->d.lm_init_kf, 0xff, 4 * Lm);
  cudaMemset(m->d.lm_line_index, 0xff, 4 * Lm);
  m->h_cap = 1 << 20;
  if (cudaMallocHost((void**)&m->h_pin, m->h_cap) != cudaSuccess || cudaMallocHost((void**)&m->h_sizes, 64) != cudaSuccess) {
    set_last_error("cudaMallocHost failed"); cudaGetLastError(); cudaFree(m->pool); delete m; return SLSLAM_ERR_CUDA;
  }
  m->small_cap = 1 << 20;
  if (cudaMalloc((void**)&m->d_small, m->small_cap) != cudaSuccess) { cudaGetLastError(); cudaFree(m->pool); delete m; return SLSLAM_ERR_CUDA; }
  m->kf_range.assign(K, make_int2(0, -1));
  m->lm_known.assign(Lm, 0);
  cudaDeviceSynchronize();
  *out = m;
  return SLSLAM_OK;
}

void slslam_map_destroy(slslam_map* m) {
  if (!m) return;
  cudaSetDevice(m->device);
  if (m->pool) cudaFree(m->pool);
  if (m->d_small) cudaFree(m->d_small);
  if (m->h_pin) cudaFreeHost(m->h_pin);
  if (m->h_sizes) cudaFreeHost(m->h_sizes);
  delete m;
}

static int map_stage(slslam_map* m, size_t bytes) {
  if (bytes <= m->h_cap) return SLSLAM_OK;
  cudaFreeHost(m->h_pin); m->h_pin = nullptr; m->h_cap = 0;
  CUDA_TRY(cudaMallocHost((void**)&m->h_pin, bytes * 2));
  m->h_cap = bytes * 2;
  return SLSLAM_OK;
}

// A new keyframe: its pose and ALL its observations (reference: add_new_keyframe, src/slam.cpp:730-761; the landmark's
// obs_vec grows by one entry per observing keyframe, here the entries of one keyframe are one contiguous range).
int slslam_map_add_keyframe(slslam_map* m, int32_t kf_id, const double* T12, int32_t n_obs, const int32_t* lm_ids, const double* obs8) {
  set_last_error("");
  if (!m || !T12 || kf_id < 0 || kf_id >= m->d.max_kf || n_obs < 0 || (n_obs > 0 && (!lm_ids || !obs8))) return SLSLAM_ERR_INVALID;
  if (m->kf_range[kf_id].y >= 0) { set_last_error("keyframe id already in the map"); return SLSLAM_ERR_INVALID; }
  if (m->num_obs + n_obs > m->d.max_obs) { set_last_error("map observation capacity exceeded"); return SLSLAM_ERR_UNSUPPORTED; }
  for (int i = 0; i < n_obs; ++i) if (lm_ids[i] < 0 || lm_ids[i] >= m->d.max_lm) return SLSLAM_ERR_INVALID;
  cudaSetDevice(m->device);
  const size_t bytes = 96 + (size_t)n_obs * (4 + 4 + 64) + 256;
  int rc = map_stage(m, bytes);
  if (rc != SLSLAM_OK) return rc;
  // one staging buffer, four small copies (the destinations are different arrays)
  char* h = m->h_pin;
  memcpy(h, T12, 96);
  int* hl = (int*)(h + 96); int* hk = hl + n_obs; double* ho = (double*)(h + 96 + up256((size_t)8 * n_obs));
  for (int i = 0; i < n_obs; ++i) { hl[i] = lm_ids[i]; hk[i] = kf_id; }
  if (n_obs) memcpy(ho, obs8, (size_t)64 * n_obs);
  CUDA_TRY(cudaMemcpyAsync(m->d.kf_T + 12 * (size_t)kf_id, h, 96, cudaMemcpyHostToDevice, nullptr));
  if (n_obs) {
    CUDA_TRY(cudaMemcpyAsync(m->d.obs_lm + m->num_obs, hl, (size_t)4 * n_obs, cudaMemcpyHostToDevice, nullptr));
    CUDA_TRY(cudaMemcpyAsync(m->d.obs_kf + m->num_obs, hk, (size_t)4 * n_obs, cudaMemcpyHostToDevice, nullptr));
    CUDA_TRY(cudaMemcpyAsync(m->d.obs_xy + 8 * (size_t)m->num_obs, ho, (size_t)64 * n_obs, cudaMemcpyHostToDevice, nullptr));
  }
  CUDA_TRY(cudaStreamSynchronize(nullptr));      // the staging buffer is reused by the next call
  m->kf_range[kf_id] = make_int2(m->num_obs, n_obs);
  m->num_obs += n_obs;
  return SLSLAM_OK;
}

// New landmarks: line = (closest point, direction) in the frame of init_kf (landmark_t::line, slam.cpp:190-219).
int slslam_map_add_landmarks(slslam_map* m, int32_t n, const int32_t* lm_ids, const int32_t* init_kf_ids, const double* line_av6) {
  set_last_error("");
  if (!m || n < 0 || (n > 0 && (!lm_ids || !init_kf_ids || !line_av6))) return SLSLAM_ERR_INVALID;
  for (int i = 0; i < n; ++i) {
    if (lm_ids[i] < 0 || lm_ids[i] >= m->d.max_lm || init_kf_ids[i] < 0 || init_kf_ids[i] >= m->d.max_kf) return SLSLAM_ERR_INVALID;
  }
  cudaSetDevice(m->device);
  // landmark ids of one call are usually consecutive: copy runs
  int i = 0;
  while (i < n) {
    int j = i + 1;
    while (j < n && lm_ids[j] == lm_ids[j - 1] + 1) ++j;
    CUDA_TRY(cudaMemcpy(m->d.lm_line + 6 * (size_t)lm_ids[i], line_av6 + 6 * (size_t)i, (size_t)48 * (j - i), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(m->d.lm_init_kf + lm_ids[i], init_kf_ids + i, (size_t)4 * (j - i), cudaMemcpyHostToDevice));
    for (int k = i; k < j; ++k) m->lm_known[lm_ids[k]] = 1;
    i = j;
  }
  return SLSLAM_OK;
}

int slslam_map_set_poses(slslam_map* m, int32_t n, const int32_t* kf_ids, const double* T12) {
  set_last_error("");
  if (!m || n < 0 || (n > 0 && (!kf_ids || !T12))) return SLSLAM_ERR_INVALID;
  cudaSetDevice(m->device);
  for (int i = 0; i < n; ++i) {
    if (kf_ids[i] < 0 || kf_ids[i] >= m->d.max_kf) return SLSLAM_ERR_INVALID;
    CUDA_TRY(cudaMemcpyAsync(m->d.kf_T + 12 * (size_t)kf_ids[i], T12 + 12 * (size_t)i, 96, cudaMemcpyHostToDevice, nullptr));
  }
  CUDA_TRY(cudaStreamSynchronize(nullptr));
  return SLSLAM_OK;
}

int slslam_map_get_poses(slslam_map* m, int32_t n, const int32_t* kf_ids, double* T12_out) {
  set_last_error("");
  if (!m || n < 0 || (n > 0 && (!kf_ids || !T12_out))) return SLSLAM_ERR_INVALID;
  cudaSetDevice(m->device);
  for (int i = 0; i < n; ++i) {
    if (kf_ids[i] < 0 || kf_ids[i] >= m->d.max_kf) return SLSLAM_ERR_INVALID;
    CUDA_TRY(cudaMemcpyAsync(T12_out + 12 * (size_t)i, m->d.kf_T + 12 * (size_t)kf_ids[i], 96, cudaMemcpyDeviceToHost, nullptr));
  }
  CUDA_TRY(cudaStreamSynchronize(nullptr));
  return SLSLAM_OK;
}

int slslam_map_get_landmarks(slslam_map* m, int32_t n, const int32_t* lm_ids, double* line_av6_out) {
  set_last_error("");
  if (!m || n < 0 || (n > 0 && (!lm_ids || !line_av6_out))) return SLSLAM_ERR_INVALID;
  cudaSetDevice(m->device);
  for (int i = 0; i < n; ++i) {
    if (lm_ids[i] < 0 || lm_ids[i] >= m->d.max_lm) return SLSLAM_ERR_INVALID;
    CUDA_TRY(cudaMemcpyAsync(line_av6_out + 6 * (size_t)i, m->d.lm_line + 6 * (size_t)lm_ids[i], 48, cudaMemcpyDeviceToHost, nullptr));
  }
  CUDA_TRY(cudaStreamSynchronize(nullptr));
  return SLSLAM_OK;
}

// What SLAM::bundle_adjustment does (src/slam.cpp:795-975) for the window keyframes `ba_kf_ids` with graph-distance
// ranks `ba_order` (ba_kfs, slam.cpp:1376-1382): keyframes with rank < window_size are free, the others constant.
int slslam_map_bundle_adjust(slslam_map* m, int32_t n_ba, const int32_t* ba_kf_ids, const int32_t* ba_order, int32_t window_size,
                             int32_t max_iterations, int32_t robust, slslam_summary* summary_out, int32_t* sizes3_out) {
  set_last_error("");
  if (!m || n_ba <= 0 || !ba_kf_ids || !ba_order || window_size <= 0 || max_iterations < 0) return SLSLAM_ERR_INVALID;
  cudaSetDevice(m->device);
  const double t_begin = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  // cameras: the free keyframes in keyframe-id order (the reference iterates its std::map by id, slam.cpp:809-832), then
  // the constant ones, also by id (the reference appends them in order of first appearance; a constant block's index
  // labels nothing the arithmetic sees, and one without observations in the window is never touched)
  std::vector<std::pair<int, int> > ks;
  for (int i = 0; i < n_ba; ++i) {
    const int id = ba_kf_ids[i];
    if (id < 0 || id >= m->d.max_kf || m->kf_range[id].y < 0) { set_last_error("window keyframe is not in the map"); return SLSLAM_ERR_INVALID; }
    ks.push_back(std::make_pair(id, ba_order[i]));
  }
  std::sort(ks.begin(), ks.end());
  for (size_t i = 1; i < ks.size(); ++i) if (ks[i].first == ks[i - 1].first) { set_last_error("duplicate window keyframe"); return SLSLAM_ERR_INVALID; }
  std::vector<int> cam_kf;
  for (auto& k : ks) if (k.second < window_size) cam_kf.push_back(k.first);
  const int Cfree = (int)cam_kf.size();
  for (auto& k : ks) if (k.second >= window_size) cam_kf.push_back(k.first);
  const int C = (int)cam_kf.size();
  // candidate observations: the ranges of the window keyframes, chronological (= by keyframe id)
  std::vector<int2> ranges;
  std::vector<int> cand_off(1, 0);
  for (auto& k : ks) {
    ranges.push_back(m->kf_range[k.first]);
    cand_off.push_back(cand_off.back() + m->kf_range[k.first].y);
  }
  const int n_cand = cand_off.back();
  const int nr = (int)ranges.size();
  // one small upload: ranges | cand_off | cam_kf
  const size_t bytes = 8 * (size_t)nr + 4 * (size_t)(nr + 1) + 4 * (size_t)C;
  if (bytes > m->small_cap || bytes > m->h_cap) return SLSLAM_ERR_UNSUPPORTED;
  char* h = m->h_pin;
  memcpy(h, ranges.data(), 8 * (size_t)nr);
  memcpy(h + 8 * (size_t)nr, cand_off.data(), 4 * (size_t)(nr + 1));
  memcpy(h + 8 * (size_t)nr + 4 * (size_t)(nr + 1), cam_kf.data(), 4 * (size_t)C);
  cudaStream_t st = nullptr;
  CUDA_TRY(cudaMemcpyAsync(m->d_small, h, bytes, cudaMemcpyHostToDevice, st));
  MapDev d = m->d;
  d.ranges = (const int2*)m->d_small; d.n_ranges = nr; d.n_cand = n_cand;
  d.cand_off = (const int*)((char*)m->d_small + 8 * (size_t)nr);
  const int* d_cam_kf = (const int*)((char*)m->d_small + 8 * (size_t)nr + 4 * (size_t)(nr + 1));
  d.C = C; d.Cfree = Cfree;
  // marks, counts, selection
  map_mark_kernel<<<(C + 127) / 128, 128, 0, st>>>(d, d_cam_kf);
  if (n_cand > 0) map_count_kernel<<<(n_cand + 255) / 256, 256, 0, st>>>(d);
  map_select_kernel<<<1, 1024, 0, st>>>(d);
  CUDA_TRY(cudaMemcpyAsync(m->h_sizes, d.sizes, 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const int L = m->h_sizes[0], N = m->h_sizes[1];
  const double t_sel = now_ms();
  m->last_C = C; m->last_L = L; m->last_N = N; m->last_cam_kf = cam_kf;
  if (sizes3_out) { sizes3_out[0] = C; sizes3_out[1] = L; sizes3_out[2] = N; }
  int rc = SLSLAM_OK;
  slslam_summary summ; memset(&summ, 0, sizeof(summ));
  if (L > 0 && N > 0) {
    map_emit_kernel<<<(n_cand + 255) / 256, 256, 0, st>>>(d);
    map_params_kernel<<<(C + L + 127) / 128, 128, 0, st>>>(d, d_cam_kf, L);
    CUDA_TRY(cudaMemcpyAsync(m->w_params0, d.w_params, 8 * (size_t)(6 * C + 4 * L), cudaMemcpyDeviceToDevice, st));
    slslam_lba_desc desc; memset(&desc, 0, sizeof(desc));
    desc.num_cameras = C; desc.num_lines = L; desc.num_observations = N; desc.max_iterations = max_iterations;
    desc.camera_index = d.w_cam; desc.line_index = d.w_line; desc.fixed_index = d.w_fixed; desc.observations = d.w_obs;
    desc.robust = robust; desc.huber_delta = 0.0; desc.baseline = -1.0;
    double* pp[1] = {d.w_params};
    rc = slslam_lba_solve_batch_device(1, &desc, pp, nullptr, &summ, st);
    if (rc == SLSLAM_OK) {
      map_writeback_cams_kernel<<<(C + 127) / 128, 128, 0, st>>>(d, d_cam_kf);
      map_writeback_lines_kernel<<<(L + 127) / 128, 128, 0, st>>>(d, L);
    }
  }
  map_reset_kernel<<<(std::max(C, d.max_lm) + 255) / 256, 256, 0, st>>>(d, d_cam_kf);
  cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); cudaGetLastError(); return SLSLAM_ERR_CUDA; }
  if (summary_out) *summary_out = summ;
  const double t_end = now_ms();
  m->tm.assemble_ms = t_sel - t_begin; m->tm.solve_and_writeback_ms = t_end - t_sel; m->tm.total_ms = t_end - t_begin;
  m->tm.h2d_bytes = (int64_t)bytes; m->tm.candidates = n_cand;
  return rc;
}

void slslam_map_last_timings(const slslam_map* m, slslam_map_timings* out) { if (m && out) *out = m->tm; }

// Diagnostics / parity: the window the last slslam_map_bundle_adjust assembled, in the reference's array layout, with the
// parameters as assembled (before the solve).  Any pointer may be NULL.  line_landmark [L], camera_keyframe [C].
int slslam_map_last_window(slslam_map* m, int32_t* camera_index, int32_t* line_index, int32_t* fixed_index, double* observations,
                           double* parameters, int32_t* line_landmark, int32_t* camera_keyframe) {
  set_last_error("");
  if (!m) return SLSLAM_ERR_INVALID;
  cudaSetDevice(m->device);
  const size_t N = (size_t)m->last_N, L = (size_t)m->last_L, C = (size_t)m->last_C;
  if (camera_index && N) CUDA_TRY(cudaMemcpy(camera_index, m->d.w_cam, 4 * N, cudaMemcpyDeviceToHost));
  if (line_index && N) CUDA_TRY(cudaMemcpy(line_index, m->d.w_line, 4 * N, cudaMemcpyDeviceToHost));
  if (fixed_index && N) CUDA_TRY(cudaMemcpy(fixed_index, m->d.w_fixed, 8 * N, cudaMemcpyDeviceToHost));
  if (observations && N) CUDA_TRY(cudaMemcpy(observations, m->d.w_obs, 64 * N, cudaMemcpyDeviceToHost));
  if (parameters && (C || L)) CUDA_TRY(cudaMemcpy(parameters, m->w_params0, 8 * (6 * C + 4 * L), cudaMemcpyDeviceToHost));
  if (line_landmark && L) CUDA_TRY(cudaMemcpy(line_landmark, m->d.w_line_lm, 4 * L, cudaMemcpyDeviceToHost));
  if (camera_keyframe) for (size_t i = 0; i < C; ++i) camera_keyframe[i] = m->last_cam_kf[i];
  return SLSLAM_OK;
}

// Batch conversions on the device (gc_av_to_orth / gc_orth_to_av, src/gc.cpp:361-460; the rotation <-> angle-axis pair
// behind gc_Rt_to_wt / gc_wt_to_Rt): mode 0 av[6n] -> orth[4n], 1 orth[4n] -> av[6n], 2 R[9n] row-major -> w[3n], 3 w -> R.
int slslam_geometry_convert(int32_t mode, int32_t n, const double* in, double* out) {
  set_last_error("");
  if (mode < 0 || mode > 3 || n < 0 || (n > 0 && (!in || !out))) return SLSLAM_ERR_INVALID;
  int rc = ensure_device(-1);
  if (rc != SLSLAM_OK) return rc;
  if (n == 0) return SLSLAM_OK;
  const int win[4] = {6, 4, 9, 3}, wout[4] = {4, 6, 3, 9};
  double *di = nullptr, *dout = nullptr;
  CUDA_TRY(cudaMalloc((void**)&di, (size_t)8 * win[mode] * n));
  if (cudaMalloc((void**)&dout, (size_t)8 * wout[mode] * n) != cudaSuccess) { cudaGetLastError(); cudaFree(di); return SLSLAM_ERR_CUDA; }
  cudaMemcpy(di, in, (size_t)8 * win[mode] * n, cudaMemcpyHostToDevice);
  map_convert_kernel<<<(n + 127) / 128, 128>>>(n, mode, di, dout);
  cudaError_t e = cudaMemcpy(out, dout, (size_t)8 * wout[mode] * n, cudaMemcpyDeviceToHost);
  cudaFree(di); cudaFree(dout);
  if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); cudaGetLastError(); return SLSLAM_ERR_CUDA; }
  return SLSLAM_OK;
}

}  // extern "C"
