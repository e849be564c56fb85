// Host side of the pose-graph path behind the C ABI (include/slslam_b200.h): validation, the symbolic plan
// (constant pose, reduced block indices, incident-edge CSR, the block structure of J^T J with its per-block
// contribution lists), upload, the stream-ordered LM loop (no host round trip inside), download.
// Replaces what POProblem::build + ceres::Solve do (reference src/po_problem.cpp:40-77, src/slam.cpp:1283-1293).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "common_host.h"
#include "po_kernels.cuh"

namespace slslam {

static thread_local float g_po_last_ms = 0.f;

struct PoPlan {
  int K = 0, E = 0, Kf = 0, n = 0;
  std::vector<int> slot, slot_pose, inc_off, inc, blk_i, blk_j, blk_off, contrib;
  std::vector<unsigned char> active;
};

static int validate_po(const slslam_po_desc& d) {
  if (d.num_poses < 0 || d.num_edges < 0 || d.max_iterations < 0) return SLSLAM_ERR_INVALID;
  if (d.num_edges > 0 && (!d.pose_index_1 || !d.pose_index_2 || !d.constraints)) return SLSLAM_ERR_INVALID;
  for (int e = 0; e < d.num_edges; ++e) {
    if (d.pose_index_1[e] < 0 || d.pose_index_1[e] >= d.num_poses) return SLSLAM_ERR_INVALID;
    if (d.pose_index_2[e] < 0 || d.pose_index_2[e] >= d.num_poses) return SLSLAM_ERR_INVALID;
  }
  for (size_t i = 0; i < 6 * (size_t)d.num_edges; ++i) if (!std::isfinite(d.constraints[i])) return SLSLAM_ERR_NUMERICAL;
  return SLSLAM_OK;
}

static void build_po_plan(const slslam_po_desc& d, bool all_free, PoPlan& p) {
  const int K = d.num_poses, E = d.num_edges;
  p.K = K; p.E = E;
  std::vector<char> used(K, 0);
  for (int e = 0; e < E; ++e) { used[d.pose_index_1[e]] = 1; used[d.pose_index_2[e]] = 1; }
  // pose1 of the first edge is held constant (reference src/po_problem.cpp:62-63); unused poses are never touched
  const int konst = (E > 0 && !all_free) ? d.pose_index_1[0] : -1;
  p.slot.assign(K, -1);
  p.Kf = 0;
  for (int k = 0; k < K; ++k) if (used[k] && k != konst) { p.slot[k] = p.Kf++; p.slot_pose.push_back(k); }
  p.n = 6 * p.Kf;
  p.active.assign(E, 0);
  std::vector<std::vector<int> > inc(p.Kf);
  std::map<std::pair<int, int>, std::vector<int> > blocks;
  for (int e = 0; e < E; ++e) {
    const int a = d.pose_index_1[e], b = d.pose_index_2[e];
    const int s1 = p.slot[a], s2 = p.slot[b];
    p.active[e] = (s1 >= 0 || s2 >= 0) ? 1 : 0;
    if (s1 >= 0) { inc[s1].push_back(e << 1); blocks[std::make_pair(s1, s1)].push_back(e << 1); }
    if (a != b && s2 >= 0) { inc[s2].push_back(e << 1 | 1); blocks[std::make_pair(s2, s2)].push_back(e << 1 | 1); }
    if (a != b && s1 >= 0 && s2 >= 0) {
      if (s2 > s1) blocks[std::make_pair(s2, s1)].push_back(e << 1 | 1);
      else blocks[std::make_pair(s1, s2)].push_back(e << 1);
    }
  }
  p.inc_off.assign(p.Kf + 1, 0);
  for (int s = 0; s < p.Kf; ++s) {
    p.inc_off[s + 1] = p.inc_off[s] + (int)inc[s].size();
    p.inc.insert(p.inc.end(), inc[s].begin(), inc[s].end());
  }
  p.blk_off.push_back(0);
  for (auto& kv : blocks) {
    p.blk_i.push_back(kv.first.first); p.blk_j.push_back(kv.first.second);
    p.contrib.insert(p.contrib.end(), kv.second.begin(), kv.second.end());
    p.blk_off.push_back((int)p.contrib.size());
  }
}

// One pooled device allocation; `take` hands out 256-byte aligned slices.
struct Pool {
  char* base = nullptr;
  size_t off = 0;
  size_t reserve(size_t bytes) { size_t r = off; off += (bytes + 255) & ~(size_t)255; return r; }
};

static int po_run(const slslam_po_desc* desc, const double* poses_in, double* poses_out, slslam_summary* summary_out,
                  double* trace_out, bool evaluate_only, double* res_out, double* j1_out, double* j2_out, double* cost_out) {
  int rc = validate_po(*desc);
  if (rc != SLSLAM_OK) return rc;
  const int K = desc->num_poses, E = desc->num_edges;
  for (size_t i = 0; i < 6 * (size_t)K; ++i) if (!std::isfinite(poses_in[i])) return SLSLAM_ERR_NUMERICAL;
  rc = ensure_device(-1);
  if (rc != SLSLAM_OK) return rc;
  PoPlan p;
  build_po_plan(*desc, evaluate_only, p);
  const int n = p.n, M = n + 1, ld = (M + 7) & ~7, nblk = (int)p.blk_i.size();
  const int max_iters = desc->max_iterations;
  const int nb32 = (n + PO_NB - 1) / PO_NB;

  Pool pool;
  const size_t Ez = std::max(E, 1), Kz = std::max(K, 1), nz = std::max(n, 1);
  const size_t o_idx1 = pool.reserve(4 * Ez), o_idx2 = pool.reserve(4 * Ez), o_cons = pool.reserve(48 * Ez);
  const size_t o_slot = pool.reserve(4 * Kz), o_spose = pool.reserve(4 * nz), o_act = pool.reserve(Ez);
  const size_t o_incoff = pool.reserve(4 * (p.inc_off.size() + 1)), o_inc = pool.reserve(4 * (p.inc.size() + 1));
  const size_t o_bi = pool.reserve(4 * (size_t)(nblk + 1)), o_bj = pool.reserve(4 * (size_t)(nblk + 1));
  const size_t o_boff = pool.reserve(4 * (size_t)(nblk + 2)), o_contrib = pool.reserve(4 * (p.contrib.size() + 1));
  const size_t o_x = pool.reserve(48 * Kz);
  const size_t o_state = pool.reserve(sizeof(PoState));
  const size_t upload_end = pool.off;
  const size_t o_xt = pool.reserve(48 * Kz);
  const size_t o_r = pool.reserve(2 * 48 * Ez), o_J1 = pool.reserve(2 * 288 * Ez), o_J2 = pool.reserve(2 * 288 * Ez);
  const size_t o_ce = pool.reserve(2 * 8 * Ez), o_mval = pool.reserve(8 * Ez);
  const size_t o_scale = pool.reserve(8 * nz), o_cn = pool.reserve(8 * nz), o_g = pool.reserve(8 * nz), o_y = pool.reserve(8 * nz);
  const size_t o_summ = pool.reserve(sizeof(slslam_summary));
  const size_t o_trace = pool.reserve(8 * (size_t)SLSLAM_TRACE_WIDTH * std::max(max_iters, 1));
  const size_t o_Ld = pool.reserve(8 * (size_t)PO_NB * PO_NB * std::max(nb32, 1));
  const size_t o_flags = pool.reserve(4 * (size_t)std::max(nb32, 1));
  const size_t o_H = evaluate_only ? pool.off : pool.reserve(8 * (size_t)M * ld);
  CUDA_TRY(cudaMalloc((void**)&pool.base, pool.off));

  std::vector<char> host(upload_end, 0);
  if (E > 0) {
    memcpy(host.data() + o_idx1, desc->pose_index_1, 4 * (size_t)E);
    memcpy(host.data() + o_idx2, desc->pose_index_2, 4 * (size_t)E);
    memcpy(host.data() + o_cons, desc->constraints, 48 * (size_t)E);
    memcpy(host.data() + o_act, p.active.data(), (size_t)E);
  }
  if (K > 0) { memcpy(host.data() + o_slot, p.slot.data(), 4 * (size_t)K); memcpy(host.data() + o_x, poses_in, 48 * (size_t)K); }
  if (p.Kf > 0) memcpy(host.data() + o_spose, p.slot_pose.data(), 4 * (size_t)p.Kf);
  memcpy(host.data() + o_incoff, p.inc_off.data(), 4 * p.inc_off.size());
  if (!p.inc.empty()) memcpy(host.data() + o_inc, p.inc.data(), 4 * p.inc.size());
  if (nblk > 0) {
    memcpy(host.data() + o_bi, p.blk_i.data(), 4 * (size_t)nblk);
    memcpy(host.data() + o_bj, p.blk_j.data(), 4 * (size_t)nblk);
    memcpy(host.data() + o_contrib, p.contrib.data(), 4 * p.contrib.size());
  }
  memcpy(host.data() + o_boff, p.blk_off.data(), 4 * p.blk_off.size());
  PoState st; memset(&st, 0, sizeof(st));
  st.radius = desc->initial_trust_region_radius > 0 ? desc->initial_trust_region_radius : 1e4;   // Ceres 1.7.0 defaults
  st.decrease_factor = 2.0;
  st.ftol = desc->function_tolerance > 0 ? desc->function_tolerance : 1e-6;
  st.gtol = desc->gradient_tolerance > 0 ? desc->gradient_tolerance : 1e-10;
  st.ptol = desc->parameter_tolerance > 0 ? desc->parameter_tolerance : 1e-8;
  st.max_iters = max_iters; st.term = SLSLAM_NO_CONVERGENCE;
  memcpy(host.data() + o_state, &st, sizeof(st));

  PoDev d; memset(&d, 0, sizeof(d));
  char* B = pool.base;
  d.K = K; d.E = E; d.n = n; d.M = M; d.ld = ld; d.nblk = nblk;
  d.idx1 = (const int*)(B + o_idx1); d.idx2 = (const int*)(B + o_idx2); d.cons = (const double*)(B + o_cons);
  d.slot = (const int*)(B + o_slot); d.slot_pose = (const int*)(B + o_spose); d.active = (const unsigned char*)(B + o_act);
  d.inc_off = (const int*)(B + o_incoff); d.inc = (const int*)(B + o_inc);
  d.blk_i = (const int*)(B + o_bi); d.blk_j = (const int*)(B + o_bj); d.blk_off = (const int*)(B + o_boff);
  d.contrib = (const int*)(B + o_contrib);
  d.x = (double*)(B + o_x); d.xt = (double*)(B + o_xt);
  d.r = (double*)(B + o_r); d.J1 = (double*)(B + o_J1); d.J2 = (double*)(B + o_J2); d.cost_e = (double*)(B + o_ce);
  d.scale = (double*)(B + o_scale); d.cn = (double*)(B + o_cn); d.g = (double*)(B + o_g); d.y = (double*)(B + o_y);
  d.mval = (double*)(B + o_mval);
  d.H = (double*)(B + o_H); d.Ld = (double*)(B + o_Ld);
  d.st = (PoState*)(B + o_state);
  d.trace = trace_out ? (double*)(B + o_trace) : nullptr;
  d.summary = (slslam_summary*)(B + o_summ);

  cudaStream_t s = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  unsigned int* d_flags = (unsigned int*)(B + o_flags);
  int bs_ctas = 1;
  rc = SLSLAM_OK;
#define PO_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { set_last_error(cudaGetErrorString(e_)); cudaGetLastError(); rc = SLSLAM_ERR_CUDA; goto done; } } while (0)
  PO_TRY(cudaMemcpyAsync(B, host.data(), upload_end, cudaMemcpyHostToDevice, s));
  if (trace_out) PO_TRY(cudaMemsetAsync(B + o_trace, 0, 8 * (size_t)SLSLAM_TRACE_WIDTH * std::max(max_iters, 1), s));
  PO_TRY(cudaEventCreate(&ev0)); PO_TRY(cudaEventCreate(&ev1));
  PO_TRY(cudaEventRecord(ev0, s));
  if (E > 0) po_linearize<<<(E + 127) / 128, 128, 0, s>>>(d, 0, 1, evaluate_only ? 1 : 0);
  if (evaluate_only) {
    PO_TRY(cudaGetLastError());
    PO_TRY(cudaStreamSynchronize(s));
    if (E > 0) {
      PO_TRY(cudaMemcpy(res_out, d.r, 48 * (size_t)E, cudaMemcpyDeviceToHost));
      if (j1_out) PO_TRY(cudaMemcpy(j1_out, d.J1, 288 * (size_t)E, cudaMemcpyDeviceToHost));
      if (j2_out) PO_TRY(cudaMemcpy(j2_out, d.J2, 288 * (size_t)E, cudaMemcpyDeviceToHost));
    }
    if (cost_out) {
      std::vector<double> ce(Ez, 0.0);
      if (E > 0) PO_TRY(cudaMemcpy(ce.data(), d.cost_e, 8 * (size_t)E, cudaMemcpyDeviceToHost));
      double c = 0.0;
      for (int e = 0; e < E; ++e) c += ce[e];
      *cost_out = c;
    }
    goto done;
  }
  if (!evaluate_only && n > 0) {
    // back-substitution spreads its 32-column blocks over co-resident CTAs (cooperative launch)
    int dev = 0, sms = 1, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, po_backsolve, 256, 0) != cudaSuccess) { cudaGetLastError(); per_sm = 1; }
    bs_ctas = std::max(1, std::min(nb32, sms * std::max(per_sm, 1)));
    bs_ctas = std::min(bs_ctas, 64);
    if ((nb32 + bs_ctas - 1) / bs_ctas > PO_BS_MAXOWN) { set_last_error("pose graph too large for the back-substitution kernel"); rc = SLSLAM_ERR_UNSUPPORTED; goto done; }
    PO_TRY(cudaMemsetAsync(d_flags, 0, 4 * (size_t)std::max(nb32, 1), s));
  }
  if (n > 0) po_colnorm_grad<<<(n + 127) / 128, 128, 0, s>>>(d, 0);
  po_refresh<<<1, 256, 0, s>>>(d, 1);
  if (n > 0) {
    for (int it = 0; it < max_iters; ++it) {
      PO_TRY(cudaMemsetAsync(d.H, 0, 8 * (size_t)M * ld, s));
      po_assemble<<<(nblk * 36 + n + 255) / 256, 256, 0, s>>>(d);
      for (int k0 = 0; k0 < n; k0 += PO_NB) {
        const int nb = std::min(PO_NB, n - k0), t0 = k0 + nb;
        po_chol_panel<<<1 + (M - t0 + PO_TR - 1) / PO_TR, PO_TR, 0, s>>>(d, k0);
        if (t0 < n) {
          const int T = (M - t0 + PO_TS - 1) / PO_TS;
          po_chol_syrk<<<T * (T + 1) / 2, 256, 0, s>>>(d, k0, nb, t0, T);
        }
      }
      {
        unsigned int gen = (unsigned int)(it + 1);
        void* args[3] = {(void*)&d, (void*)&d_flags, (void*)&gen};
        PO_TRY(cudaLaunchCooperativeKernel((const void*)po_backsolve, dim3((unsigned)bs_ctas), dim3(256), args, 0, s));
      }
      po_step<<<(6 * K + E + 255) / 256, 256, 0, s>>>(d);
      po_linearize<<<(E + 127) / 128, 128, 0, s>>>(d, 1, 0, 0);
      po_decide<<<1, 256, 0, s>>>(d);
      po_accept<<<(6 * K + 255) / 256, 256, 0, s>>>(d);
      po_colnorm_grad<<<(n + 127) / 128, 128, 0, s>>>(d, 1);
      po_refresh<<<1, 256, 0, s>>>(d, 0);
    }
  }
  po_finish<<<1, 1, 0, s>>>(d);
  PO_TRY(cudaGetLastError());
  PO_TRY(cudaEventRecord(ev1, s));
  {
    std::vector<double> xo((size_t)6 * Kz);
    slslam_summary summ;
    PO_TRY(cudaMemcpyAsync(xo.data(), d.x, 48 * (size_t)Kz, cudaMemcpyDeviceToHost, s));
    PO_TRY(cudaMemcpyAsync(&summ, d.summary, sizeof(summ), cudaMemcpyDeviceToHost, s));
    if (trace_out) PO_TRY(cudaMemcpyAsync(trace_out, d.trace, 8 * (size_t)SLSLAM_TRACE_WIDTH * std::max(max_iters, 0), cudaMemcpyDeviceToHost, s));
    PO_TRY(cudaStreamSynchronize(s));
    PO_TRY(cudaEventElapsedTime(&g_po_last_ms, ev0, ev1));
    // parameters are only overwritten once everything has succeeded
    if (K > 0) memcpy(poses_out, xo.data(), 48 * (size_t)K);
    if (summary_out) *summary_out = summ;
  }
done:
#undef PO_TRY
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  cudaFree(pool.base);
  return rc;
}

}  // namespace slslam

using namespace slslam;

extern "C" {

int slslam_po_solve_trace(const slslam_po_desc* desc, double* poses_inout, slslam_summary* summary_out, double* trace_out) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!desc || (!poses_inout && desc->num_poses > 0)) return SLSLAM_ERR_INVALID;
  return po_run(desc, poses_inout, poses_inout, summary_out, trace_out, false, nullptr, nullptr, nullptr, nullptr);
}

int slslam_po_solve(const slslam_po_desc* desc, double* poses_inout, slslam_summary* summary_out) {
  return slslam_po_solve_trace(desc, poses_inout, summary_out, nullptr);
}

int slslam_po_evaluate(const slslam_po_desc* desc, const double* poses, double* residuals, double* jac_pose1,
                       double* jac_pose2, double* cost_out) {
  slslam::set_last_error("");   // a message left by an earlier call must not be attached to this one
  if (!desc || !poses || !residuals) return SLSLAM_ERR_INVALID;
  return po_run(desc, poses, nullptr, nullptr, nullptr, true, residuals, jac_pose1, jac_pose2, cost_out);
}

float slslam_po_last_solve_ms(void) { return g_po_last_ms; }

}  // extern "C"
